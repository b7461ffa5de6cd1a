"""BASELINE configs[4] (training step, forward part + loss gradient): sup@v3.3 UB X, bf16 encoder, batch 512 x 4000 samples,
targets ~U(350, 450) bases with ~9 % X spliced in.  Times encoder forward + CTC-CRF loss forward, and additionally the loss
backward to the scores (xb_ctc_crf_loss_bwd).  The encoder backward (parameter gradients) is not implemented.
    python tools/run_config5.py [N]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bonito_oracle as bo            # weight generator only
from xna_basecaller_b200._lib import Handle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = 800
h = Handle('NACGTX', 3, max_N=N, max_T=T, bf16=True)
h.load_weights(bo.reference_state_dict(n_base=5, seed=25))
x = torch.randn(N, 4000, generator=torch.Generator().manual_seed(1234)).cuda()
rs = np.random.RandomState(3)
lens = rs.randint(350, 451, N)
tg = np.zeros((N, int(lens.max())), dtype=np.int64)
for i, L in enumerate(lens):
    seq = rs.randint(1, 5, L)
    pos = np.arange(5, L - 5, 11)[rs.rand(len(np.arange(5, L - 5, 11))) < 0.99]      # ~9 % X, >= 5 bases apart
    seq[pos] = 5
    tg[i, :L] = seq
tg, tl = torch.from_numpy(tg).cuda(), torch.from_numpy(lens).cuda()
w = torch.full((N,), 1.0 / N, device='cuda')


def step(backward):
    s = h.encoder(x)
    loss = h.ctc_loss(s, tg, tl)
    g = h.ctc_loss_bwd(s, tg, tl, w) if backward else None
    return loss, g


out = {}
for name, bwd in (('fwd_loss', False), ('fwd_loss_plus_loss_bwd', True)):
    for _ in range(2):
        loss, g = step(bwd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        loss, g = step(bwd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[name] = {'ms_per_step': ms, 'samples_per_s': N * 4000 / ms * 1e3}
out['loss_mean'] = float(loss.mean())
out['grad_abs_sum_per_step'] = float(g.abs().sum(2).mean())
out['config'] = 'configs[4]: bf16 encoder forward + CTC-CRF loss, batch %d x 4000 samples, targets 350-450 bases (~9%% X)' % N
print(json.dumps(out))
