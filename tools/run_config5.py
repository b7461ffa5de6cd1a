"""BASELINE configs[4] (CTC-CRF training step, fwd_bwd): sup@v3.3 UB X, batch 512 x 4000 samples, targets ~U(350, 450)
bases with ~9 % X spliced in (>= 5 bases apart), per-GPU data-parallel replica (no all-reduce in the timed region).

Times, with CUDA events:  forward + loss;  + loss backward to the scores;  + encoder backward to all 28 parameter tensors;
+ clip_grad_norm_(2.0) + AdamW step -- the whole of Trainer.train_one_step (bonito/training.py:91-117), through the plugin
classes (xna_basecaller_b200.crf.Model, xna_basecaller_b200.training.Trainer).
    python tools/run_config5.py [N] [bf16]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bonito_oracle as bo            # weight generator only
from xna_basecaller_b200 import util
from xna_basecaller_b200.training import Trainer

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = {'global_norm': {'state_len': 3}, 'input': {'features': 1}, 'labels': {'labels': list('NACGTX')},
       'model': {'package': 'xna_basecaller_b200.crf'},
       'encoder': {'stride': 5, 'activation': 'swish', 'features': 768, 'winlen': 19, 'scale': 5.0, 'rnn_type': 'lstm',
                   'blank_score': 2.0}}
model = util.load_symbol(cfg, 'Model')(cfg)
model.load_state_dict(bo.reference_state_dict(n_base=5, seed=25))
trainer = Trainer(model, 'cuda')
trainer.init_optimizer(2e-3)
model.train()
x = torch.randn(N, 1, 4000, generator=torch.Generator().manual_seed(1234)).cuda()
rs = np.random.RandomState(3)
lens = rs.randint(350, 451, N)
tg = np.zeros((N, int(lens.max())), dtype=np.int64)
for i, L in enumerate(lens):
    seq = rs.randint(1, 5, L)
    cand = np.arange(5, L - 5, 11)
    seq[cand[rs.rand(len(cand)) < 0.99]] = 5                  # ~9 % X, >= 5 bases apart
    tg[i, :L] = seq
tg, tl = torch.from_numpy(tg).cuda(), torch.from_numpy(lens).cuda()


def fwd_loss():
    with torch.no_grad():
        model.eval()
        loss = model.seqdist.ctc_loss(model(x).float(), tg, tl)
        model.train()
    return loss


def fwd_bwd():
    for p in model.parameters():
        p.grad = None
    loss = model.seqdist.ctc_loss(model(x).float(), tg, tl)
    loss.backward()
    return loss


def full_step():
    losses, norm = trainer.train_one_step((x, tg, tl))
    return losses['loss']


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


out = {'config': 'configs[4]: training step fwd_bwd, batch %d x 4000 samples, targets 350-450 bases (~9%% X); fp16 forward '
                 'operands, bf16 gradient transport, fp32 accumulation and master weights' % N}
for name, fn, reps in (('fwd_loss', fwd_loss, 5), ('fwd_loss_bwd_to_all_parameters', fwd_bwd, 3), ('train_one_step', full_step, 3)):
    ms, val = timed(fn, reps)
    out[name] = {'ms_per_step': ms, 'samples_per_s': N * 4000 / ms * 1e3, 'loss': float(val)}
h = model.seqdist.engine.handle
h.set_profiling(True)
h.stage_times()
fwd_bwd()
torch.cuda.synchronize()
out['stages_ms_of_one_fwd_bwd'] = {k: v[0] for k, v in h.stage_times().items()}
print(json.dumps(out))
