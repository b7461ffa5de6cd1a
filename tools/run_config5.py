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


# The reference's PyTorch path on THIS GPU for the same step, as far as it can be restated without the seqdist wheel: the module
# graph through torch autograd under fp16 autocast (cuDNN LSTM forward + backward, cuBLAS, cuDNN convolutions), fp32 master
# weights, a fixed cotangent on the scores in place of the loss backward.  Encoder only: a LOWER bound on the reference's step.
def torch_encoder_fwd_bwd():
    sd = model.state_dict()
    params = {k: v.detach().clone().cuda().requires_grad_() for k, v in sd.items() if k.startswith('encoder.')}
    lstms = []
    for layer in range(4, 9):
        m = torch.nn.LSTM(768, 768).cuda()
        with torch.no_grad():
            for name in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0'):
                getattr(m, name).copy_(sd['encoder.%d.rnn.%s' % (layer, name)])
        m.flatten_parameters()
        lstms.append(m)
    cot = torch.randn(800, N, 750, generator=torch.Generator().manual_seed(2)).cuda() * 1e-3

    def step():
        for p_ in list(params.values()) + [q for m in lstms for q in m.parameters()]:
            p_.grad = None
        with torch.autocast('cuda', dtype=torch.float16):
            y = bo.conv_stem(params, x).permute(2, 0, 1).contiguous()
            for m, rev in zip(lstms, bo.LSTM_DIRECTIONS):
                y = m(y.flip(0))[0].flip(0) if rev else m(y)[0]
            scores = bo.crf_head(params, y, 5)
        scores.float().backward(cot)
        return scores.detach().float().mean()
    return timed(step, 3)


try:
    ms, _ = torch_encoder_fwd_bwd()
    out['torch_cudnn_encoder_fwd_bwd_same_gpu'] = {'ms_per_step': ms, 'samples_per_s': N * 4000 / ms * 1e3,
                                                   'what': 'torch %s autograd, fp16 autocast, cuDNN LSTM: encoder forward + '
                                                           'backward only (no loss, no optimiser)' % torch.__version__}
except Exception as e:                                   # comparator only: never fail the measurement of the product
    out['torch_cudnn_encoder_fwd_bwd_same_gpu'] = {'error': repr(e)[:200]}
torch.cuda.empty_cache()
h = model.seqdist.engine.handle
h.set_profiling(True)
h.stage_times()
fwd_bwd()
torch.cuda.synchronize()
out['stages_ms_of_one_fwd_bwd'] = {k: v[0] for k, v in h.stage_times().items()}
print(json.dumps(out))
