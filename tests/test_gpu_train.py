"""GPU parity of the training step (BASELINE configs[4], fwd_bwd): the encoder backward (xb_encoder_fwd_train /
xb_encoder_bwd) against torch autograd through the fp32 oracle restatement of the reference modules
(bonito/training.py:91-117 runs exactly that autograd).

Tolerance: gradients travel as bf16 and the forward runs on fp16 operands, so every gradient tensor is compared by its
relative error in the 2-norm (<= 3e-2) and by the cosine of the angle to the reference (>= 0.999)."""
import numpy as np
import pytest
import torch

from make_golden import ALPHABETS, REF_SCALE, synthetic_signal
from oracle import bonito_oracle as bo

pytestmark = pytest.mark.gpu

REL_TOL, COS_TOL = 3e-2, 0.999


def _reference_grads(sd, x, cotangent, n_base):
    params = {k: v.clone().requires_grad_() for k, v in sd.items()}
    scores = bo.encoder_forward(params, x, n_base)
    scores.backward(cotangent)
    return scores.detach(), {k: p.grad for k, p in params.items()}


def _compare(got, ref, skip=()):
    worst = {}
    for k, g in ref.items():
        if k in skip or g is None:
            continue
        a, b = got[k].double().cpu().flatten(), g.double().flatten()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        cos = (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)).item()
        worst[k] = (rel, cos)
    return worst


@pytest.mark.parametrize('n_base,N,L', [(5, 8, 500), (6, 16, 300)])
def test_encoder_backward_matches_autograd(n_base, N, L):
    from xna_basecaller_b200._lib import Handle
    sd = bo.reference_state_dict(n_base=n_base, seed=12, **REF_SCALE)
    x = synthetic_signal(41, N, L)
    T = L // 5
    g = torch.Generator().manual_seed(9)
    C, NZ = n_base ** 3, n_base + 1
    cot = torch.randn(T, N, C * NZ, generator=g) * 1e-2
    ref_scores, ref = _reference_grads(sd, x, cot, n_base)
    h = Handle(ALPHABETS[n_base], 3, max_N=N, max_T=T, train=True)
    h.load_weights(sd)
    scores = h.encoder_train(x.cuda())
    assert (scores.cpu() - ref_scores).abs().max().item() < 1e-2
    # the training forward is the inference forward plus stores: same bits
    assert torch.equal(scores, h.encoder(x.cuda()))
    got = h.encoder_backward(cot.cuda())
    torch.cuda.synchronize()
    worst = _compare(got, ref, skip=[k for k in ref if k.endswith('bias_hh_l0')])
    for k, (rel, cos) in sorted(worst.items()):
        print('%-32s rel %.4f  cos %.6f' % (k, rel, cos))
    for k, (rel, cos) in worst.items():
        assert rel <= REL_TOL and cos >= COS_TOL, (k, rel, cos)
    # bias_hh receives the gradient of bias_ih (they enter the gates as a sum)
    for layer in range(4, 9):
        assert torch.equal(got['encoder.%d.rnn.bias_hh_l0' % layer], got['encoder.%d.rnn.bias_ih_l0' % layer])
    h.close()
