"""GPU parity at BASELINE size (configs[1]: N = 512 chunks x 4000 samples, T = 800, n_base 5).

* score-level parity: the CRF scores of sampled chunks of the full batch against the fp32 oracle run on exactly those
  chunks (chunks are independent), for fp16 weights and for bf16 weights, against the north-star tolerance
  (max abs error <= 1e-2 on scores in +-5).  800 recurrent steps x 5 layers is where 16-bit rounding and
  tanh.approx would accumulate if they did;
* decode parity at the same size: labels of sampled chunks bit-equal to the C oracle on the same fp32 scores;
* determinism: the h hand-off between the 24 CTAs of an LSTM group goes through L2 with a release / acquire pair;
  a missed hand-off would read a stale h tile and silently change basecalls.  The same batch is pushed through
  encoder + decode 100 times and every repetition must reproduce the first one bit for bit.
"""
import numpy as np
import pytest
import torch

from make_golden import ALPHABETS, REF_SCALE, synthetic_signal
from oracle import bonito_oracle as bo
from oracle import cexact

pytestmark = pytest.mark.gpu

N_FULL, L_FULL = 512, 4000
SCORE_TOL = 1e-2              # BASELINE.json north_star: max abs error on CRF scores, 16-bit operands vs fp32 reference
PICK = [0, 31, 32, 95, 96, 300, 480, 511]   # first / last chunk of sub-batches and groups of the persistent LSTM


def _weights(head_gain=1.0):
    sd = bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE)
    sd['encoder.9.linear.weight'] = sd['encoder.9.linear.weight'] * head_gain
    return sd


def _mixed_decode_handle(x):
    """Handle whose head gain gives decodes that mix blanks and moves (a random-init head decodes to the empty
    string, a strongly amplified one emits a base at every step; both make decode parity vacuous)."""
    from xna_basecaller_b200._lib import Handle
    h = Handle(ALPHABETS[5], 3, max_N=N_FULL, max_T=L_FULL // 5)
    lo, hi, seen = 0.02, 8.0, []                     # decoded length grows with the gain: bisect in log space
    for _ in range(14):
        gain = (lo * hi) ** 0.5
        h.load_weights(_weights(gain))
        _, _, lens = h.decode(h.encoder(x[:16]), want_qstring=False)
        mean = lens.float().mean().item()
        seen.append((round(gain, 3), mean))
        if 100 < mean < 700:
            return h, gain, mean
        lo, hi = (gain, hi) if mean <= 100 else (lo, gain)
    raise AssertionError('no head gain gives a mixed decode: %s' % seen)


@pytest.fixture(scope='module')
def full_batch():
    return synthetic_signal(77, N_FULL, L_FULL)


@pytest.mark.parametrize('bf16', [False, True])
def test_scores_at_baseline_size(full_batch, bf16):
    from xna_basecaller_b200._lib import Handle
    sd = _weights()
    h = Handle(ALPHABETS[5], 3, max_N=N_FULL, max_T=L_FULL // 5, bf16=bf16)
    h.load_weights(sd)
    scores = h.encoder(full_batch.cuda())
    got = scores[:, PICK].cpu()
    with torch.no_grad():
        ref = bo.encoder_forward(sd, full_batch[PICK], 5)
    err = (got - ref).abs()
    per_chunk = err.amax(dim=(0, 2))
    print('%s weights, N=%d T=%d: max |score err| per sampled chunk %s; mean %.2e'
          % ('bf16' if bf16 else 'fp16', N_FULL, L_FULL // 5, ['%.4f' % v for v in per_chunk.tolist()], err.mean().item()))
    assert err.max().item() <= SCORE_TOL
    # error must not grow along the sequence (800 recurrent steps): compare the first and last 100 steps
    assert err[-100:].max().item() <= SCORE_TOL and err[:100].max().item() <= SCORE_TOL
    h.close()


def test_decode_at_baseline_size(full_batch):
    """Labels and packed rows of sampled chunks of the full batch == C oracle on the same fp32 scores; the whole batch
    obeys the size-independent properties (lengths = count of non-blank labels, rows left-packed, letters in alphabet)."""
    h, gain, _ = _mixed_decode_handle(full_batch.cuda())
    scores = h.encoder(full_batch.cuda())
    seq, _, lens, labels = h.decode(scores, want_labels=True, want_qstring=False)
    torch.cuda.synchronize()
    want = cexact.crf_decode(scores[:, PICK].cpu().numpy(), 5)
    assert np.array_equal(labels[PICK].cpu().numpy(), want)
    lab, seqn, lensn = labels.cpu().numpy(), seq.cpu().numpy(), lens.cpu().numpy()
    assert np.array_equal(lensn, (lab != 0).sum(1))
    letters = np.frombuffer(b'NACGTX', dtype='u1')
    for i in range(N_FULL):
        row = lab[i][lab[i] != 0]
        assert np.array_equal(seqn[i, :lensn[i]].astype('u1'), letters[row])
        assert not seqn[i, lensn[i]:].any()
    assert 100 < lensn.mean() < 700, 'degenerate decodes: the parity above would be vacuous'
    print('head gain %.2f: mean decoded length %.1f bases per 4000-sample chunk' % (gain, lensn.mean()))
    h.close()


def test_encoder_decode_determinism_stress(full_batch):
    """100 repetitions of encoder + decode at N = 512, T = 800: bit-identical scores and packed rows every time."""
    x = full_batch.cuda()
    h, _, _ = _mixed_decode_handle(x)
    s0 = h.encoder(x)
    seq0, _, lens0 = h.decode(s0, want_qstring=False)
    bad = 0
    for rep in range(100):
        s = h.encoder(x)
        seq, _, lens = h.decode(s, want_qstring=False)
        if not (torch.equal(s, s0) and torch.equal(seq, seq0) and torch.equal(lens, lens0)):
            bad += 1
    assert bad == 0, '%d of 100 repetitions differed from the first run' % bad
    h.close()
