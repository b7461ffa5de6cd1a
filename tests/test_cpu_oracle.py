"""CPU tests of the oracle itself (no GPU): the restatements under oracle/ against the golden vectors that the
reference's own code produced (tests/golden/make_golden.py), against brute force, and against each other.
An oracle that drifts from the reference would make every GPU parity test meaningless, so it is pinned here."""
import itertools

import numpy as np
import pytest
import torch

from make_golden import ALPHABETS, REF_SCALE, synthetic_scores, synthetic_signal, synthetic_targets
from oracle import bonito_oracle as bo
from oracle import cexact
from oracle import seqdist_restated as sr


# ----------------------------------------------------------------------------- exact math
def test_exact_expf_logf_within_2ulp():
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(-80, 80, 20000), rs.uniform(-1, 1, 20000), [-86.0, -87.0, 0.0, 88.0]]).astype(np.float32)
    got = cexact.expf(x).astype(np.float64)
    want = np.exp(x.astype(np.float64))
    ok = want > 1e-37
    ulp = np.spacing(want[ok].astype(np.float32)).astype(np.float64)
    assert (np.abs(got[ok] - want[ok]) <= 2 * ulp).all()
    assert cexact.expf(np.float32([-1e38, -np.inf, -87.0])).tolist() == [0.0, 0.0, 0.0]
    y = np.concatenate([rs.uniform(1e-30, 1e30, 10), np.exp(rs.uniform(-60, 60, 40000))]).astype(np.float32)
    got = cexact.logf(y).astype(np.float64)
    want = np.log(y.astype(np.float64))
    ulp = np.maximum(np.spacing(np.abs(want).astype(np.float32)).astype(np.float64), 2 ** -24)
    assert (np.abs(got - want) <= 2 * ulp).all()
    assert cexact.logf(np.float32([1.0]))[0] == 0.0


def test_exact_math_fast_variants():
    """The branch-trimmed exp / log the CUDA sweeps use return the same bits as the general ones on their domain."""
    rs = np.random.RandomState(1)
    x = np.concatenate([-np.abs(rs.uniform(0, 100, 200000)), -rs.exponential(3.0, 200000), [0.0, -0.0, -86.0, -86.5, -1e38, -np.inf]]).astype(np.float32)
    assert np.array_equal(cexact.expf_le0(x).view(np.uint32), cexact.expf(x).view(np.uint32))
    y = np.concatenate([np.exp(rs.uniform(-18.5, 10, 400000)), [1e-8, 1.0, 750.0, 1.00000001e-8]]).astype(np.float32)
    assert np.array_equal(cexact.logf_norm(y).view(np.uint32), cexact.logf(y).view(np.uint32))


# The reference computes the posteriors in fp32 LOG space (seqdist's logsumexp scans): alphas and betas grow to ~1e3 over
# 800 steps, where one fp32 ulp is 6e-5, so that formulation carries ~2e-5 .. 2e-4 of absolute error in a posterior.  The
# linear-domain contract of the CUDA decode is accurate to ~2e-7 against a float64 evaluation; its distance to the fp32
# reference values is therefore the REFERENCE's rounding, not the kernel's.
POST_TOL_VS_FP32_REFERENCE = 5e-5
POST_TOL_VS_FP64 = 1e-6


@pytest.mark.parametrize('n_base,T', [(5, 800), (6, 400), (4, 800)])
def test_linear_domain_decode_is_closer_to_float64_than_the_log_domain_one(n_base, T):
    """Accuracy of the two formulations against the float64 restatement at BASELINE length, and label agreement."""
    s = synthetic_scores(3, T, 2, n_base)
    crf = bo.CRF(3, ALPHABETS[n_base])
    p64 = crf.posteriors(s.double())
    labels, post = cexact.crf_decode(s.numpy(), n_base, want_post=True)
    _, post_ld = cexact.crf_decode_logdomain(s.numpy(), n_base, want_post=True)
    err_lin, err_log = np.abs(post - p64.numpy()).max(), np.abs(post_ld - p64.numpy()).max()
    assert err_lin < POST_TOL_VS_FP64 and err_lin < err_log
    want = crf.viterbi((p64 + 1e-8).log()).numpy().T           # decode_batch in float64
    assert (labels == want).mean() == 1.0
    # the hand-over format of the fused route: decoding exp(scores) gives the same bits
    labels_e, post_e = cexact.crf_decode(cexact.score_exp(s.numpy()), n_base, want_post=True, exp_input=True)
    assert np.array_equal(labels_e, labels) and np.array_equal(post_e.view(np.uint32), post.view(np.uint32))


# ----------------------------------------------------------------------------- CRF vs golden / brute force
@pytest.mark.parametrize('n_base', [4, 5, 6])
@pytest.mark.parametrize('seed', [0, 1])
def test_crf_restatement_matches_reference_golden(golden, n_base, seed):
    g = golden['crf']
    key = 'n%d_s%d_' % (n_base, seed)
    crf = bo.CRF(3, ALPHABETS[n_base])
    s = synthetic_scores(seed, 160, 3, n_base)
    np.testing.assert_allclose(crf.logZ(s).numpy(), g[key + 'logZ'], rtol=2e-6)
    post = crf.posteriors(s)
    assert np.abs(post[::9, :, ::7].numpy() - g[key + 'post_sub']).max() < 2e-6
    assert np.abs(post.sum(2).numpy() - 1).max() < 1e-5
    assert crf.decode_batch(s) == list(g[key + 'strings'])
    assert np.array_equal(crf.viterbi(s).numpy().astype(np.int8), g[key + 'paths_raw'])
    tg, tl = synthetic_targets(100 + seed, 3, n_base, 30, 50)
    np.testing.assert_allclose(crf.ctc_loss(s, tg, tl, reduction='none').numpy(), g[key + 'ctc_loss'], rtol=1e-5)
    np.testing.assert_allclose(crf.forward_scores(s)[-1].numpy(), g[key + 'alpha_last'], rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(crf.backward_scores(s)[0].numpy(), g[key + 'beta_first'], rtol=2e-6, atol=1e-4)


@pytest.mark.parametrize('n_base', [4, 5, 6])
def test_oracle_ctc_loss_gradient_matches_reference_golden(golden, n_base):
    """d ctc_loss / d scores of the restated path == the reference class's autograd result, and equals the closed form
    (P_full - P_target) / (N * len) the CUDA backward kernel implements (rows sum to 0: both posteriors sum to 1)."""
    g = golden['crf_grad']
    crf = bo.CRF(3, ALPHABETS[n_base])
    for seed in (0, 1):
        s = synthetic_scores(seed, 160, 3, n_base).requires_grad_()
        tg, tl = synthetic_targets(100 + seed, 3, n_base, 30, 50)
        crf.ctc_loss(s, tg, tl).backward()
        key = 'n%d_s%d_' % (n_base, seed)
        np.testing.assert_allclose(s.grad[::9, :, ::7].numpy(), g[key + 'grad_sub'], rtol=5e-4, atol=5e-7)
        np.testing.assert_allclose(s.grad.abs().sum(2).numpy(), g[key + 'grad_abs_rowsum'], rtol=1e-4)
        assert s.grad.sum(2).abs().max().item() < 1e-6


@pytest.mark.parametrize('n_base', [4, 5, 6])
def test_c_checker_matches_reference_golden(golden, n_base):
    """The bit-exact C checker decodes the golden inputs to exactly the reference's paths and strings."""
    g = golden['crf']
    for seed in (0, 1):
        key = 'n%d_s%d_' % (n_base, seed)
        s = synthetic_scores(seed, 160, 3, n_base).numpy()
        labels, post = cexact.crf_decode(s, n_base, want_post=True)
        assert np.array_equal(labels.T, g[key + 'paths'])
        assert cexact.strings(labels, ALPHABETS[n_base]) == list(g[key + 'strings'])
        assert np.abs(post[::9, :, ::7] - g[key + 'post_sub']).max() < POST_TOL_VS_FP32_REFERENCE
        ld_labels, ld_post = cexact.crf_decode_logdomain(s, n_base, want_post=True)     # second, independent formulation
        assert np.array_equal(ld_labels.T, g[key + 'paths'])
        assert np.abs(ld_post - post).max() < POST_TOL_VS_FP32_REFERENCE
        np.testing.assert_allclose(cexact.crf_logz(s, n_base), g[key + 'logZ'], rtol=2e-6)
        assert np.array_equal(cexact.crf_viterbi(s, n_base).T, g[key + 'paths_raw'])


def test_crf_brute_force_tiny():
    """n_base 2, state_len 2, T 5: enumerate every state path; logZ and the best path must agree."""
    n, sl, T = 2, 2, 5
    crf = bo.CRF(sl, ['N', 'A', 'C'])
    rs = np.random.RandomState(3)
    s = torch.from_numpy(rs.uniform(-3, 3, size=(T, 1, crf.C * crf.NZ)).astype(np.float32))
    Ms = s.reshape(T, crf.C, crf.NZ).numpy().astype(np.float64)
    idx = crf.idx.numpy()
    total, best = [], (-np.inf, None)
    for states in itertools.product(range(crf.C), repeat=T + 1):
        for edges in itertools.product(range(crf.NZ), repeat=T):
            if all(idx[states[t + 1], edges[t]] == states[t] for t in range(T)):
                w = sum(Ms[t, states[t + 1], edges[t]] for t in range(T))
                total.append(w)
                if w > best[0]:
                    best = (w, edges)
    logz = np.logaddexp.reduce(total)
    assert abs(crf.logZ(s).item() - logz) < 1e-4
    assert abs(crf.logZ(s, 'max').item() - best[0]) < 1e-4
    assert crf.viterbi(s)[:, 0].tolist() == list(best[1])
    assert cexact.crf_viterbi(s.numpy(), n, sl)[0].tolist() == list(best[1])
    # autograd formulation (seqdist style) agrees with the explicit alpha/beta one
    post_autograd = sr.posteriors(s.reshape(T, 1, crf.C, crf.NZ), crf.idx.to(torch.int64)) if hasattr(sr, 'posteriors') else None
    if post_autograd is not None:
        assert np.abs(post_autograd.reshape(T, 1, -1).numpy() - crf.posteriors(s).numpy()).max() < 1e-5


@pytest.mark.parametrize('n_base,T,N', [(5, 64, 4), (6, 40, 3), (4, 1, 2), (5, 2, 1)])
def test_c_checker_vs_torch_restatement(n_base, T, N):
    s = synthetic_scores(40 + T, T, N, n_base)
    crf = bo.CRF(3, ALPHABETS[n_base])
    labels, post = cexact.crf_decode(s.numpy(), n_base, want_post=True)
    want_post = crf.posteriors(s)
    assert np.abs(post - want_post.numpy()).max() < POST_TOL_VS_FP32_REFERENCE
    assert np.abs(post - crf.posteriors(s.double()).numpy()).max() < POST_TOL_VS_FP64
    want = crf.viterbi((want_post + 1e-8).log()).numpy().T
    # near-ties aside the two agree; on continuous random scores they agree everywhere
    assert (labels == want).mean() > 0.999
    alpha = cexact.crf_alpha(s.numpy(), n_base)
    np.testing.assert_allclose(alpha, crf.forward_scores(s).numpy(), rtol=2e-6, atol=1e-4)


def test_c_checker_ties_take_first_index():
    """Saturated scores (exact ties everywhere): arg-max must take the first flat edge index, as torch.argmax."""
    n_base, T, N = 5, 12, 2
    C, NZ = 125, 6
    s = np.full((T, N, C, NZ), 5.0, dtype=np.float32)
    s[..., 0] = 2.0
    labels = cexact.crf_decode(s.reshape(T, N, -1), n_base)
    crf = bo.CRF(3, ALPHABETS[n_base])
    st = torch.from_numpy(s.reshape(T, N, -1))
    want = crf.viterbi((crf.posteriors(st) + 1e-8).log()).numpy().T
    assert labels.shape == (N, T)
    assert set(np.unique(labels)) <= set(range(NZ))
    assert np.array_equal(labels, want)


def test_all_blank_decodes_to_empty():
    """Random-init degeneracy: every non-blank score below the blank score -> empty strings."""
    s = synthetic_scores(5, 50, 2, 5).numpy().reshape(50, 2, 125, 6).copy()
    s[..., 1:] = np.minimum(s[..., 1:], 1.0) - 4.0
    labels = cexact.crf_decode(s.reshape(50, 2, -1), 5)
    assert (labels == 0).all()
    assert cexact.strings(labels, ALPHABETS[5]) == ['', '']


# ----------------------------------------------------------------------------- encoder vs golden
@pytest.mark.parametrize('n_base', [5, 6])
def test_encoder_restatement_matches_reference_golden(golden, n_base):
    g = golden['encoder']
    sd = bo.reference_state_dict(n_base=n_base, seed=11)
    x = synthetic_signal(21, 2, 500)
    key = 'n%d_' % n_base
    stem = bo.conv_stem(sd, x)
    assert np.abs(stem[:, ::16, :].numpy() - g[key + 'stem_sub']).max() < 1e-5
    p = 'encoder.4.rnn.'
    l1 = bo.lstm_layer(stem.permute(2, 0, 1).contiguous(), sd[p + 'weight_ih_l0'], sd[p + 'weight_hh_l0'],
                       sd[p + 'bias_ih_l0'], sd[p + 'bias_hh_l0'], True)
    assert np.abs(l1[:, :, ::16].numpy() - g[key + 'lstm1_sub']).max() < 1e-4
    scores = bo.encoder_forward(sd, x, n_base)
    assert np.abs(scores.numpy() - g[key + 'scores']).max() < 2e-3
    lib = bo.encoder_forward(sd, x, n_base, library=True)
    assert np.abs(lib.numpy() - g[key + 'scores']).max() < 2e-3
    crf = bo.CRF(3, ALPHABETS[n_base])
    strings = crf.decode_batch(torch.from_numpy(g[key + 'scores']))
    assert strings == list(g[key + 'strings'])
    packed = bo.left_pack(strings, scores.shape[0])
    assert np.array_equal(packed['sequence'], g[key + 'cs_sequence'])
    assert np.array_equal(packed['qstring'], g[key + 'cs_qstring'])
    assert np.array_equal(packed['moves'], g[key + 'cs_moves'].astype(bool))


def test_encoder_reference_scale(golden):
    sd = bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE)
    x = synthetic_signal(23, 2, 500)
    assert np.abs(bo.encoder_forward(sd, x, 5).numpy() - golden['encoder']['r5_scores']).max() < 1e-3


# ----------------------------------------------------------------------------- chunk / stitch / basecall
def _stitch_cases():
    for cs_, ov in ((4000, 500), (3600, 500), (1000, 100)):
        for L in (cs_ - 1, cs_, cs_ + 1, 7500, 10000, 10001, 2 * cs_ - ov, 19999):
            yield cs_, ov, L


@pytest.mark.parametrize('cs_,ov,L', list(_stitch_cases()))
def test_chunk_stitch_restatement_matches_reference_golden(golden, cs_, ov, L):
    g = golden['stitch']
    key = 'c%d_o%d_L%d_' % (cs_, ov, L)
    sig = torch.arange(L, dtype=torch.float32)
    ch = bo.chunk(sig, cs_, ov)
    assert np.array_equal(ch[:, 0, 0].numpy().astype(np.int64), g[key + 'first'])
    assert np.array_equal(ch[-1, 0, -3:].numpy().astype(np.int64), g[key + 'lastrow'])
    T = cs_ // 5
    lab = torch.arange(ch.shape[0] * T, dtype=torch.int32).reshape(ch.shape[0], T)
    assert np.array_equal(bo.stitch(lab, cs_, ov, L, 5).numpy(), g[key + 'stitched'])
    assert np.array_equal(bo.stitch(lab, cs_, ov, L, 5, reverse=True).numpy(), g[key + 'stitched_rev'])


def test_basecall_restatement_matches_reference_golden(golden):
    g = golden['basecall']
    sd = bo.reference_state_dict(n_base=5, seed=11)
    crf = bo.CRF(3, ALPHABETS[5])
    rs = np.random.RandomState(77)
    reads = [('read%d' % i, rs.randn(int(L)).astype(np.float32)) for i, L in enumerate(g['lengths'])]

    def score_fn(batch):
        with torch.no_grad():
            return bo.encoder_forward(sd, batch, 5)

    out = dict(bo.basecall(score_fn, crf, reads, 1000, 100, 4))
    assert list(out) == ['read%d' % i for i in range(len(reads))]
    for rid, res in out.items():
        assert res['sequence'] == str(g[rid + '_sequence'])
        assert res['qstring'] == str(g[rid + '_qstring'])
        assert len(res['sig_move']) == int(g[rid + '_sig_move_len'])
        assert not res['sig_move'].any()


# ----------------------------------------------------------------------------- beam search
@pytest.mark.parametrize('T', [1, 2, 3, 5])
def test_beam_search_finds_the_most_probable_sequence_brute_force(T):
    """n_base 2, state_len 2: enumerate every path, group the paths by the base sequence they spell (start state, emitted
    labels, final state -- the last state_len bases are still inside the state), logsumexp per sequence.  The beam search
    (width 32 < number of sequences from T = 2 on, so it prunes) must return the labels of the most probable sequence."""
    n, sl = 2, 2
    crf = bo.CRF(sl, ['N', 'A', 'C'])
    rs = np.random.RandomState(4 + T)
    idx = crf.idx.numpy()
    for _ in range(12):
        s = rs.uniform(-3, 3, size=(T, 1, crf.C * crf.NZ)).astype(np.float32)
        Ms = s.reshape(T, crf.C, crf.NZ).astype(np.float64)
        groups = {}
        for states in itertools.product(range(crf.C), repeat=T + 1):
            for edges in itertools.product(range(crf.NZ), repeat=T):
                if all(idx[states[t + 1], edges[t]] == states[t] for t in range(T)):
                    w = sum(Ms[t, states[t + 1], edges[t]] for t in range(T))
                    groups.setdefault((states[0], states[-1]) + tuple(e for e in edges if e), []).append(w)
        best = max(groups.items(), key=lambda kv: np.logaddexp.reduce(kv[1]))[0]
        labels, quals = cexact.crf_beam_search(s, n, sl, beam_width=32)
        assert tuple(int(x) for x in labels[0] if x) == best[2:]
        assert ((quals[0] > 0) == (labels[0] != 0)).all() and (quals[0][quals[0] > 0] >= 34).all() and quals[0].max() <= 83


def test_beam_search_on_confident_scores_equals_viterbi():
    """Where one path dominates (confident scores) the most probable sequence is the Viterbi path's sequence."""
    rs = np.random.RandomState(1)
    T, n_base = 120, 5
    C, NZ = 125, 6
    s = np.full((T, 1, C, NZ), -5.0, dtype=np.float32)
    s[..., 0] = 2.0
    state = 17
    for t in range(T):                       # plant a path: stay or move with a large margin
        if rs.rand() < 0.5:
            new = (state % 25) * 5 + rs.randint(5)
            s[t, 0, new, 1 + state // 25] = 5.0
            state = new
        else:
            s[t, 0, state, 0] = 5.0
    s = s.reshape(T, 1, -1)
    labels, _ = cexact.crf_beam_search(s, n_base)
    assert np.array_equal(labels, cexact.crf_viterbi(s, n_base))
    assert np.array_equal(labels, cexact.crf_decode(s, n_base))
