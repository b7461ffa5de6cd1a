"""GPU parity tests of the CUDA CRF decode against the oracle (through the C ABI).

Bar: labels / packed strings bit-exact against oracle/c/crf_exact.c and against the golden vectors
produced by the reference's own CTC_CRF code; posteriors / logZ / alpha / beta within fp32 tolerances
(the reference's torch.logsumexp / softmax association order is not reproducible bit for bit)."""
import numpy as np
import pytest
import torch

from make_golden import ALPHABETS, synthetic_scores, synthetic_targets
from oracle import bonito_oracle as bo
from oracle import cexact

pytestmark = pytest.mark.gpu

# max abs error on posteriors (values in [0,1]) against the fp32 torch restatement / the reference's golden values.  The
# reference works in fp32 log space (alphas ~1e3, one ulp 6e-5) and is itself 2e-5..2e-4 away from a float64 evaluation;
# the linear-domain kernels are within POST_TOL_F64 of float64, so this bound measures the reference's rounding.
POST_TOL = 5e-5
POST_TOL_F64 = 1e-6
LOGZ_RTOL = 2e-6     # relative error on logZ (|logZ| ~ 1e3 at T=160: fp32 ulp 6e-5)


@pytest.fixture(scope='module')
def handles():
    from xna_basecaller_b200._lib import Handle
    hs = {n: Handle(ALPHABETS[n], 3, max_N=64, max_T=800, encoder=False) for n in (4, 5, 6)}
    yield hs
    for h in hs.values():
        h.close()


@pytest.mark.parametrize('n_base', [4, 5, 6])
@pytest.mark.parametrize('seed', [0, 1])
def test_decode_matches_reference_golden(handles, golden, n_base, seed):
    h = handles[n_base]
    g = golden['crf']
    key = 'n%d_s%d_' % (n_base, seed)
    s = synthetic_scores(seed, 160, 3, n_base)
    seq, qs, lens, labels, post = h.decode(s, want_labels=True, want_post=True)
    torch.cuda.synchronize()
    labels = labels.cpu().numpy()
    assert np.array_equal(labels.T, g[key + 'paths'])                       # bit-exact Viterbi paths
    strings = [bytes(seq[i, :lens[i]].cpu().numpy().astype('u1')).decode() for i in range(3)]
    assert strings == list(g[key + 'strings'])                              # decode_batch strings
    assert np.array_equal((qs.cpu().numpy() != 0), (seq.cpu().numpy() != 0))
    assert set(np.unique(qs.cpu().numpy())) <= {0, ord('O')}
    post = post.cpu().numpy()
    assert np.abs(post[::9, :, ::7] - g[key + 'post_sub']).max() < POST_TOL
    assert np.abs(post.sum(2) - 1).max() < 1e-5
    lz = h.logZ(s).cpu().numpy()
    np.testing.assert_allclose(lz, g[key + 'logZ'], rtol=LOGZ_RTOL)
    np.testing.assert_allclose(h.forward_scores(s)[-1].cpu().numpy(), g[key + 'alpha_last'], rtol=LOGZ_RTOL, atol=1e-4)
    np.testing.assert_allclose(h.backward_scores(s)[0].cpu().numpy(), g[key + 'beta_first'], rtol=LOGZ_RTOL, atol=1e-4)
    raw = h.viterbi(s).cpu().numpy()
    assert np.array_equal(raw.T, g[key + 'paths_raw'])                      # CTC_CRF.viterbi on raw scores


@pytest.mark.parametrize('n_base,T,N', [(5, 800, 64), (6, 800, 32), (4, 333, 17), (5, 1, 3), (5, 2, 1)])
def test_decode_bit_exact_vs_c_oracle(handles, n_base, T, N):
    h = handles[n_base]
    s = synthetic_scores(1000 + T + N, T, N, n_base)
    seq, qs, lens, labels, post = h.decode(s, want_labels=True, want_post=True)
    torch.cuda.synchronize()
    o_labels, o_post = cexact.crf_decode(s.numpy(), n_base, want_post=True)
    assert np.array_equal(labels.cpu().numpy(), o_labels)
    # posteriors are produced by the same arithmetic contract: bit-equal, not just close
    assert np.array_equal(post.cpu().numpy().view(np.uint32), o_post.view(np.uint32))
    o_seq, o_qs, o_lens = cexact.pack(o_labels, ALPHABETS[n_base])
    assert np.array_equal(seq.cpu().numpy(), o_seq)
    assert np.array_equal(qs.cpu().numpy(), o_qs)
    assert np.array_equal(lens.cpu().numpy(), o_lens)
    assert np.array_equal(h.logZ(s).cpu().numpy().view(np.uint32), cexact.crf_logz(s.numpy(), n_base).view(np.uint32))
    assert np.array_equal(h.viterbi(s).cpu().numpy(), cexact.crf_viterbi(s.numpy(), n_base))
    assert np.array_equal(h.forward_scores(s).cpu().numpy().view(np.uint32),
                          cexact.crf_alpha(s.numpy(), n_base).view(np.uint32))


def test_decode_saturated_ties(handles):
    """5*tanh saturates: many exactly equal scores.  First-index tie-breaking must match."""
    n_base = 5
    h = handles[n_base]
    rs = np.random.RandomState(3)
    s = rs.choice(np.array([-5.0, 5.0, 2.0, 0.0], dtype=np.float32), size=(200, 8, 125, 6))
    s[..., 0] = 2.0
    s = torch.from_numpy(s.reshape(200, 8, -1))
    labels = h.decode(s, want_labels=True)[3].cpu().numpy()
    assert np.array_equal(labels, cexact.crf_decode(s.numpy(), n_base))
    assert np.array_equal(h.viterbi(s).cpu().numpy(), cexact.crf_viterbi(s.numpy(), n_base))


def test_decode_all_blank_is_empty(handles):
    """Random-init models give every non-blank score < blank: decode must be the empty string."""
    h = handles[5]
    s = torch.full((100, 4, 125, 6), -1.0)
    s[..., 0] = 2.0
    seq, qs, lens = h.decode(s.reshape(100, 4, -1))
    assert lens.cpu().tolist() == [0, 0, 0, 0]
    assert int(seq.abs().sum()) == 0


@pytest.mark.parametrize('n_base', [5, 6])
def test_posteriors_vs_torch_restatement(handles, n_base):
    h = handles[n_base]
    s = synthetic_scores(77, 120, 4, n_base)
    ref = bo.CRF(3, ALPHABETS[n_base]).posteriors(s)
    got = h.posteriors(s).cpu()
    assert (got - ref).abs().max().item() < POST_TOL
    ref64 = bo.CRF(3, ALPHABETS[n_base]).posteriors(s.double())
    assert (got.double() - ref64).abs().max().item() < POST_TOL_F64


@pytest.mark.parametrize('n_base,T,N', [(5, 300, 9), (6, 200, 5), (4, 64, 3)])
def test_exp_hand_over_format(handles, n_base, T, N):
    """xb_crf_decode_exp on exp(scores) (the fused route's hand-over format) == xb_crf_decode on the scores, bit for bit,
    and both == the C oracle; the exponential is the contract's xb_score_exp (clamp to [-80, 80])."""
    h = handles[n_base]
    s = synthetic_scores(500 + T, T, N, n_base)
    s[0, 0, :7] = torch.tensor([-200.0, 200.0, -80.0, 80.0, -81.0, 0.0, 1e-30])       # clamp edges
    e = torch.from_numpy(cexact.score_exp(s.numpy()))
    a = h.decode(s, want_labels=True, want_post=True)
    b = h.decode(e, want_labels=True, want_post=True, exp_input=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    o_labels, o_post = cexact.crf_decode(s.numpy(), n_base, want_post=True)
    assert np.array_equal(a[3].cpu().numpy(), o_labels)
    assert np.array_equal(a[4].cpu().numpy().view(np.uint32), o_post.view(np.uint32))


@pytest.mark.parametrize('n_base', [4, 5, 6])
def test_max_semiring_methods(handles, n_base):
    """CTC_CRF.logZ / forward_scores / backward_scores / posteriors with S = Max (crf/model.py:41-61, 92-95) through the
    plugin class against the torch restatement: max-plus arithmetic is exact, so values agree to rounding of the same
    additions and the one-hot posteriors agree exactly."""
    from xna_basecaller_b200.crf.model import CTC_CRF, Max
    s = synthetic_scores(31 + n_base, 150, 4, n_base)
    ref = bo.CRF(3, ALPHABETS[n_base])
    sd = CTC_CRF(3, ALPHABETS[n_base])
    sc = s.cuda()
    np.testing.assert_allclose(sd.logZ(sc, Max).cpu().numpy(), ref.logZ(s, 'max').numpy(), rtol=1e-6)
    np.testing.assert_allclose(sd.forward_scores(sc, Max).cpu().numpy(), ref.forward_scores(s, 'max').numpy(), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(sd.backward_scores(sc, Max).cpu().numpy(), ref.backward_scores(s, 'max').numpy(), rtol=1e-6, atol=1e-4)
    post = sd.posteriors(sc, Max).cpu()
    want = ref.posteriors(s, 'max')
    assert post.sum(2).eq(1).all() and (post.eq(0) | post.eq(1)).all()
    assert (post.argmax(2) == want.argmax(2)).float().mean().item() > 0.999        # ties in fp32 sums aside: identical
    assert torch.equal(post.argmax(2) % (n_base + 1), sd.viterbi(sc).cpu())
    with pytest.raises(NotImplementedError):
        sd.logZ(sc, 'tropical')
    sd.engine.close()


@pytest.mark.parametrize('n_base', [4, 5, 6])
def test_ctc_loss_matches_reference_golden(handles, golden, n_base):
    h = handles[n_base]
    g = golden['crf']
    for seed in (0, 1):
        s = synthetic_scores(seed, 160, 3, n_base)
        tg, tl = synthetic_targets(100 + seed, 3, n_base, 30, 50)
        loss = h.ctc_loss(s, tg, tl).cpu().numpy()
        np.testing.assert_allclose(loss, g['n%d_s%d_ctc_loss' % (n_base, seed)], rtol=2e-5)


def test_ctc_loss_long_targets(handles):
    n_base = 5
    h = handles[n_base]
    s = synthetic_scores(5, 800, 6, n_base)
    tg, tl = synthetic_targets(9, 6, n_base, 350, 450)
    ref = bo.CRF(3, ALPHABETS[n_base]).ctc_loss(s, tg, tl, reduction='none')
    np.testing.assert_allclose(h.ctc_loss(s, tg, tl).cpu().numpy(), ref.numpy(), rtol=5e-5)


@pytest.mark.parametrize('n_base', [4, 5, 6])
def test_ctc_loss_backward_matches_reference_golden(handles, golden, n_base):
    """xb_ctc_crf_loss_bwd with grad_loss = 1/N (reduction='mean') against the gradient autograd gives the reference class."""
    h = handles[n_base]
    g = golden['crf_grad']
    for seed in (0, 1):
        s = synthetic_scores(seed, 160, 3, n_base)
        tg, tl = synthetic_targets(100 + seed, 3, n_base, 30, 50)
        grad = h.ctc_loss_bwd(s, tg, tl, torch.full((3,), 1.0 / 3)).cpu()
        key = 'n%d_s%d_' % (n_base, seed)
        np.testing.assert_allclose(grad[::9, :, ::7].numpy(), g[key + 'grad_sub'], rtol=5e-4, atol=5e-7)
        np.testing.assert_allclose(grad.abs().sum(2).numpy(), g[key + 'grad_abs_rowsum'], rtol=1e-3)


@pytest.mark.parametrize('normalise', [True, False])
def test_ctc_loss_backward_weighted_long_targets(handles, normalise):
    """Per-sequence upstream gradients (incl. 0 = a clipped loss, and a negative one), T = 800, targets of 350-450 bases
    with repeated k-mers (several positions scatter into one edge), against autograd through the oracle."""
    n_base = 5
    h = handles[n_base]
    s = synthetic_scores(5, 800, 4, n_base)
    tg, tl = synthetic_targets(9, 4, n_base, 350, 450)
    w = torch.tensor([1.0, 0.0, -0.5, 0.25])
    sr = s.clone().requires_grad_()
    loss = bo.CRF(3, ALPHABETS[n_base]).ctc_loss(sr, tg, tl, reduction='none', normalise_scores=normalise)
    (loss * w).sum().backward()
    grad = h.ctc_loss_bwd(s, tg, tl, w, normalise=normalise).cpu()
    # alpha + beta - logz is a difference of numbers ~1e3 in fp32: ~1e-3 relative on posteriors <= |w| / len ~ 3e-3
    assert (grad - sr.grad).abs().max().item() < 2e-5
    assert grad[:, 1].abs().max().item() == 0.0


@pytest.mark.parametrize('reverse', [False, True])
def test_stitch_matches_reference_golden(handles, golden, reverse):
    """xb_stitch against the reference's util.stitch output, forward and reverse=True (util.py:180-184)."""
    h = handles[5]
    g = golden['stitch']
    for cs, ov in ((4000, 500), (3600, 500), (1000, 100)):
        T = cs // 5
        lens, firsts, counts, rows, expect = [], [], [], [], []
        for L in (cs - 1, cs, cs + 1, 7500, 10000, 10001, 2 * cs - ov, 19999):
            key = 'c%d_o%d_L%d_' % (cs, ov, L)
            nch = len(g[key + 'first'])
            # rows hold 1 + (flat index % 120) so that every position is a distinct-ish non-zero byte
            lab = (np.arange(nch * T).reshape(nch, T) % 120 + 1).astype(np.int8)
            firsts.append(sum(counts))
            counts.append(nch)
            lens.append(L)
            rows.append(lab)
            expect.append((g[key + ('stitched_rev' if reverse else 'stitched')] % 120 + 1).astype(np.int8))
        out, out_len = h.stitch(torch.from_numpy(np.concatenate(rows)), firsts, counts, lens, cs, ov, reverse=reverse)
        out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
        for i, e in enumerate(expect):
            assert out_len[i] == len(e)
            assert np.array_equal(out[i, :len(e)], e)


def test_full_size_properties():
    """BASELINE configs 2 and 3 sizes: properties that do not need the oracle at full size, plus an
    oracle spot check on a few sequences cut out of the big batch (chunks are independent)."""
    from xna_basecaller_b200._lib import Handle
    for n_base, N in ((5, 512), (6, 1024)):
        h = Handle(ALPHABETS[n_base], 3, max_N=N, max_T=800, encoder=False)
        C, NZ = n_base ** 3, n_base + 1
        g = torch.Generator(device='cuda').manual_seed(7)
        s = torch.empty(800, N, C, NZ, device='cuda').uniform_(-5, 5, generator=g)
        s[..., 0] = 2.0
        s = s.reshape(800, N, -1)
        seq, qs, lens, labels, post = h.decode(s, want_labels=True, want_post=True)
        torch.cuda.synchronize()
        assert (post.sum(2) - 1).abs().max().item() < 1e-5
        assert post.min().item() >= 0
        assert int(labels.min()) >= 0 and int(labels.max()) <= n_base
        assert torch.equal((labels != 0).sum(1).int(), lens)
        # idempotence / determinism: a second run gives the same bits
        seq2, _, lens2 = h.decode(s, want_qstring=False)
        assert torch.equal(seq, seq2) and torch.equal(lens, lens2)
        # batch independence + oracle: first, middle and last sequences
        pick = [0, N // 2 + 1, N - 1]
        sub = s[:, pick].contiguous().cpu().numpy()
        assert np.array_equal(labels[pick].cpu().numpy(), cexact.crf_decode(sub, n_base))
        del s, post
        h.close()
        torch.cuda.empty_cache()


@pytest.mark.parametrize('n_base,T,N,width', [(5, 400, 6, 32), (6, 300, 4, 32), (4, 500, 5, 16), (5, 50, 3, 1)])
def test_beam_search_bit_exact_vs_c_oracle(handles, n_base, T, N, width):
    """xb_crf_beam_search == the plain-C statement of the algorithm (labels via moves / letters, qualities), and its
    sequences are at least as long-lived as Viterbi's on the same scores (sanity: similar lengths)."""
    h = handles[n_base]
    s = synthetic_scores(900 + T, T, N, n_base)
    seq, qs, moves, lens = h.beam_search(s, beam_width=width, beam_cut=100.0)
    torch.cuda.synchronize()
    o_labels, o_quals = cexact.crf_beam_search(s.numpy(), n_base, beam_width=width, beam_cut=100.0)
    o_seq, _, o_lens = cexact.pack(o_labels, ALPHABETS[n_base])
    assert np.array_equal(moves.cpu().numpy(), o_labels != 0)
    assert np.array_equal(seq.cpu().numpy(), o_seq) and np.array_equal(lens.cpu().numpy(), o_lens)
    for i in range(N):
        assert np.array_equal(qs[i, :o_lens[i]].cpu().numpy().astype(np.uint8), o_quals[i][o_quals[i] > 0])
        assert not qs[i, o_lens[i]:].any()
    vit_lens = h.decode(s, want_qstring=False)[2].cpu().numpy()
    assert np.abs(lens.cpu().numpy() - vit_lens).max() <= max(8, T // 20)
