"""GPU tests of the drop-in plugin surface: the reference's own call sequence (load_symbol -> Model(config) ->
load_state_dict -> .half().eval().to(device) -> basecall(model, reads, ...)) running on the CUDA path, checked
against golden outputs of the reference code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from make_golden import ALPHABETS, synthetic_scores, synthetic_signal, synthetic_targets
from oracle import bonito_oracle as bo
from test_cpu_host import sup_config

pytestmark = pytest.mark.gpu

SCORE_TOL_F16_AMPLIFIED = 6e-2   # same bound as tests/test_gpu_encoder.py for the gain-12 head of the golden weights


@pytest.fixture(scope='module')
def model5():
    from xna_basecaller_b200 import util
    cfg = sup_config(ALPHABETS[5])
    Model = util.load_symbol(cfg, 'Model')
    m = Model(cfg)
    m.load_state_dict(bo.reference_state_dict(n_base=5, seed=11))
    m = m.half().eval().to('cuda')
    return m


def test_model_forward_and_decode_batch(model5, golden):
    x = synthetic_signal(21, 2, 500)
    with torch.no_grad():
        scores = model5(x.half().cuda())
    gold = torch.from_numpy(golden['encoder']['n5_scores'])
    assert scores.shape == gold.shape and scores.dtype == torch.float32
    assert (scores.cpu() - gold).abs().max().item() <= SCORE_TOL_F16_AMPLIFIED
    assert model5.decode_batch(scores) == list(golden['encoder']['n5_strings'])
    assert model5.decode(scores[:, 0]) == str(golden['encoder']['n5_strings'][0])
    assert model5.seqdist.engine.handle.launches > 0


def test_layers_run_one_at_a_time(model5, golden):
    """nn.LSTM / nn.LinearCRFEncoder called as stand-alone modules give the same result as the fused Serial."""
    x = synthetic_signal(21, 2, 500).half().cuda()
    enc = model5.encoder
    with torch.no_grad():
        y = enc[:4](x)                                   # fused stem -> (T, N, 768)
        assert y.shape == (100, 2, 768)
        stem_gold = torch.from_numpy(golden['encoder']['n5_stem_sub'])          # (N, 768/16, T)
        assert (y.float().cpu().permute(1, 2, 0)[:, ::16, :] - stem_gold).abs().max().item() < 2e-2
        for layer in enc[4:9]:
            y = layer(y)
        scores = enc[9](y)
        whole = model5(x)
    assert torch.equal(scores, whole)


def test_compute_scores_matches_reference_golden(model5, golden):
    from xna_basecaller_b200.crf.basecall import compute_scores
    x = synthetic_signal(21, 2, 500)
    out = compute_scores(model5, x)
    g = golden['encoder']
    assert out['sequence'].dtype == torch.int8 and out['qstring'].dtype == torch.int8 and out['moves'].dtype == bool
    assert np.array_equal(out['sequence'].numpy(), g['n5_cs_sequence'])
    assert np.array_equal(out['qstring'].numpy(), g['n5_cs_qstring'])
    assert np.array_equal(out['moves'], g['n5_cs_moves'].astype(bool))


def test_basecall_matches_reference_golden(model5, golden):
    from xna_basecaller_b200 import util
    g = golden['basecall']

    class Read:
        def __init__(self, rid, sig):
            self.read_id, self.signal = rid, sig

    rs = np.random.RandomState(77)
    reads = [Read('read%d' % i, rs.randn(int(L)).astype(np.float32)) for i, L in enumerate(g['lengths'])]
    basecall = util.load_symbol(model5.config, 'basecall')
    out = list(basecall(model5, iter(reads), chunksize=1000, overlap=100, batchsize=4))
    assert [r.read_id for r, _ in out] == [r.read_id for r in reads]
    # (1) bit-exact against the oracle's decode + stitch of the SAME (CUDA) scores
    crf = bo.CRF(3, ALPHABETS[5])

    def score_fn(batch):
        with torch.no_grad():
            return model5(batch.cuda()).cpu()

    want = dict(bo.basecall(score_fn, crf, [(r.read_id, r.signal) for r in reads], 1000, 100, 4))
    same, worst = 0, 0.0
    for rd, res in out:
        assert set(res) == {'sequence', 'qstring', 'sig_move'}
        assert res['sequence'] == want[rd.read_id]['sequence']
        assert res['qstring'] == 'O' * len(res['sequence'])
        assert len(res['sig_move']) == int(g[rd.read_id + '_sig_move_len']) and not res['sig_move'].any()
        # (2) against the reference's own fp32 run: identical-read rate, differing reads must be near-ties
        gold = str(g[rd.read_id + '_sequence'])
        same += res['sequence'] == gold
        worst = max(worst, _edit_distance(res['sequence'], gold) / max(len(gold), 1))
    print('identical-read rate vs the fp32 reference run: %d/%d, worst edit rate %.3f' % (same, len(reads), worst))
    assert same >= len(reads) // 2 and worst <= 0.03      # differing reads differ in a few near-tie steps only


def _edit_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[-1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def test_ctc_crf_methods(model5, golden):
    sd = model5.seqdist
    g = golden['crf']
    s = synthetic_scores(0, 160, 3, 5).cuda()
    np.testing.assert_allclose(sd.logZ(s).cpu().numpy(), g['n5_s0_logZ'], rtol=2e-6)
    post = sd.posteriors(s)
    assert np.abs(post[::9, :, ::7].cpu().numpy() - g['n5_s0_post_sub']).max() < 2e-5
    assert np.array_equal(sd.viterbi(s).cpu().numpy().astype(np.int8), g['n5_s0_paths_raw'])
    lp = (post + 1e-8).log()
    paths = sd.viterbi(lp).to(torch.int16).T.cpu().numpy()
    strings = [sd.path_to_str(p) for p in paths]
    assert sum(a == b for a, b in zip(strings, g['n5_s0_strings'])) >= 2     # torch.log vs exact log: near-ties only
    tg, tl = synthetic_targets(100, 3, 5, 30, 50)
    loss = sd.ctc_loss(s, tg.cuda(), tl.cuda(), reduction='none')
    np.testing.assert_allclose(loss.cpu().numpy(), g['n5_s0_ctc_loss'], rtol=1e-5)
    assert abs(sd.ctc_loss(s, tg.cuda(), tl.cuda()).item() - g['n5_s0_ctc_loss'].mean()) < 1e-4
    tp, ip = sd.compute_transition_probs(s, sd.backward_scores(s))
    ref = bo.CRF(3, ALPHABETS[5])
    tp_ref, ip_ref = ref.compute_transition_probs(s.cpu(), ref.backward_scores(s.cpu()))
    assert (tp.cpu() - tp_ref).abs().max().item() < 2e-4 and (ip.cpu() - ip_ref).abs().max().item() < 2e-4   # betas ~1e3: fp32 ulp 6e-5
    want = bo.CRF(3, ALPHABETS[5]).reverse_complement(s.cpu())
    assert torch.equal(sd.reverse_complement(s).cpu(), want)
    np.testing.assert_allclose((sd.normalise(s)).cpu().numpy(), bo.CRF(3, ALPHABETS[5]).normalise(s.cpu()).numpy(), atol=1e-4)


def test_ctc_loss_backward_through_plugin(model5, golden):
    """CTC_CRF.ctc_loss is differentiable w.r.t. the scores like the reference's (training.py:100-108 calls
    loss.backward()); loss_clip zeroes the gradient of clipped sequences through ordinary autograd."""
    from make_golden import synthetic_scores, synthetic_targets
    sd = model5.seqdist
    s = synthetic_scores(0, 160, 3, 5).cuda().requires_grad_()
    tg, tl = synthetic_targets(100, 3, 5, 30, 50)
    sd.ctc_loss(s, tg.cuda(), tl.cuda()).backward()
    g = golden['crf_grad']
    np.testing.assert_allclose(s.grad[::9, :, ::7].cpu().numpy(), g['n5_s0_grad_sub'], rtol=5e-4, atol=5e-7)
    s2 = s.detach().clone().requires_grad_()
    per = sd.ctc_loss(s2, tg.cuda(), tl.cuda(), reduction='none')
    clip = float(per.detach().sort().values[1])                      # clips the largest loss only
    sd.ctc_loss(s2, tg.cuda(), tl.cuda(), loss_clip=clip).backward()
    worst = int(per.detach().argmax())
    assert s2.grad[:, worst].abs().max().item() == 0.0
    assert s2.grad.abs().max().item() > 0.0



def test_read_set_pipeline_matches_basecall(model5):
    """Device-side chunking + stitching of a whole read set gives exactly the strings of the reference-shaped
    basecall() iterator (short, exact-multiple, stub and multi-chunk reads; ragged last batch)."""
    from xna_basecaller_b200 import pipeline
    from xna_basecaller_b200.crf import basecall

    class Read:
        def __init__(self, rid, sig):
            self.read_id, self.signal = rid, sig

    rs = np.random.RandomState(5)
    lengths = [700, 1000, 1001, 2350, 1900, 3100, 999, 4600, 1, 2800]
    sigs = [rs.randn(L).astype(np.float32) for L in lengths]
    want = [res['sequence'] for _, res in basecall(model5, iter([Read(i, s) for i, s in enumerate(sigs)]),
                                                   chunksize=1000, overlap=100, batchsize=4)]
    got, counters = pipeline.ReadSetBasecaller(model5, chunksize=1000, overlap=100, batchsize=7).basecall(sigs)
    assert got == want
    assert counters['reads'] == len(sigs) and counters['samples'] == sum(lengths)
    assert counters['chunks'] == sum(len(pipeline.plan_chunks([L], 1000, 100)['chunk_read']) for L in lengths)
    # int16 signal takes the same route (values are exactly representable)
    ints = [np.round(s * 50).astype(np.int16) for s in sigs[:4]]
    a, _ = pipeline.ReadSetBasecaller(model5, 1000, 100, 5).basecall(ints)
    b, _ = pipeline.ReadSetBasecaller(model5, 1000, 100, 5).basecall([x.astype(np.float32) for x in ints])
    assert a == b
    shard, table = pipeline.basecall_sharded(model5, sigs, 1000, 100, 7, rank=1, world=2)
    assert shard == {i: want[i] for i in range(1, len(sigs), 2)} and table['reads'] == [5.0]


def test_compute_scores_reverse_and_device_batches(model5):
    """reverse=True permutes the score tensor (CTC_CRF.reverse_complement) between encoder and decode; a batch that
    already lives on the device takes the same two-call route.  Both must equal the oracle's decode of the same scores."""
    from xna_basecaller_b200.crf.basecall import compute_scores
    x = synthetic_signal(33, 3, 500)
    crf = bo.CRF(3, ALPHABETS[5])
    with torch.no_grad():
        scores = model5(x.cuda()).cpu()
    for reverse in (False, True):
        s = crf.reverse_complement(scores) if reverse else scores
        want = bo.left_pack(crf.decode_batch(s), s.shape[0])
        got = compute_scores(model5, x.cuda() if not reverse else x, reverse=reverse)
        assert np.array_equal(got['sequence'].numpy(), want['sequence'])
        assert np.array_equal(got['qstring'].numpy(), want['qstring'])


def test_read_set_pipeline_edge_cases(model5):
    from xna_basecaller_b200 import pipeline
    caller = pipeline.ReadSetBasecaller(model5, chunksize=1000, overlap=100, batchsize=3)
    strings, counters = caller.basecall([])
    assert strings == [] and counters['reads'] == 0 and counters['chunks'] == 0
    one, c = caller.basecall([np.zeros(5, dtype=np.float32)])            # a single very short read: one padded chunk
    assert len(one) == 1 and c['chunks'] == 1
    rs = np.random.RandomState(9)
    sig = rs.randn(1000).astype(np.float32)                               # exactly one chunk: stitch is the identity
    (a,), _ = caller.basecall([sig])
    (b,), _ = pipeline.ReadSetBasecaller(model5, 1000, 100, 64).basecall([sig])
    assert a == b and len(a) > 0


def test_bf16_end_to_end_identical_read_rate(golden):
    """North-star item: with end-to-end bf16 the identical-read rate against the reference's fp32 run is reported, and
    reads that differ do so in a few near-tie steps only (bounded edit rate).  Reference-scale weights."""
    from make_golden import REF_SCALE
    from xna_basecaller_b200 import util
    cfg = sup_config(ALPHABETS[5])
    sd = bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE)
    sd['encoder.9.linear.weight'] = sd['encoder.9.linear.weight'] * 3.0     # avoid the all-blank degeneracy
    m = util.load_symbol(cfg, 'Model')(cfg)
    m.load_state_dict(sd)
    crf = bo.CRF(3, ALPHABETS[5])
    rs = np.random.RandomState(21)
    reads = [('r%d' % i, rs.randn(int(L)).astype(np.float32)) for i, L in enumerate([1000, 1700, 2350, 3100])]

    def ref_scores(batch):
        with torch.no_grad():
            return bo.encoder_forward(sd, batch, 5)

    want = {k: v['sequence'] for k, v in bo.basecall(ref_scores, crf, reads, 1000, 100, 4)}

    class Read:
        def __init__(self, rid, sig):
            self.read_id, self.signal = rid, sig

    for dtype, max_rate in ((torch.float16, 0.02), (torch.bfloat16, 0.03)):      # measured: 4/4 identical reads in both modes
        mm = m.to(dtype).eval().to('cuda')
        got = {r.read_id: res['sequence'] for r, res in
               util.load_symbol(cfg, 'basecall')(mm, iter([Read(k, s) for k, s in reads]), chunksize=1000, overlap=100, batchsize=4)}
        same = sum(got[k] == want[k] for k in want)
        worst = max(_edit_distance(got[k], want[k]) / max(len(want[k]), 1) for k in want)
        print('%s: identical-read rate %d/%d vs the fp32 reference path, worst edit rate %.3f, mean length %.0f'
              % (dtype, same, len(want), worst, np.mean([len(v) for v in want.values()])))
        assert worst <= max_rate


def test_compute_scores_beam_search_branch():
    """compute_scores' beam-search branch (crf/basecall.py:33-46: a model built with expand_blanks=False): scores without
    blank columns, blank score inserted, beam search -> {'sequence', 'qstring', 'moves'} with real qualities and moves."""
    from xna_basecaller_b200 import util
    from xna_basecaller_b200.crf.basecall import compute_scores
    from oracle import cexact
    cfg = sup_config(ALPHABETS[5])
    cfg['encoder']['expand_blanks'] = False
    m = util.load_symbol(cfg, 'Model')(cfg)
    m.load_state_dict(bo.reference_state_dict(n_base=5, seed=11))
    m = m.half().eval().to('cuda')
    x = synthetic_signal(35, 3, 500)
    out = compute_scores(m, x, beam_width=32, beam_cut=100.0, blank_score=2.0)
    assert set(out) == {'sequence', 'qstring', 'moves'} and out['moves'].shape == (3, 100) and out['moves'].dtype == bool
    with torch.no_grad():
        s = m(x.cuda()).float().cpu()                                   # (T, N, C * n_base): no blank columns
    full = torch.nn.functional.pad(s.view(100, 3, -1, 5), (1, 0), value=2.0).view(100, 3, -1)
    labels, quals = cexact.crf_beam_search(full.numpy(), 5)
    seq, _, lens = cexact.pack(labels, ALPHABETS[5])
    assert np.array_equal(out['sequence'].numpy(), seq) and np.array_equal(out['moves'], labels != 0)
    assert int(lens.sum()) > 0
    q = out['qstring'].numpy()
    assert ((q != 0) == (seq != 0)).all() and q[q != 0].min() >= 34


def test_read_set_pipeline_reverse_and_lazy_stream(model5):
    """(1) ReadSetBasecaller(reverse=True) (device-side reverse stitch, scores reverse-complemented between encoder and
    decode) gives the strings of the reference-shaped basecall(..., reverse=True); (2) basecall_reads is LAZY: it pulls
    reads from the iterator block by block (crf/basecall.py:96-119 hands its consumer an iterator, not a list) and yields
    results in input order while later blocks have not been pulled yet."""
    from xna_basecaller_b200 import pipeline
    from xna_basecaller_b200.crf import basecall

    class Read:
        def __init__(self, rid, sig):
            self.read_id, self.signal = rid, sig

    rs = np.random.RandomState(8)
    lengths = [700, 1000, 2350, 1900, 3100, 999, 4600, 2800, 1001, 1500]
    sigs = [rs.randn(L).astype(np.float32) for L in lengths]
    want = [res['sequence'] for _, res in basecall(model5, iter([Read(i, s) for i, s in enumerate(sigs)]),
                                                   chunksize=1000, overlap=100, batchsize=4, reverse=True)]
    got, _ = pipeline.ReadSetBasecaller(model5, chunksize=1000, overlap=100, batchsize=7, reverse=True).basecall(sigs)
    assert got == want and any(len(s) > 0 for s in got)

    pulled = []

    def source():
        for i, s in enumerate(sigs):
            pulled.append(i)
            yield Read(i, s)

    fwd = [res['sequence'] for _, res in basecall(model5, iter([Read(i, s) for i, s in enumerate(sigs)]),
                                                  chunksize=1000, overlap=100, batchsize=4)]
    out = pipeline.basecall_reads(model5, source(), chunksize=1000, overlap=100, batchsize=7, block_reads=3)
    first_read, first = next(out)
    assert first_read.read_id == 0 and first['sequence'] == fwd[0]
    assert len(pulled) <= 9, 'the whole iterator was consumed before the first result came out: %s' % pulled
    rest = [(r.read_id, res['sequence']) for r, res in out]
    assert [i for i, _ in rest] == list(range(1, len(sigs))) and [s for _, s in rest] == fwd[1:]
    assert first['qstring'] == 'O' * len(first['sequence']) and not first['sig_move'].any()
