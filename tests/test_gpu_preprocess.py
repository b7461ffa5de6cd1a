"""xb_preprocess_reads (csrc/preprocess.cu) against oracle/preprocess.py and the reference-generated golden outputs:
same trim index, same float32 bits."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
from make_golden import RAW_LENGTHS, RAW_SCALING, synthetic_raw_read

from oracle import preprocess as pp

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preprocess.npz'))


@pytest.fixture(scope='module')
def handle():
    from xna_basecaller_b200._lib import Handle
    return Handle('NACGTX', 3, max_N=4, max_T=40, encoder=False)


def run(handle, raws, scalings, offsets):
    lens = np.array([len(r) for r in raws], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    raw = torch.from_numpy(np.concatenate(raws) if lens.sum() else np.zeros(0, dtype=np.int16))
    out, out_len, stats = handle.preprocess(raw, off, lens, scalings, offsets)
    torch.cuda.synchronize()
    out, out_len, stats = out.cpu().numpy(), out_len.cpu().numpy(), stats.cpu().numpy()
    return [out[o:o + n] for o, n in zip(off, out_len)], out_len, stats


def test_matches_reference_golden(handle):
    raws = [synthetic_raw_read(1000 + i, n, stall=i % 3 != 2) for i, n in enumerate(RAW_LENGTHS)]
    sigs, out_len, stats = run(handle, raws, [RAW_SCALING] * len(raws), [-240 + i for i in range(len(raws))])
    for i, sig in enumerate(sigs):
        want = GOLD['r%d_signal' % i]
        assert int(stats[i, 0]) == int(GOLD['r%d_trim' % i])
        assert out_len[i] == len(want)
        assert np.array_equal(sig.view(np.uint32), want.view(np.uint32)), i
    assert set(stats[:, 3].astype(int)) == {0, 1}


def test_matches_oracle_many_reads(handle):
    rs = np.random.RandomState(3)
    lengths = [int(x) for x in np.clip(rs.normal(9000, 3000, 60), 20, 30000)] + [8010, 8011, 7999, 100, 99, 101, 11, 10, 3, 0]
    raws = [synthetic_raw_read(50 + i, n, stall=i % 4 != 3) if n else np.zeros(0, dtype=np.int16) for i, n in enumerate(lengths)]
    scalings = (0.15 + 0.05 * rs.rand(len(raws))).astype(np.float64)
    offsets = rs.randint(-300, 300, len(raws)).astype(np.int32)
    sigs, out_len, stats = run(handle, raws, scalings, offsets)
    for i, raw in enumerate(raws):
        want, start, med, mad, mode = pp.preprocess(raw, scalings[i], int(offsets[i]))
        assert (int(stats[i, 0]), int(stats[i, 3]), out_len[i]) == (start, mode, len(want)), (i, len(raw))
        assert np.array_equal(sigs[i].view(np.uint32), want.view(np.uint32)), (i, len(raw))
        if mode != 2:
            assert stats[i, 1] == med and stats[i, 2] == mad


def test_degenerate_signals(handle):
    """Constant reads (MAD = eps), a read with heavy ties, a very long read (radix select over 600k samples)."""
    rs = np.random.RandomState(8)
    raws = [np.full(9000, 321, dtype=np.int16), np.full(500, -7, dtype=np.int16),
            rs.randint(0, 4, 12000).astype(np.int16), (400 + 30 * rs.randn(600000)).astype(np.int16)]
    sigs, out_len, stats = run(handle, raws, [RAW_SCALING] * 4, [0, 5, 0, -100])
    for i, raw in enumerate(raws):
        want, start, med, mad, mode = pp.preprocess(raw, RAW_SCALING, [0, 5, 0, -100][i])
        assert int(stats[i, 0]) == start and out_len[i] == len(want)
        assert np.array_equal(sigs[i].view(np.uint32), want.view(np.uint32)), i


def test_raw_reads_through_the_read_set_pipeline():
    """Raw int16 reads + (scaling, offset) through ReadSetBasecaller == the same reads normalised by the oracle first."""
    from oracle import bonito_oracle as bo
    from xna_basecaller_b200.crf import Model
    from xna_basecaller_b200.pipeline import ReadSetBasecaller
    from test_cpu_host import sup_config
    model = Model(sup_config(list('NACGTX')))
    model.load_state_dict(bo.reference_state_dict(n_base=5, seed=11))
    model = model.half().eval().to('cuda')
    raws = [synthetic_raw_read(200 + i, n) for i, n in enumerate((5200, 9100, 3000, 12500))]
    scal, offs = [RAW_SCALING] * 4, [-240, -100, 0, 55]
    caller = ReadSetBasecaller(model, chunksize=2000, overlap=200, batchsize=8)
    got, counters = caller.basecall(raws, scaling=scal, offset=offs)
    want, _ = caller.basecall([pp.preprocess(r, s, o)[0] for r, s, o in zip(raws, scal, offs)])
    assert got == want and all(len(s) > 0 for s in got)
    assert counters['samples'] == sum(len(r) for r in raws)


def test_raw_reads_to_fastq():
    """Raw reads -> xb_preprocess_reads -> encoder / decode / stitch on the device -> io.Writer: FASTQ records + summary rows."""
    import io as pyio
    from oracle import bonito_oracle as bo
    from test_cpu_host import sup_config
    from xna_basecaller_b200 import io as xio
    from xna_basecaller_b200.crf import Model
    from xna_basecaller_b200.pipeline import basecall_reads

    class Read:
        def __init__(self, i, raw):
            self.read_id, self.signal, self.scaling, self.offset = 'raw-%d' % i, raw, RAW_SCALING, -200 + 10 * i
            self.run_id, self.filename, self.channel, self.mux = 'runA', 'f.fast5', 7 + i, 1
            self.start, self.duration, self.template_start, self.template_duration = 0.0, 1.0, 0.0, 1.0

    model = Model(sup_config(list('NACGTX')))
    model.load_state_dict(bo.reference_state_dict(n_base=5, seed=11))
    model = model.half().eval().to('cuda')
    reads = [Read(i, synthetic_raw_read(300 + i, n)) for i, n in enumerate((5200, 9100, 3000))]
    fd = pyio.StringIO()
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        w = xio.Writer('wfq', basecall_reads(model, reads, chunksize=2000, overlap=200, batchsize=8, raw=True), fd=fd,
                       group_key='sup', summary=os.path.join(tmp, 'summary.tsv'))
        w.start()
        w.join()
        rows = open(os.path.join(tmp, 'summary.tsv')).read().splitlines()
    lines = fd.getvalue().splitlines()
    assert len(lines) == 4 * len(reads) and len(rows) == 1 + len(reads)
    for i, read in enumerate(reads):
        assert lines[4 * i].startswith('@raw-%d RG:Z:runA_sup\tqs:i:40' % i)
        seq, qual = lines[4 * i + 1], lines[4 * i + 3]
        assert len(seq) > 0 and set(seq) <= set('ACGTX') and qual == 'O' * len(seq)
    assert [x[0] for x in w.log] == [r.read_id for r in reads]
