"""Generate tests/golden/*.npz by running the REAL reference code (build container only).

    python tests/golden/make_golden.py

Everything here executes /root/reference/ub-bonito/bonito/{nn,util,crf/model,crf/basecall}.py
through ``oracle.refshim`` (the only stand-ins are the absent third-party modules: seqdist is
``oracle.seqdist_restated``, koi.decode.to_str is restated).  Inputs are NOT stored: tests
regenerate them from the seeds below (numpy RandomState / torch.Generator on CPU are stable
for a fixed build), so the fixtures stay small.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim, bonito_oracle as bo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ALPHABETS = {4: ['N', 'A', 'C', 'G', 'T'], 5: ['N', 'A', 'C', 'G', 'T', 'X'],
             6: ['N', 'A', 'C', 'G', 'T', 'X', 'Y']}


REF_SCALE = dict(head_gain=1.0, head_shift=0.0, wih_gain=1.0, whh_gain=1.0)


def synthetic_scores(seed, T, N, n_base, state_len=3, blank=2.0):
    """SURVEY 8c golden (ii) / config 3: non-blank ~U(-5,5), blank column constant."""
    rs = np.random.RandomState(seed)
    C = n_base ** state_len
    s = rs.uniform(-5, 5, size=(T, N, C, n_base + 1)).astype(np.float32)
    s[..., 0] = blank
    return torch.from_numpy(s.reshape(T, N, -1))


def synthetic_targets(seed, N, n_base, lo, hi):
    rs = np.random.RandomState(seed)
    lengths = rs.randint(lo, hi + 1, size=N)
    tg = np.zeros((N, hi), dtype=np.int64)
    for i, l in enumerate(lengths):
        tg[i, :l] = rs.randint(1, n_base + 1, size=l)
    return torch.from_numpy(tg), torch.from_numpy(lengths.astype(np.int64))


def synthetic_signal(seed, N, L):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(N, 1, L, generator=g)


def make_crf_grad(RM):
    """Gradient of the reference's CTC_CRF.ctc_loss w.r.t. the scores (autograd through the restated seqdist), sub-sampled."""
    out = {}
    for n_base, alphabet in ALPHABETS.items():
        sd = RM.CTC_CRF(3, alphabet)
        for seed in (0, 1):
            s = synthetic_scores(seed, 160, 3, n_base).requires_grad_()
            tg, tl = synthetic_targets(100 + seed, 3, n_base, 30, 50)
            sd.ctc_loss(s, tg, tl).backward()                      # reduction='mean', normalise_scores=True
            key = 'n%d_s%d_' % (n_base, seed)
            out[key + 'grad_sub'] = s.grad[::9, :, ::7].numpy()
            out[key + 'grad_abs_rowsum'] = s.grad.abs().sum(2).numpy()
    np.savez_compressed(os.path.join(OUT, 'crf_grad.npz'), **out)
    print('crf_grad.npz', os.path.getsize(os.path.join(OUT, 'crf_grad.npz')))


def synthetic_raw_read(seed, n, stall=True):
    """int16 DAC samples with the features the pre-processing reacts to: k-mer-like level steps, an open-pore stall at
    the start (what trim() removes) and a quiet stretch (what norm_by_noisiest_section avoids)."""
    rs = np.random.RandomState(seed)
    x = 400 + 60 * rs.randn() + 35 * rs.randn(n) + np.repeat(rs.randn(n // 8 + 1) * 25, 8)[:n]
    if stall:
        k = min(rs.randint(100, 1500), n)
        x[:k] += 250 + 30 * rs.randn(k)
    if n > 3000:
        q = rs.randint(n // 3, n // 2)
        x[q:q + rs.randint(200, 900)] *= 0.2
    return np.clip(np.round(x), -2000, 2000).astype(np.int16)


RAW_LENGTHS = (12000, 9000, 8700, 8011, 7000, 4000, 2500, 1000, 450, 150, 99, 30, 12, 15001)
RAW_SCALING = 1437.976 / 8192


def make_preprocess():
    """The reference's own Read.__init__ arithmetic (fast5.py:88-100) on synthetic raw reads."""
    import warnings
    F5 = refshim.install_io()['bonito.fast5']
    out = {}
    for i, n in enumerate(RAW_LENGTHS):
        raw, offset = synthetic_raw_read(1000 + i, n, stall=i % 3 != 2), -240 + i
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            scaled = np.array(np.float64(RAW_SCALING) * (raw + offset), dtype=np.float32)
            trim_start, _ = F5.trim(scaled[:8000])
            scaled = scaled[trim_start:]
            if len(scaled) > 8000:
                med, mad = F5.med_mad(scaled)
                signal = (scaled - med) / mad
            else:
                signal = F5.norm_by_noisiest_section(scaled)
        out['r%d_trim' % i] = np.array(trim_start)
        out['r%d_signal' % i] = signal.astype(np.float32)
        assert signal.dtype == np.float32
    np.savez_compressed(os.path.join(OUT, 'preprocess.npz'), **out)
    print('preprocess.npz', os.path.getsize(os.path.join(OUT, 'preprocess.npz')))


class FakeRead:
    """The attributes of bonito.fast5.Read that Writer / summary_row touch (fast5.py:22-128)."""

    def __init__(self, i, n):
        self.read_id, self.filename, self.run_id = 'read-%04d' % i, 'batch_%d.fast5' % (i // 2), 'run%02d' % (i % 3)
        self.channel, self.mux, self.read_number = 100 + i, 1 + i % 4, 7 * i
        self.start, self.duration = 1.5 * i, 2.25 + i
        self.template_start, self.template_duration = self.start + 0.01, self.duration - 0.01
        self.start_time = '2021-03-0%dT10:00:0%d' % (1 + i % 9, i % 10)
        self.signal = np.zeros(n, dtype=np.float32)

    def tagdata(self):
        return ['mx:i:%s' % self.mux, 'ch:i:%s' % self.channel, 'st:Z:%s' % self.start_time, 'rn:i:%s' % self.read_number,
                'f5:Z:%s' % self.filename]


def fake_results():
    rs = np.random.RandomState(5)
    out = []
    for i, n in enumerate((1200, 0, 4000, 333)):
        seq = ''.join(rs.choice(list('ACGTX'), size=n // 10))
        out.append((FakeRead(i, n), {'sequence': seq, 'qstring': 'O' * len(seq), 'sig_move': np.zeros(n, dtype=bool)}))
    return out


def make_io():
    """FASTQ text + summary table written by the reference's own Writer thread (mode 'wfq') for four fake reads."""
    import io as pyio
    import json
    import tempfile
    mods = refshim.install_io()
    RIO = mods['bonito.io']
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            fd = pyio.StringIO()
            RIO.summary_file = lambda: os.path.join(tmp, 'summary.tsv')     # location only; the format is the reference's
            w = RIO.Writer('wfq', iter(fake_results()), aligner=None, fd=fd, groups=[], group_key='sup_v3.3')
            w.start(); w.join()
            summary = open([f for f in os.listdir(tmp) if f.endswith('summary.tsv')][0]).read()
        finally:
            os.chdir(cwd)
    buf = pyio.StringIO()
    RIO.write_fasta('r1 extra', 'ACGTX', fd=buf)
    RIO.write_fastq('r2', 'ACX', 'OOO', fd=buf)
    RU = mods['bonito.util']
    out = {'fastq': fd.getvalue(), 'summary': summary, 'log': w.log, 'records': buf.getvalue(),
           'qscores': {q: float(RU.mean_qscore_from_qstring(q)) for q in ('OOOO', '!5I', '+', 'IIIIIIIIII5')},
           'row_unaligned': {k: (v if isinstance(v, (str, int, float)) else float(v))
                             for k, v in RIO.summary_row(FakeRead(3, 10), 17, 12.5, alignment=None).items()}}
    with open(os.path.join(OUT, 'io.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('io.json', os.path.getsize(os.path.join(OUT, 'io.json')))


def main():
    if '--only-io' in sys.argv:
        return make_io()
    if '--only-preprocess' in sys.argv:
        return make_preprocess()
    mods = refshim.install()
    RM, RU, RB = mods['bonito.crf.model'], mods['bonito.util'], mods['bonito.crf.basecall']
    if '--only-crf-grad' in sys.argv:
        return make_crf_grad(RM)
    make_crf_grad(RM)
    make_io()
    make_preprocess()

    # ---- (i)+(ii) CRF: reference CTC_CRF methods on synthetic scores
    crf = {}
    for n_base, alphabet in ALPHABETS.items():
        sd = RM.CTC_CRF(3, alphabet)
        for seed in (0, 1):
            T, N = 160, 3
            s = synthetic_scores(seed, T, N, n_base)
            key = 'n%d_s%d_' % (n_base, seed)
            post = sd.posteriors(s)
            crf[key + 'logZ'] = sd.logZ(s).numpy()
            crf[key + 'post_sub'] = post[::9, :, ::7].numpy()
            crf[key + 'post_rowsum'] = post.sum(2).numpy()
            lp = (post + 1e-8).log()
            crf[key + 'paths'] = sd.viterbi(lp).numpy().astype(np.int8)
            crf[key + 'paths_raw'] = sd.viterbi(s).numpy().astype(np.int8)
            model = RM.SeqdistModel(torch.nn.Identity(), sd)
            crf[key + 'strings'] = np.array(model.decode_batch(s))
            tg, tl = synthetic_targets(100 + seed, N, n_base, 30, 50)
            crf[key + 'ctc_loss'] = sd.ctc_loss(s, tg, tl, reduction='none').numpy()
            crf[key + 'alpha_last'] = sd.forward_scores(s)[-1].numpy()
            crf[key + 'beta_first'] = sd.backward_scores(s)[0].numpy()
    np.savez_compressed(os.path.join(OUT, 'crf.npz'), **crf)

    # ---- (iii) encoder: reference Model with deterministic weights
    enc = {}
    for n_base in (5, 6):
        cfg = refshim.reference_config(ALPHABETS[n_base])
        model = RM.Model(cfg).eval()
        sd = bo.reference_state_dict(n_base=n_base, seed=11)
        model.load_state_dict(sd)
        x = synthetic_signal(21, 2, 500)
        with torch.no_grad():
            stem = model.encoder[2](model.encoder[1](model.encoder[0](x)))
            l1 = model.encoder[4](stem.permute(2, 0, 1))
            scores = model(x)
        key = 'n%d_' % n_base
        enc[key + 'stem_sub'] = stem[:, ::16, :].numpy()
        enc[key + 'lstm1_sub'] = l1[:, :, ::16].numpy()
        enc[key + 'scores'] = scores.numpy()
        enc[key + 'strings'] = np.array(model.decode_batch(scores))
        cs = RB.compute_scores(model, x)
        enc[key + 'cs_sequence'] = cs['sequence'].numpy()
        enc[key + 'cs_qstring'] = cs['qstring'].numpy()
        enc[key + 'cs_moves'] = np.asarray(cs['moves'])
    # reference-scale weights (unit gains, as the reference's own init): the <=1e-2 tolerance case
    cfg = refshim.reference_config(ALPHABETS[5])
    model = RM.Model(cfg).eval()
    model.load_state_dict(bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE))
    x = synthetic_signal(23, 2, 500)
    with torch.no_grad():
        enc['r5_scores'] = model(x).numpy()
    np.savez_compressed(os.path.join(OUT, 'encoder.npz'), **enc)

    # ---- (iv) chunk / stitch on index arrays
    st = {}
    for cs_, ov in ((4000, 500), (3600, 500), (1000, 100)):
        for L in (cs_ - 1, cs_, cs_ + 1, 7500, 10000, 10001, 2 * cs_ - ov, 19999):
            sig = torch.arange(L, dtype=torch.float32)
            ch = RU.chunk(sig, cs_, ov)
            key = 'c%d_o%d_L%d_' % (cs_, ov, L)
            st[key + 'first'] = ch[:, 0, 0].numpy().astype(np.int64)       # start sample of each chunk
            st[key + 'lastrow'] = ch[-1, 0, -3:].numpy().astype(np.int64)
            T = cs_ // 5
            lab = torch.arange(ch.shape[0] * T, dtype=torch.int32).reshape(ch.shape[0], T)
            st[key + 'stitched'] = RU.stitch(lab, cs_, ov, L, 5).numpy()
            st[key + 'stitched_rev'] = RU.stitch(lab, cs_, ov, L, 5, reverse=True).numpy()
    np.savez_compressed(os.path.join(OUT, 'stitch.npz'), **st)

    # ---- end to end: reference basecall() on a few short reads (CPU, fp32)
    class Read:
        def __init__(self, rid, sig):
            self.read_id, self.signal = rid, sig

    cfg = refshim.reference_config(ALPHABETS[5])
    model = RM.Model(cfg).eval()
    model.load_state_dict(bo.reference_state_dict(n_base=5, seed=11))
    rs = np.random.RandomState(77)
    lengths = [700, 1000, 1001, 2350, 1900, 3100]
    reads = [Read('read%d' % i, rs.randn(L).astype(np.float32)) for i, L in enumerate(lengths)]
    e2e = {'lengths': np.array(lengths)}
    for rd, res in RB.basecall(model, reads, chunksize=1000, overlap=100, batchsize=4):
        e2e[rd.read_id + '_sequence'] = np.array(res['sequence'])
        e2e[rd.read_id + '_qstring'] = np.array(res['qstring'])
        e2e[rd.read_id + '_sig_move_len'] = np.array(len(res['sig_move']))
    np.savez_compressed(os.path.join(OUT, 'basecall.npz'), **e2e)
    for f in ('crf.npz', 'encoder.npz', 'stitch.npz', 'basecall.npz'):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
