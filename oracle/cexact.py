"""ctypes front-end of oracle/c/crf_exact.c (the bit-exact CPU checker for the CUDA decode).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libxb_oracle.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'c', 'crf_exact.c')
    hdr = os.path.join(_HERE, '..', 'xna_basecaller_b200', 'csrc', 'xb_exact_math.h')
    stale = (not os.path.exists(_SO)
             or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(['make', '-s', '-B', '-C', os.path.join(_HERE, 'c')])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(n_base, state_len):
    return n_base ** state_len, n_base + 1


def expf(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().xbo_expf_array(_p(x), _p(y), ctypes.c_long(x.size))
    return y


def logf(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().xbo_logf_array(_p(x), _p(y), ctypes.c_long(x.size))
    return y


def _apply(name, x):
    x = _f32(x)
    y = np.empty_like(x)
    getattr(lib(), name)(_p(x), _p(y), ctypes.c_long(x.size))
    return y


def expf_le0(x):
    return _apply('xbo_expf_le0_array', x)


def score_exp(x):
    """exp of a CRF score as the linear-domain decode defines it (clamp to [-80, 80], xb_expf_mid)."""
    return _apply('xbo_score_exp_array', x)


def logf_norm(x):
    return _apply('xbo_logf_norm_array', x)


def crf_alpha(scores, n_base, state_len=3):
    s = _f32(scores)
    T, N, _ = s.shape
    C, _ = _dims(n_base, state_len)
    out = np.empty((T + 1, N, C), dtype=np.float32)
    assert lib().xbo_crf_alpha(_p(s), T, N, n_base, state_len, _p(out)) == 0
    return out


def crf_logz(scores, n_base, state_len=3):
    s = _f32(scores)
    T, N, _ = s.shape
    out = np.empty(N, dtype=np.float32)
    assert lib().xbo_crf_logz(_p(s), T, N, n_base, state_len, _p(out)) == 0
    return out


def crf_decode_logdomain(scores, n_base, state_len=3, want_post=False, want_lp=False):
    """The log-domain restatement of decode_batch (round 1's kernel contract), kept as an independent second
    formulation to cross-check the linear-domain one: labels (N,T) int8 [, posteriors] [, log(post+1e-8)]."""
    s = _f32(scores)
    T, N, _ = s.shape
    labels = np.empty((N, T), dtype=np.int8)
    post = np.empty_like(s) if want_post else None
    lp = np.empty_like(s) if want_lp else None
    rc = lib().xbo_crf_decode(_p(s), T, N, n_base, state_len,
                              _p(post) if want_post else None, _p(lp) if want_lp else None, _p(labels))
    assert rc == 0
    out = [labels]
    if want_post:
        out.append(post)
    if want_lp:
        out.append(lp)
    return out[0] if len(out) == 1 else tuple(out)


def crf_decode(scores, n_base, state_len=3, want_post=False, exp_input=False, threads=1):
    """decode_batch as the CUDA kernels compute it -- the linear-domain (scaled) contract written above
    xbo_crf_decode_lin_range in c/crf_exact.c: labels (N,T) int8 [, posteriors (T,N,C*NZ)].
    exp_input: `scores` already holds exp(scores) (what the fused head hands to the decode).
    threads > 1 spreads the batch over host threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    s = _f32(scores)
    T, N, _ = s.shape
    labels = np.empty((N, T), dtype=np.int8)
    post = np.empty_like(s) if want_post else None
    threads = max(1, min(threads, N))
    bounds = [(N * i // threads, N * (i + 1) // threads) for i in range(threads)]
    fn = lib().xbo_crf_decode_lin_range

    def work(b):
        return fn(_p(s), int(exp_input), T, N, b[0], b[1], n_base, state_len, _p(post) if want_post else None, _p(labels))

    with ThreadPoolExecutor(threads) as ex:
        assert all(rc == 0 for rc in ex.map(work, bounds))
    return (labels, post) if want_post else labels


def crf_decode_threads(scores, n_base, state_len=3, threads=1):
    return crf_decode(scores, n_base, state_len, threads=threads)


def crf_beam_search(scores, n_base, state_len=3, beam_width=32, beam_cut=100.0):
    """Beam search over the CRF lattice (the contract of csrc/beam_search.cu): labels (N,T) int8, quals (N,T) uint8."""
    s = _f32(scores)
    T, N, _ = s.shape
    labels = np.empty((N, T), dtype=np.int8)
    quals = np.empty((N, T), dtype=np.uint8)
    rc = lib().xbo_crf_beam_search(_p(s), T, N, n_base, state_len, int(beam_width), ctypes.c_float(beam_cut), _p(labels), _p(quals))
    assert rc == 0, rc
    return labels, quals


def crf_viterbi(scores, n_base, state_len=3):
    s = _f32(scores)
    T, N, _ = s.shape
    labels = np.empty((N, T), dtype=np.int8)
    assert lib().xbo_crf_viterbi(_p(s), T, N, n_base, state_len, _p(labels)) == 0
    return labels


def pack(labels, alphabet):
    """path_to_str + left-pack: sequence (N,T) int8, qstring (N,T) int8, lens (N,) int32."""
    lab = np.ascontiguousarray(labels, dtype=np.int8)
    N, T = lab.shape
    seq = np.empty((N, T), dtype=np.int8)
    qs = np.empty((N, T), dtype=np.int8)
    lens = np.empty(N, dtype=np.int32)
    abc = ''.join(alphabet).encode()
    assert lib().xbo_pack(_p(lab), N, T, ctypes.c_char_p(abc), _p(seq), _p(qs), _p(lens)) == 0
    return seq, qs, lens


def strings(labels, alphabet):
    seq, _, lens = pack(labels, alphabet)
    return [seq[i, :lens[i]].astype('u1').tobytes().decode() for i in range(len(lens))]
