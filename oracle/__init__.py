"""CPU oracle for the ub-bonito basecalling forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, link or execute it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The product
package ``xna_basecaller_b200`` never imports this package and fails loudly when
its CUDA library is missing.

Layers of the oracle (all CPU):

``seqdist_restated``  restatement of the third-party ``ont-seqdist==0.0.4`` calls the
                      reference makes (not vendored in /root/reference, GPU-only cupy):
                      Log/Max semirings, sparse logZ with autograd, posteriors, ctc_simple.
``bonito_oracle``     restatement of the reference's own Python for the path
                      (``bonito/nn.py``, ``bonito/crf/model.py``, ``bonito/crf/basecall.py``,
                      ``bonito/util.py`` chunk/stitch/batchify), torch fp32 on CPU.
``refshim``           imports the REAL reference modules from /root/reference (only where
                      that tree exists, i.e. the build container) with the restated seqdist
                      plugged in; used to pin ``bonito_oracle`` and to generate
                      ``tests/golden/*.npz``.
``c/crf_exact.c``     plain-C restatement of the CRF decode arithmetic with the portable
                      exp/log of ``xb_exact_math.h``: the bit-exact checker for the CUDA decode.

PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path
 and ``ont-seqdist`` cannot be installed (no network, cupy/GPU
only).  The encoder half of the oracle IS pinned: it is checked against outputs of the
real ``bonito.nn`` / ``bonito.crf.model.Model`` code run in the build container
(``tests/golden/make_golden.py``).  The CRF arithmetic that lives in seqdist is
"parity unpinned" against seqdist itself; it is pinned instead against a dense
brute-force enumeration and against the reference's own ``CTC_CRF`` methods executed on
top of the restated seqdist.
"""
