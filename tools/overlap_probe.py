"""Does the CRF decode of batch i hide under the encoder of batch i+1 when they run on two streams?
    python tools/overlap_probe.py [steps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bonito_oracle as bo            # weight generator only
from xna_basecaller_b200._lib import Handle

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = 512
h = Handle('NACGTX', 3, max_N=N, max_T=800)
h.load_weights(bo.reference_state_dict(n_base=5, seed=25))
x = torch.randn(N, 4000, generator=torch.Generator().manual_seed(1234)).cuda()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def run(K, overlap):
    keep = []
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(sA)
    for i in range(K):
        with torch.cuda.stream(sA):
            sc = h.encoder(x)
            e = torch.cuda.Event()
            e.record(sA)
        sd = sB if overlap else sA
        with torch.cuda.stream(sd):
            sd.wait_event(e)
            out = h.decode(sc, want_qstring=False)
        keep.append((sc, out))
    sA.wait_stream(sB)
    t1.record(sA)
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / K, keep[-1][1][2][:4].tolist()


run(3, False)
for ov in (False, True, False, True):
    ms, lens = run(K, ov)
    print('overlap' if ov else 'serial ', '%.2f ms/step' % ms, lens)
