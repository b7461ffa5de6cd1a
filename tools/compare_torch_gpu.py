"""Context number: the reference's PyTorch path on the same B200 -- its nn modules in fp16
through torch's cuDNN / cuBLAS (restated functionally in oracle/bonito_oracle.py) and the restated-seqdist torch ops
for the CRF decode (the real seqdist wheel is unavailable).  Not part of the product or of bench.py.
    python tools/compare_torch_gpu.py [N]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bonito_oracle as bo

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device('cuda')
sd = {k: v.to(dev).half() for k, v in bo.reference_state_dict(n_base=5, seed=25).items()}
x = torch.randn(N, 1, 4000, generator=torch.Generator().manual_seed(1234)).to(dev).half()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


with torch.no_grad():
    enc_s, scores = timed(lambda: bo.encoder_forward(sd, x, 5, library=True))
    crf = bo.CRF(3, list('NACGTX'))
    for name in ('idx', 'src_edges', 'src_dst'):
        setattr(crf, name, getattr(crf, name).to(dev))
    nd = min(N, 64)                                      # the pure-torch scan is slow: time a slice, scale linearly
    sub = scores[:, :nd].float().contiguous()

    def decode():
        post = crf.posteriors(sub, 'log') + 1e-8
        return crf.viterbi(post.log())
    dec_s, _ = timed(decode, reps=1)
dec_full = dec_s * N / nd
print(json.dumps({'what': 'reference PyTorch path on this GPU (torch %s: cuDNN LSTM fp16, cuBLAS; restated-seqdist torch CRF)' % torch.__version__,
                  'N': N, 'encoder_ms': 1e3 * enc_s, 'encoder_samples_per_s': N * 4000 / enc_s,
                  'decode_ms_scaled_from_%d_chunks' % nd: 1e3 * dec_full,
                  'end_to_end_samples_per_s': N * 4000 / (enc_s + dec_full)}))
