"""Throughput of xb_preprocess_reads on a config-4-like read set (lengths ~ N(10000, 1000) clipped to [4000, 20000]).
    python tools/preprocess_bench.py [n_reads]
Algorithmic bytes: 2 B read (int16) + 4 B written (fp32) per sample."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xna_basecaller_b200._lib import Handle

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rs = np.random.RandomState(11)
lens = np.clip(rs.normal(10000, 1000, n_reads), 4000, 20000).astype(np.int64)
off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
total = int(lens.sum())
raw = (400 + 40 * torch.randn(total, device='cuda')).to(torch.int16)
step = torch.repeat_interleave(torch.randn(total // 8 + 1, device='cuda') * 25, 8)[:total]
raw = (raw.float() + step).to(torch.int16)
h = Handle('NACGTX', 3, max_N=4, max_T=40, encoder=False)
scal = np.full(n_reads, 1437.976 / 8192)
offs = rs.randint(-300, 300, n_reads).astype(np.int32)
for _ in range(2):
    out = h.preprocess(raw, off, lens, scal, offs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ro, rl = torch.as_tensor(off, device='cuda'), torch.as_tensor(lens, dtype=torch.int32, device='cuda')
sc, of = torch.as_tensor(scal, device='cuda'), torch.as_tensor(offs, device='cuda')
a.record()
for _ in range(5):
    out = h.preprocess(raw, ro, rl, sc, of)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(json.dumps({'reads': n_reads, 'samples': total, 'ms': ms, 'samples_per_s': total / ms * 1e3,
                  'algorithmic_GBps': total * 6 / ms / 1e6, 'trim_mean': float(out[2][:, 0].mean())}))
