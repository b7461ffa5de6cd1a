"""BASELINE configs[2]: sup@v3.3 UB XY (n_base 6, C=216, NZ=7): CRF forward-backward posteriors + Viterbi, batch 1024,
synthetic fp32 scores (800, 1024, 1512) with blank column 2.0 (seed 7).  Reports time per batch and the achieved
fraction of the HBM roofline for the algorithmic bytes (3*S + 1) per (t, chunk), S = 6048 B.
    python tools/run_config3.py [N]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xna_basecaller_b200._lib import Handle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T, C, NZ = 800, 216, 7
h = Handle('NACGTXY', 3, max_N=N, max_T=T, encoder=False)
g = torch.Generator(device='cuda').manual_seed(7)
s = torch.empty(T, N, C, NZ, device='cuda').uniform_(-5, 5, generator=g)
s[..., 0] = 2.0
s = s.reshape(T, N, -1)
for _ in range(3):
    out = h.decode(s, want_post=True, want_qstring=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
h.set_profiling(True); h.stage_times()
e0.record()
for _ in range(reps):
    out = h.decode(s, want_post=True, want_qstring=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
stages = {k: v[0] / reps for k, v in h.stage_times().items() if v[1]}
S = 4 * C * NZ
algo = T * N * (3 * S + 1)
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
print(json.dumps({'config': 'configs[2]: XY n_base 6, posteriors + Viterbi, batch %d x T 800, fp32 scores %.2f GB' % (N, T * N * S / 1e9),
                  'ms_per_batch': ms, 'samples_per_s': N * 4000 / (ms / 1e3), 'stages_ms': stages,
                  'algorithmic_GB': algo / 1e9, 'achieved_GBps': algo / (ms / 1e3) / 1e9, 'hbm_peak_GBps': peak,
                  'frac_of_hbm_roofline': algo / (ms / 1e3) / 1e9 / peak, 'mean_decoded_len': float(out[2].float().mean())}))
