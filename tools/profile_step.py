"""One configs[1] step (N=512 x 4000 samples, n_base 5) of the hot path, for ncu captures:
    python tools/profile_step.py [N] [steps] [n_base]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bonito_oracle as bo            # weight generator only
from xna_basecaller_b200._lib import Handle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_base = int(sys.argv[3]) if len(sys.argv) > 3 else 5
alphabet = 'NACGTXY'[:n_base + 1]
h = Handle(alphabet, 3, max_N=N, max_T=800)
h.load_weights(bo.reference_state_dict(n_base=n_base, seed=25))
x = torch.randn(N, 4000, generator=torch.Generator().manual_seed(1234)).cuda()
for _ in range(steps):
    s = h.encoder(x)
    out = h.decode(s, want_qstring=False)
torch.cuda.synchronize()
print('ok', h.launches, 'launches; decoded lens', out[2][:4].tolist())
