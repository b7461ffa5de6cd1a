"""Stall breakdown of the A-stationary GEMM launches of one encoder pass (XB_INPROJ_DEBUG=1): five input projections + head."""
import os, sys
os.environ['XB_INPROJ_DEBUG'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import bonito_oracle as bo
from xna_basecaller_b200._lib import Handle
h = Handle('NACGTX', 3, max_N=512, max_T=800)
h.load_weights(bo.reference_state_dict(n_base=5, seed=25))
x = torch.randn(512, 4000, device='cuda')
for _ in range(2):
    s = h.encoder(x)
torch.cuda.synchronize()
