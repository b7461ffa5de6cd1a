import os, sys
os.environ['XB_INPROJ_DEBUG']='1'
sys.path.insert(0,'/root/repo')
import torch
from oracle import bonito_oracle as bo
from xna_basecaller_b200._lib import Handle
h = Handle('NACGTX', 3, max_N=512, max_T=800)
h.load_weights(bo.reference_state_dict(n_base=5, seed=25))
x = (torch.randn(800, 512, 768, device='cuda')*0.5).half()
for _ in range(2):
    s = h.crf_head(x)
torch.cuda.synchronize()
