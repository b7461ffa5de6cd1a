"""Debug helper: per-phase clock64 timeline of CTA 0 of the persistent LSTM kernel (XB_LSTM_DEBUG=1)."""
import ctypes, os, sys
os.environ['XB_LSTM_DEBUG'] = '1'
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bonito_oracle as bo
from xna_basecaller_b200._lib import Handle
N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 200
h = Handle('NACGTX', 3, max_N=N, max_T=T)
h.load_weights(bo.reference_state_dict(5, seed=1))
x = (torch.randn(T, N, 768, device='cuda') * 0.5).half()
for _ in range(2):
    y = h.lstm(0, x, False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
h.lib.xb_debug_lstm_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert h.lib.xb_debug_lstm_timeline(h.h, buf) == 0
a = np.array(buf[:]).reshape(8, 16)
names = ['poll_start', 'poll_done', 'tma_issued', 'mma_dempty', 'mma_hfull', 'mma_issued', 'mma_x', 'epi_dfull', 'epi_p1', 'epi_p2', 'epi_stored', 'epi_red']
for s in range(1, 8):
    base = a[s, 0]
    print(os.environ.get('XB_LSTM_NO3D', '3d'), 'step', 64 + s, ' '.join('%s=%d' % (n, a[s, i] - base) for i, n in enumerate(names)), ' | step period', a[s, 0] - a[s - 1, 0])
