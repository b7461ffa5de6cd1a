"""Debug helper: per-phase clock64 timeline of CTA 0 of the persistent LSTM kernel (XB_LSTM_DEBUG=1), all sub-batches."""
import ctypes, os, sys
os.environ['XB_LSTM_DEBUG'] = '1'
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bonito_oracle as bo
from xna_basecaller_b200._lib import Handle
N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 200
SUB = 6 if os.environ.get('XB_LSTM_VARIANT') == '1' else 3
h = Handle('NACGTX', 3, max_N=N, max_T=T)
h.load_weights(bo.reference_state_dict(5, seed=1))
x = (torch.randn(T, N, 768, device='cuda') * 0.5).half()
for _ in range(2):
    y = h.lstm(0, x, False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
h.lib.xb_debug_lstm_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert h.lib.xb_debug_lstm_timeline(h.h, buf) == 0
a = np.array(buf[:]).reshape(8, 8, 16)
names = ['poll0', 'poll1', 'tma', 'mma_rdy', 'mma_iss', '-', 'e_transp', 'e_math', 'tma_done', 'e_dfull', 'e_ld', 'e_gfull', 'e_stored', 'e_red']
base = a[1, 0, 0]
for s in range(1, 7):
    for sub in range(SUB):
        print('s%d sub%d ' % (64 + s, sub) + ' '.join('%s=%d' % (n, a[s, sub, i] - base) for i, n in sorted(enumerate(names), key=lambda kv: a[s, sub, kv[0]]) if n != '-'))
