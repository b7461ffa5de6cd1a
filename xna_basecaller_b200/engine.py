"""Engine: the xb_handle a model (or a stand-alone layer / CTC_CRF object) computes through.

One Engine per module tree.  It creates the C-ABI handle lazily on the device of the first call, sized to
the largest (N, T) seen so far, and pushes the module's parameters through xb_load_*_weights whenever they
change (load_state_dict, .half(), .to(), an optimiser step): parameters stay ordinary torch Parameters under
the reference's state_dict keys, the kernel-layout copies live inside the handle.
"""
import torch

from . import _lib


def _stamp(tensors):
    return tuple((t.data_ptr(), t._version, t.dtype, str(t.device)) for t in tensors if t is not None)


class Engine:
    def __init__(self, alphabet='NACGT', state_len=3):
        self.alphabet = ''.join(alphabet)
        self.state_len = state_len
        self.handle = None
        self.key = None
        self.stamps = {}

    # ------------------------------------------------------------------ handle life cycle
    def get(self, device, N, T, bf16=False, encoder=True, train=False):
        """Handle on `device` with capacity >= (N, T); train=True: one that keeps what the encoder backward needs."""
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('xna_basecaller_b200 computes on a CUDA device (sm_100a) only; got tensors on %s. '
                               'There is no CPU fallback.' % device)
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        h = self.handle
        if (h is None or h.device != device or h.bf16 != bf16 or h.max_N < N or h.max_T < T
                or (encoder and not h.has_encoder) or (train and not h.train)):
            max_N = max(N, h.max_N if h is not None and h.device == device else 0)
            max_T = max(T, h.max_T if h is not None and h.device == device else 0)
            need_enc = encoder or (h is not None and h.has_encoder)
            train = train or (h is not None and h.train)
            # The outgoing handle is NOT destroyed here: an in-flight pipelined batch (crf/basecall.py::_submit_scores), a
            # _CTCLoss context between forward and backward or a ReadSetBasecaller may still hold it.  Dropping our
            # reference lets Handle.__del__ run xb_destroy when its last user lets go.
            self.handle = _lib.Handle(self.alphabet, self.state_len, max_N=max_N, max_T=max_T, device=device,
                                      bf16=bf16, encoder=need_enc, train=train)
            self.handle.has_encoder = need_enc
            self.stamps = {}
        return self.handle

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None
            self.stamps = {}

    # ------------------------------------------------------------------ weights
    def sync(self, slot, tensors, loader):
        """Re-run `loader(handle)` when any tensor of weight group `slot` changed since the last call."""
        stamp = _stamp(tensors)
        if self.stamps.get(slot) != stamp:
            loader(self.handle)
            self.stamps[slot] = stamp
