"""CTC-CRF model with the reference's class surface (ub-bonito/bonito/crf/model.py): get_stride (:14-21),
CTC_CRF (:24-135), conv / rnn_encoder (:138-201), SeqdistModel (:204-221), Model (:224-237).

CTC_CRF's scans (logZ, forward/backward scores, posteriors, viterbi, ctc_loss) are the seqdist calls of the
reference (seqdist.sparse.logZ / fwd_scores_cupy / bwd_scores_cupy, ctc_simple.logZ_cupy) re-implemented as
CUDA kernels in libxna_b200.so (crf_decode.cu, misc_kernels.cu); SeqdistModel.decode_batch runs the fused
posteriors -> log -> max-marginal arg-max -> label path (xb_crf_decode) and only ships (N, T) int8 to the host.
"""
import numpy as np
import torch

from ..engine import Engine
from ..nn import Module, Convolution, LinearCRFEncoder, Serial, Permute, layers, from_dict


class Log:
    """Marker for the log semiring (seqdist.core.Log): sum = logsumexp, dsum = softmax."""
    zero, one = -1e38, 0.0


class Max:
    """Marker for the max semiring (seqdist.core.Max): sum = max, dsum = one-hot at the arg-max."""
    zero, one = -1e38, 0.0


def get_stride(m):
    if hasattr(m, 'stride'):
        return m.stride if isinstance(m.stride, int) else m.stride[0]
    if isinstance(m, Convolution):
        return get_stride(m.conv)
    if isinstance(m, Serial):
        return int(np.prod([get_stride(x) for x in m]))
    return 1


class _CTCLoss(torch.autograd.Function):
    """Per-sequence CTC-CRF loss with its gradient w.r.t. the scores (xb_ctc_crf_loss_fwd / _bwd), so that
    `loss.backward()` of the reference's training step (training.py:100-108) reaches the score tensor."""

    @staticmethod
    def forward(ctx, scores, handle, targets, lengths, normalise):
        ctx.handle, ctx.normalise = handle, normalise
        ctx.save_for_backward(scores, targets, lengths)
        return handle.ctc_loss(scores, targets, lengths, normalise=normalise)

    @staticmethod
    def backward(ctx, grad):
        scores, targets, lengths = ctx.saved_tensors
        g = ctx.handle.ctc_loss_bwd(scores, targets, lengths, grad, normalise=ctx.normalise)
        return g.to(scores.dtype), None, None, None, None


class CTC_CRF:
    """Transition lattice over n_base^state_len states with n_base+1 edges into every state: edge 0 stays
    (blank), edge 1+j arrives from the state whose leading base j was dropped."""

    def __init__(self, state_len, alphabet, engine=None):
        self.alphabet = alphabet
        self.state_len = state_len
        self.n_base = len(alphabet[1:])
        C = self.n_base ** state_len
        states = torch.arange(C)
        moves = torch.arange(self.n_base)[None, :] * self.n_base ** (state_len - 1) + (states // self.n_base)[:, None]
        self.idx = torch.cat([states[:, None], moves], dim=1).to(torch.int32)
        self.engine = engine if engine is not None else Engine(''.join(alphabet), state_len)

    def n_score(self):
        return len(self.alphabet) * self.n_base ** self.state_len

    def _handle(self, scores):
        T, N, W = scores.shape
        if W != self.n_score():
            raise ValueError('scores last dim %d != n_score() = %d' % (W, self.n_score()))
        return self.engine.get(scores.device, N, T, bf16=bool(self.engine.handle and self.engine.handle.bf16),
                               encoder=False)

    @staticmethod
    def _semiring(S):
        """0 = Log, 1 = Max; accepts this module's markers, seqdist's semiring objects or their names."""
        name = S if isinstance(S, str) else getattr(S, '__name__', getattr(S, 'name', type(S).__name__))
        name = str(name).lower()
        if S is Log or name == 'log':
            return 0
        if S is Max or name == 'max':
            return 1
        raise NotImplementedError('semiring %r: the CUDA scans implement Log and Max' % (S,))

    def logZ(self, scores, S=Log):
        return self._handle(scores).logZ(scores, self._semiring(S))

    def normalise(self, scores):
        return scores - self.logZ(scores)[:, None] / len(scores)

    def forward_scores(self, scores, S=Log):
        return self._handle(scores).forward_scores(scores, self._semiring(S))

    def backward_scores(self, scores, S=Log):
        return self._handle(scores).backward_scores(scores, self._semiring(S))

    def posteriors(self, scores, S=Log):
        """Edge posteriors d(sum logZ_S)/d scores, (T, N, C*NZ) fp32 (seqdist.core.SequenceDist.posteriors): the softmax
        over all edges of a step for Log, the one-hot at the arg-max edge of the max-marginals for Max."""
        if self._semiring(S) == 1:
            return self._handle(scores).posteriors_max(scores)[1]
        return self._handle(scores).posteriors(scores)

    def viterbi(self, scores):
        """Arg-max edge of the max-marginals per step, as edge % NZ: (T, N) int64 labels."""
        labels_nt = self._handle(scores).viterbi(scores)
        return labels_nt.T.to(torch.int64)

    def path_to_str(self, path):
        letters = np.frombuffer(''.join(self.alphabet).encode(), dtype='u1')
        path = np.asarray(path)
        return letters[path[path != 0]].tobytes().decode()

    def beam_search(self, scores, beam_width=32, beam_cut=100.0):
        """Beam-search decode (the role of koi.decode.beam_search, which is ACGT-only): (sequence, qstring, moves, lens),
        sequence / qstring (N, T) int8 left-packed, moves (N, T) bool."""
        return self._handle(scores).beam_search(scores, beam_width, beam_cut)

    def decode_packed(self, scores):
        """decode_batch without the host strings: (seq (N,T) int8 left-packed letters, qstring, lens)."""
        return self._handle(scores).decode(scores)

    def ctc_loss(self, scores, targets, target_lengths, loss_clip=None, reduction='mean', normalise_scores=True):
        loss = _CTCLoss.apply(scores, self._handle(scores), targets, target_lengths, normalise_scores)
        if loss_clip:
            loss = torch.clamp(loss, 0.0, loss_clip)
        if reduction == 'mean':
            return loss.mean()
        if reduction in ('none', None):
            return loss
        raise ValueError('Unknown reduction type {}'.format(reduction))

    def compute_transition_probs(self, scores, betas):
        """Per-state transition probabilities in (old_state, emitted_base) layout and initial state probabilities
        (crf/model.py:63-76; only the reference's unused duplex CLI consumes them).  betas come from the CUDA backward
        scan (backward_scores); the re-layout and the two small softmaxes are torch ops on the same device."""
        T, N, _ = scores.shape
        lt = scores.reshape(T, N, -1, self.n_base + 1) + betas[1:, :, :, None]
        lt = torch.cat([lt[:, :, :, [0]], lt[:, :, :, 1:].transpose(3, 2).reshape(T, N, -1, self.n_base)], dim=-1)
        return torch.softmax(lt, dim=-1), torch.softmax(betas[0], dim=-1)

    # Index permutations of the score tensor (no arithmetic): device-side tensor views, as in the reference.
    def reverse_complement(self, scores):
        T, N, _ = scores.shape
        n, sl = self.n_base, self.state_len
        s = scores.reshape(T, N, *([n] * sl), n + 1)
        blanks = s[..., 0].permute(0, 1, *range(sl + 1, 1, -1)).reshape(T, N, -1, 1).flip([0, 2])
        emis = s[..., 1:].permute(0, 1, *range(sl, 1, -1), sl + 2, sl + 1).reshape(T, N, -1, n).flip([0, 2, 3])
        return torch.cat([blanks, emis], dim=-1).reshape(T, N, -1)


def conv(c_in, c_out, ks, stride=1, bias=False, activation=None):
    return Convolution(c_in, c_out, ks, stride=stride, padding=ks // 2, bias=bias, activation=activation)


def rnn_encoder(n_base, state_len, insize=1, stride=5, winlen=19, activation='swish', rnn_type='lstm', features=768,
                scale=5.0, blank_score=None, expand_blanks=True, extra_linear=False, drop_rate=0, drop_rate_bottom=0):
    rnn = layers[rnn_type]
    drop = (lambda: [torch.nn.Dropout(p=drop_rate_bottom)]) if drop_rate_bottom else (lambda: [])
    stem = [
        conv(insize, 4, ks=5, bias=True, activation=activation), *drop(),
        conv(4, 16, ks=5, bias=True, activation=activation), *drop(),
        conv(16, features, ks=winlen, stride=stride, bias=True, activation=activation), *drop(),
        Permute([2, 0, 1]),
    ]
    stack = []
    for i in range(5):
        stack.append(rnn(features, features, reverse=(i % 2 == 0)))
        if i < 4:
            stack += drop()
    head = LinearCRFEncoder(features, n_base, state_len, activation='tanh', scale=scale, blank_score=blank_score,
                            expand_blanks=expand_blanks, extra_linear=extra_linear, drop_rate=drop_rate)
    return Serial(stem + stack + [head])


class SeqdistModel(Module):
    def __init__(self, encoder, seqdist):
        super().__init__()
        self.seqdist = seqdist
        self.encoder = encoder
        self.stride = get_stride(encoder)
        self.alphabet = seqdist.alphabet
        if isinstance(encoder, Serial):
            encoder.set_engine(seqdist.engine)      # one handle for encoder + decode

    def forward(self, x):
        return self.encoder(x)

    def decode_batch(self, x):
        seq, _, lens = self.seqdist.decode_packed(x)
        seq, lens = seq.cpu().numpy(), lens.cpu().numpy()
        return [seq[i, :lens[i]].astype('u1').tobytes().decode() for i in range(seq.shape[0])]

    def decode(self, x):
        return self.decode_batch(x.unsqueeze(1))[0]


class Model(SeqdistModel):

    def __init__(self, config):
        seqdist = CTC_CRF(state_len=config['global_norm']['state_len'], alphabet=config['labels']['labels'])
        if 'type' in config['encoder']:      # new-style config
            encoder = from_dict(config['encoder'])
        else:                                # old-style
            encoder = rnn_encoder(seqdist.n_base, seqdist.state_len, insize=config['input']['features'],
                                  **config['encoder'])
        super().__init__(encoder, seqdist)
        self.config = config
