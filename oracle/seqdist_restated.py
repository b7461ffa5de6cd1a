"""Restatement (torch, CPU, any float dtype) of the ``ont-seqdist==0.0.4`` calls the reference makes.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

seqdist is a third-party dependency pinned at ``ont-seqdist-cuda102==0.0.4``
(/root/reference/ub-bonito/requirements.txt:18); its source is NOT in /root/reference and its
kernels are cupy RawKernels (GPU only), so it can be neither vendored nor run here.  What follows
restates its published algorithm from the reference's call sites:

  bonito/crf/model.py:9-11    imports  seqdist.sparse, seqdist.ctc_simple.{logZ_cupy,viterbi_alignments},
                              seqdist.core.{SequenceDist, Max, Log, semiring}
  bonito/crf/model.py:41-46   seqdist.sparse.logZ(Ms, idx, alpha_0, beta_T, S)
  bonito/crf/model.py:51-61   seqdist.sparse.fwd_scores_cupy / bwd_scores_cupy (K=1)
  bonito/crf/model.py:92-95   SequenceDist.posteriors(scores, Max)
  bonito/crf/model.py:122,135 ctc_simple.logZ_cupy / viterbi_alignments
  bonito/crf/model.py:216     SequenceDist.posteriors(scores)  (Log)

Algorithm (sparse transition lattice):  Ms[t, n, c, k] is the score of the k-th incoming edge of
state c at step t, idx[c, k] the state that edge leaves.  With a semiring S = (zero, one, mul, sum):

    alpha_0 = v0,            alpha_{t+1}[c] = S.sum_k  S.mul(Ms[t, c, k], alpha_t[idx[c, k]])
    beta_T  = vT,            beta_t[c']     = S.sum_{(c,k): idx[c,k]=c'} S.mul(Ms[t, c, k], beta_{t+1}[c])
    logZ[n] = S.sum_c S.mul(alpha_T[c], vT[c])
    d logZ / d Ms[t] = S.dsum over ALL (c, k) of  S.mul(S.mul(Ms[t,c,k], alpha_t[idx[c,k]]), beta_{t+1}[c])

``Log``:  mul = +, sum = logsumexp, dsum = softmax, zero = -1e38, one = 0.
``Max``:  mul = +, sum = max,       dsum = one-hot at the (first) arg-max.
``posteriors(scores, S)`` is autograd of ``logZ(scores, S).sum()`` with respect to ``scores``.

PARITY UNPINNED against seqdist itself (no source, no wheel, no GPU here).  Known open point:
seqdist's kernels are launched with C//K threads (K=4 by default); how 0.0.4 treats C=125 (not a
multiple of 4) cannot be verified here.  This restatement uses the mathematically complete state set.
"""
from collections import namedtuple

import torch

semiring = namedtuple('semiring', ('zero', 'one', 'mul', 'sum', 'dsum'))


def _one_hot_argmax(x, dim=0):
    # first index on ties (torch CPU argmax returns the first maximal element)
    return torch.zeros_like(x).scatter_(dim, x.argmax(dim, True), 1.0)


Log = semiring(zero=-1e38, one=0., mul=torch.add, sum=torch.logsumexp, dsum=torch.softmax)
Max = semiring(zero=-1e38, one=0., mul=torch.add,
               sum=(lambda x, dim=0: torch.max(x, dim=dim)[0]), dsum=_one_hot_argmax)


def _source_lists(idx):
    """For every source state c' the flat edge ids (c*NZ+k) with idx[c,k]==c'.

    Every state of the CTC-CRF lattice has the same out-degree (NZ), so a stable argsort of
    the flattened table reshapes to (C, NZ)."""
    C, NZ = idx.shape
    order = idx.flatten().to(torch.int64).argsort(stable=True)
    return order.reshape(C, NZ)


def fwd_scores(Ms, idx, v0, S=Log):
    """alpha (T+1, N, C)."""
    T, N, C, NZ = Ms.shape
    idx = idx.to(torch.int64)
    alpha = Ms.new_full((T + 1, N, C), S.zero)
    alpha[0] = v0
    for t in range(T):
        alpha[t + 1] = S.sum(S.mul(Ms[t], alpha[t][:, idx]), dim=-1)
    return alpha


def bwd_scores(Ms, idx, vT, S=Log):
    """beta (T+1, N, C)."""
    T, N, C, NZ = Ms.shape
    src = _source_lists(idx)                      # (C, NZ) flat edge ids leaving each state
    dst_state = src // NZ                         # the state each of those edges enters
    beta = Ms.new_full((T + 1, N, C), S.zero)
    beta[T] = vT
    Mflat = Ms.reshape(T, N, C * NZ)
    for t in range(T - 1, -1, -1):
        beta[t] = S.sum(S.mul(Mflat[t][:, src], beta[t + 1][:, dst_state]), dim=-1)
    return beta


class _SparseLogZ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Ms, idx, v0, vT, S):
        alpha = fwd_scores(Ms, idx, v0, S)
        ctx.save_for_backward(Ms, idx, vT, alpha)
        ctx.S = S
        return S.sum(S.mul(alpha[-1], vT), dim=1)

    @staticmethod
    def backward(ctx, grad):
        Ms, idx, vT, alpha = ctx.saved_tensors
        S = ctx.S
        T, N, C, NZ = Ms.shape
        beta = bwd_scores(Ms, idx, vT, S)
        edge = S.mul(S.mul(Ms, alpha[:-1][:, :, idx.to(torch.int64)]), beta[1:, :, :, None])
        edge = S.dsum(edge.reshape(T, N, -1), dim=2).reshape(T, N, C, NZ)
        return grad[None, :, None, None] * edge, None, None, None, None


class sparse:
    """Stand-in for the ``seqdist.sparse`` module namespace."""

    @staticmethod
    def logZ(Ms, idx, v0, vT, S=Log, K=4):
        return _SparseLogZ.apply(Ms, idx, v0, vT, S)

    @staticmethod
    def fwd_scores_cupy(Ms, idx, v0, S=Log, K=4):
        return fwd_scores(Ms, idx, v0, S)

    @staticmethod
    def bwd_scores_cupy(Ms, idx, vT, S=Log, K=4):
        return bwd_scores(Ms, idx, vT, S)


class SequenceDist:
    """``seqdist.core.SequenceDist``: only what bonito/crf/model.py inherits and uses."""

    def __init__(self):
        pass

    def logZ(self, scores, S=Log):
        raise NotImplementedError

    def viterbi(self, scores):
        raise NotImplementedError

    def ctc_loss(self, scores, targets, target_lengths):
        raise NotImplementedError

    def posteriors(self, scores, S=Log):
        with torch.enable_grad():
            x = scores.detach().requires_grad_(True)
            total = self.logZ(x, S).sum()
            (g,) = torch.autograd.grad(total, x)
        return g


# --------------------------------------------------------------------------- ctc_simple
def _simple_alpha(stay, move, S):
    """alpha (T+1, N, L) of the stay/move lattice; paths start at position 0."""
    T, N, L = stay.shape
    alpha = stay.new_full((T + 1, N, L), S.zero)
    alpha[0, :, 0] = S.one
    for t in range(T):
        a = alpha[t]
        stayed = S.mul(stay[t], a)
        moved = torch.cat([a.new_full((N, 1), S.zero), S.mul(move[t], a[:, :-1])], dim=1)
        alpha[t + 1] = S.sum(torch.stack([stayed, moved], dim=-1), dim=-1)
    return alpha


class _SimpleLogZ(torch.autograd.Function):
    """log-sum over monotone alignments (stay or move by one at every step), ending at
    position target_length-1 after the last step.  Gradients by the beta recursion."""

    @staticmethod
    def forward(ctx, stay, move, target_lengths, S):
        T, N, L = stay.shape
        alpha = _simple_alpha(stay, move, S)
        beta_T = stay.new_full((N, L), S.zero)
        beta_T[torch.arange(N), target_lengths.to(torch.int64) - 1] = S.one
        ctx.save_for_backward(stay, move, alpha, beta_T)
        ctx.S = S
        return S.sum(S.mul(alpha[-1], beta_T), dim=1)

    @staticmethod
    def backward(ctx, grad):
        stay, move, alpha, beta_T = ctx.saved_tensors
        S = ctx.S
        T, N, L = stay.shape
        beta = stay.new_full((T + 1, N, L), S.zero)
        beta[T] = beta_T
        for t in range(T - 1, -1, -1):
            b = beta[t + 1]
            stayed = S.mul(stay[t], b)
            moved = torch.cat([S.mul(move[t], b[:, 1:]), b.new_full((N, 1), S.zero)], dim=1)
            beta[t] = S.sum(torch.stack([stayed, moved], dim=-1), dim=-1)
        g_stay = S.mul(S.mul(alpha[:-1], stay), beta[1:])
        g_move = S.mul(S.mul(alpha[:-1, :, :-1], move), beta[1:, :, 1:])
        g = S.dsum(torch.cat([g_stay, g_move], dim=2), dim=2)
        g = g * grad[None, :, None]
        return g[:, :, :L], g[:, :, L:], None, None


class ctc_simple:
    """Stand-in for the ``seqdist.ctc_simple`` module namespace."""

    @staticmethod
    def logZ_cupy(stay_scores, move_scores, target_lengths, S=Log):
        return _SimpleLogZ.apply(stay_scores, move_scores, target_lengths, S)

    @staticmethod
    def viterbi_alignments(stay_scores, move_scores, target_lengths):
        with torch.enable_grad():
            s = stay_scores.detach().requires_grad_(True)
            m = move_scores.detach().requires_grad_(True)
            z = _SimpleLogZ.apply(s, m, target_lengths, Max).sum()
            gs, gm = torch.autograd.grad(z, (s, m))
        return gs, gm
