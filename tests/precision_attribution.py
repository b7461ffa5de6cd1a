"""Where does the 16-bit score error come from?  CPU emulation of the CUDA encoder's rounding points on the fp32 oracle
(test infrastructure; run by hand: `python tests/precision_attribution.py [L] [N]`).

The CUDA encoder rounds to 16 bits at: the weights (conv3, W_ih, W_hh, head), the stem output, the hoisted input
projection G = x W_ih^T + b (stored between the two LSTM kernels), and h_t (the MMA operand of the next step and the
next layer's input).  Accumulation, cell state and scores are fp32.  Each variant below rounds a subset of those points
to bf16 / fp16 and reports the max abs score error against the all-fp32 oracle at reference-scale weights.
"""
import sys

import torch

sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'golden'))
from make_golden import REF_SCALE, synthetic_signal  # noqa: E402
from oracle import bonito_oracle as bo  # noqa: E402


def q(x, dt):
    return x if dt is None else x.to(dt).float()


def encoder(sd, x, n_base, w_dt, act_dt, g_dt, h_dt):
    sdq = dict(sd)
    for k in sd:
        if k.endswith('weight') and (k.startswith('encoder.2') or 'rnn' in k or k.startswith('encoder.9')):
            sdq[k] = q(sd[k], w_dt)
    y = q(bo.conv_stem(sdq, x).permute(2, 0, 1).contiguous(), act_dt)
    for layer, rev in zip(range(4, 9), bo.LSTM_DIRECTIONS):
        p = 'encoder.%d.rnn.' % layer
        w_ih, w_hh = sdq[p + 'weight_ih_l0'], sdq[p + 'weight_hh_l0']
        b = sd[p + 'bias_ih_l0'] + sd[p + 'bias_hh_l0']
        T, N, H = y.shape
        gin = q(y @ w_ih.t() + b, g_dt)
        h = y.new_zeros(N, H)
        c = y.new_zeros(N, H)
        out = y.new_empty(T, N, H)
        for t in (range(T - 1, -1, -1) if rev else range(T)):
            g = gin[t] + h @ w_hh.t()
            i, f, gg, o = g.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = q(torch.sigmoid(o) * torch.tanh(c), h_dt)
            out[t] = h
        y = out
    return bo.crf_head(sdq, y, n_base)


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.set_num_threads(8)
    sd = bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE)
    x = synthetic_signal(23, N, L)
    ref = bo.encoder_forward(sd, x, 5)
    bf, hf = torch.bfloat16, torch.float16
    variants = [
        ('fp16 everywhere (product default)', hf, hf, hf, hf),
        ('bf16 everywhere', bf, bf, bf, bf),
        ('bf16, G kept in fp16', bf, bf, hf, bf),
        ('bf16, G in fp32', bf, bf, None, bf),
        ('bf16 weights only', bf, None, None, None),
        ('bf16 activations h only', None, None, None, bf),
        ('bf16 G only', None, None, bf, None),
        ('bf16 weights, fp16 activations + G', bf, hf, hf, hf),
    ]
    for name, w, a, g, h in variants:
        err = (encoder(sd, x, 5, w, a, g, h) - ref).abs()
        print('%-40s max |score err| %.4f   mean %.5f' % (name, err.max().item(), err.mean().item()))


if __name__ == '__main__':
    main()
