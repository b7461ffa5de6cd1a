"""Layer classes of the encoder under the reference's registry names and constructor arguments
(ub-bonito/bonito/nn.py): `layers` / `register` (:10-16), Swish (:22-24), Serial (:27-36), Reverse (:39-54),
Convolution (:57-84), LinearCRFEncoder (:87-153), Permute (:156-167), LSTM / RNNWrapper (:176-235),
to_dict / from_dict (:238-259).

Parameters are ordinary torch Parameters with the reference's names (conv.weight, rnn.weight_ih_l0,
linear.weight ...), so reference checkpoints load unchanged; the arithmetic runs in libxna_b200.so:

  Serial.forward   recognises [Convolution(1,4,5) Convolution(4,16,5) Convolution(16,768,19,s5) Permute(2,0,1)]
                   and runs it as the fused stem (xb_conv_stem_fwd), then LSTM layers (xb_lstm_fwd) and the
                   CRF head (xb_crf_head_fwd); activations stay 16-bit (T, N, 768) on the device in between.
  LSTM.forward     one layer on a (T, N, 768) CUDA tensor; `reverse` walks time backwards by indexing.
  LinearCRFEncoder.forward   scale * tanh(x W^T + b) with the blank score expanded in the epilogue -> fp32.

There is no CPU or eager-PyTorch fallback: tensors must live on an sm_100a device, and shapes outside the
sup@v3.3 architecture (features 768, stem 1->4->16->768) raise.
"""
import torch
from torch.nn import Module
from torch.nn.init import orthogonal_

from .engine import Engine

layers = {}

FEATURES = 768


def register(layer):
    layer.name = layer.__name__.lower()
    layers[layer.name] = layer
    return layer


register(torch.nn.ReLU)
register(torch.nn.Tanh)


@register
class Swish(torch.nn.SiLU):
    pass


def _engine_of(module):
    eng = getattr(module, '_xb_engine', None)
    if eng is None:
        eng = Engine()
        object.__setattr__(module, '_xb_engine', eng)
    return eng


def _adopt(module, engine, slot_counter):
    """Hand `engine` down a module tree and number the LSTM layers (their weight slot in the handle)."""
    object.__setattr__(module, '_xb_engine', engine)
    if isinstance(module, LSTM):
        object.__setattr__(module, '_xb_slot', slot_counter[0])
        slot_counter[0] += 1
    for child in module.children():
        _adopt(child, engine, slot_counter)


def _is_bf16(module):
    p = next(module.parameters(), None)
    return p is not None and p.dtype == torch.bfloat16


@register
class Serial(torch.nn.Sequential):

    def __init__(self, sublayers):
        if isinstance(sublayers, dict):      # slicing a Sequential re-enters here with an OrderedDict:
            super().__init__(sublayers)      # a view of modules that keep their engine and weight slots
            first = next(iter(sublayers.values()), None)
            object.__setattr__(self, '_xb_engine', getattr(first, '_xb_engine', None))
            return
        super().__init__(*sublayers)
        head = next((m for m in self.modules() if isinstance(m, LinearCRFEncoder)), None)
        if head is not None:
            eng = Engine('N' + 'ACGTXYZWVU'[:head.n_base], head.state_len)
        else:
            eng = Engine()
        _adopt(self, eng, [0])

    def set_engine(self, engine):
        _adopt(self, engine, [0])

    def _stem(self):
        """The four leading modules when they are the sup@v3.3 convolution stem, else None."""
        mods = list(self)
        if len(mods) < 4:
            return None
        c1, c2, c3, pm = mods[:4]
        if not (isinstance(c1, Convolution) and isinstance(c2, Convolution) and isinstance(c3, Convolution)
                and isinstance(pm, Permute)):
            return None
        want = [(1, 4, 5, 1, 2), (4, 16, 5, 1, 2), (16, FEATURES, 19, 5, 9)]
        for c, w in zip((c1, c2, c3), want):
            got = (c.conv.in_channels, c.conv.out_channels, c.conv.kernel_size[0], c.conv.stride[0], c.conv.padding[0])
            if got != w or not isinstance(c.activation, Swish) or c.conv.bias is None:
                return None
        if list(pm.dims) != [2, 0, 1]:
            return None
        return c1, c2, c3

    def forward(self, x):
        mods = list(self)
        stem = self._stem()
        start = 0
        if stem is not None and x.dim() == 3 and x.shape[1] == 1:
            x = _conv_stem_forward(self, stem, x)
            start = 4
        for m in mods[start:]:
            if isinstance(m, torch.nn.Dropout):
                if m.training and m.p > 0:
                    raise RuntimeError('dropout is not part of the B200 forward path; call model.eval()')
                continue
            x = m(x)
        return x

    def sync_weights(self, handle=None):
        """Push every parameter of the tree whose storage changed into the engine's handle."""
        eng = _engine_of(self)
        stem = self._stem()
        if stem is not None:
            _sync_stem(eng, stem)
        for m in self.modules():
            if isinstance(m, (LSTM, LinearCRFEncoder)):
                m._sync(eng)

    def to_dict(self, include_weights=False):
        return {'sublayers': [to_dict(layer, include_weights) for layer in self._modules.values()]}


def _sync_stem(eng, stem):
    tensors = [t for c in stem for t in (c.conv.weight, c.conv.bias)]
    eng.sync('conv', tensors, lambda hd: hd.load_conv_weights(*tensors))


def _conv_stem_forward(owner, stem, x):
    eng = _engine_of(owner)
    N, _, L = x.shape
    if L % 5:
        raise ValueError('chunk length %d is not a multiple of the stride 5' % L)
    h = eng.get(x.device, N, L // 5, bf16=_is_bf16(owner))
    _sync_stem(eng, stem)
    return h.conv_stem(x)


@register
class Reverse(Module):

    def __init__(self, sublayers):
        super().__init__()
        self.layer = Serial(sublayers) if isinstance(sublayers, list) else sublayers

    def forward(self, x):
        inner = self.layer
        if isinstance(inner, LSTM):                 # direction by indexing, no flips
            return inner(x, reverse=not inner.reverse)
        return inner(x.flip(0)).flip(0)

    def to_dict(self, include_weights=False):
        if isinstance(self.layer, Serial):
            return self.layer.to_dict(include_weights)
        return {'sublayers': to_dict(self.layer, include_weights)}


@register
class Convolution(Module):

    def __init__(self, insize, size, winlen, stride=1, padding=0, bias=True, activation=None):
        super().__init__()
        self.conv = torch.nn.Conv1d(insize, size, winlen, stride=stride, padding=padding, bias=bias)
        self.activation = layers.get(activation, lambda: activation)()

    def forward(self, x):
        raise RuntimeError(
            'xna_basecaller_b200.nn.Convolution only runs as part of the fused stem '
            'Serial([Convolution(1,4,5), Convolution(4,16,5), Convolution(16,768,19,stride=5), Permute([2,0,1]), ...]); '
            'there is no stand-alone or CPU convolution on this path')

    def to_dict(self, include_weights=False):
        res = {
            'insize': self.conv.in_channels,
            'size': self.conv.out_channels,
            'bias': self.conv.bias is not None,
            'winlen': self.conv.kernel_size[0],
            'stride': self.conv.stride[0],
            'padding': self.conv.padding[0],
            'activation': self.activation.name if self.activation else None,
        }
        if include_weights:
            res['params'] = {'W': self.conv.weight, 'b': self.conv.bias if self.conv.bias is not None else []}
        return res


@register
class LinearCRFEncoder(Module):

    def __init__(self, insize, n_base, state_len, bias=True, scale=None, activation=None, blank_score=None,
                 expand_blanks=True, extra_linear=False, drop_rate=0):
        super().__init__()
        self.scale = scale
        self.n_base = n_base
        self.state_len = state_len
        self.blank_score = blank_score
        self.expand_blanks = expand_blanks
        self.extra_linear = extra_linear
        if extra_linear:
            self.linear_ext = torch.nn.Linear(insize, insize, bias=bias)
        self.dropout = torch.nn.Dropout(p=drop_rate)
        size = (n_base + 1) * n_base ** state_len if blank_score is None else n_base ** (state_len + 1)
        self.linear = torch.nn.Linear(insize, size, bias=bias)
        self.activation = layers.get(activation, lambda: activation)()

    def forward(self, x):
        if self.extra_linear:
            raise RuntimeError('extra_linear heads have no B200 kernel (the sup@v3.3 models do not use them)')
        if self.dropout.training and self.dropout.p > 0:
            raise RuntimeError('dropout is not part of the B200 forward path; call model.eval()')
        if not isinstance(self.activation, torch.nn.Tanh) or self.blank_score is None:
            raise RuntimeError('the B200 CRF head computes scale*tanh(.) with a constant blank score '
                               '(activation="tanh", blank_score set), as in the sup@v3.3 config')
        if self.linear.in_features != FEATURES:
            raise RuntimeError('the B200 CRF head is built for %d input features' % FEATURES)
        eng = _engine_of(self)
        T, N, _ = x.shape
        h = eng.get(x.device, N, T, bf16=_is_bf16(self))
        self._sync(eng)
        return h.crf_head(x.to(h.dtype16), expand_blanks=self.expand_blanks)

    def _sync(self, eng):
        h = eng.handle
        if h.n_base != self.n_base or h.state_len != self.state_len:
            raise RuntimeError('engine alphabet (n_base %d, state_len %d) does not match the head (%d, %d)'
                               % (h.n_base, h.state_len, self.n_base, self.state_len))
        tensors = [self.linear.weight, self.linear.bias]
        key = (self.scale, self.blank_score, self.expand_blanks)
        eng.sync(('head', key), tensors, lambda hd: hd.load_head_weights(
            self.linear.weight, self.linear.bias, self.scale, self.blank_score, self.expand_blanks))

    def to_dict(self, include_weights=False):
        res = {
            'insize': self.linear.in_features,
            'n_base': self.n_base,
            'state_len': self.state_len,
            'bias': self.linear.bias is not None,
            'scale': self.scale,
            'activation': self.activation.name if self.activation else None,
            'blank_score': self.blank_score,
        }
        if include_weights:
            res['params'] = {'W': self.linear.weight, 'b': self.linear.bias if self.linear.bias is not None else []}
            if self.extra_linear:
                res['params']['W_ext'] = self.linear_ext.weight
                res['params']['b_ext'] = self.linear_ext.bias if self.linear_ext.bias is not None else []
        return res


@register
class Permute(Module):

    def __init__(self, dims):
        super().__init__()
        self.dims = dims

    def forward(self, x):
        return x.permute(*self.dims)

    def to_dict(self, include_weights=False):
        return {'dims': self.dims}


def truncated_normal(size, dtype=torch.float32, device=None, num_resample=5):
    x = torch.empty(size + (num_resample,), dtype=torch.float32, device=device).normal_()
    i = ((x < 2) & (x > -2)).max(-1, keepdim=True)[1]
    return torch.clamp_(x.gather(-1, i).squeeze(-1), -2, 2)


class RNNWrapper(Module):
    """Holds a torch.nn RNN module for its parameters and initialisation (orthogonal weights, truncated-normal
    bias_ih, frozen zero bias_hh: nn.py:195-213); forward goes to the persistent tcgen05 kernel."""

    def __init__(self, rnn_type, *args, reverse=False, orthogonal_weight_init=True, disable_state_bias=True,
                 bidirectional=False, **kwargs):
        super().__init__()
        if reverse and bidirectional:
            raise Exception("'reverse' and 'bidirectional' should not both be set to True")
        self.reverse = reverse
        self.rnn = rnn_type(*args, bidirectional=bidirectional, **kwargs)
        self.init_orthogonal(orthogonal_weight_init)
        self.init_biases()
        if disable_state_bias:
            self.disable_state_bias()

    def forward(self, x, reverse=None):
        rnn = self.rnn
        if not isinstance(rnn, torch.nn.LSTM) or rnn.bidirectional or rnn.num_layers != 1 \
                or rnn.hidden_size != FEATURES or rnn.input_size != FEATURES:
            raise RuntimeError('the B200 recurrent kernel is a single-layer unidirectional LSTM(768, 768)')
        eng = _engine_of(self)
        T, N, _ = x.shape
        h = eng.get(x.device, N, T, bf16=_is_bf16(self))
        self._sync(eng)
        return h.lstm(getattr(self, '_xb_slot', 0), x.to(h.dtype16), self.reverse if reverse is None else reverse)

    def _sync(self, eng):
        rnn = self.rnn
        slot = getattr(self, '_xb_slot', 0)
        if slot >= 5:
            raise RuntimeError('the engine holds five LSTM weight slots; layer %d does not fit' % slot)
        w = [rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0]
        eng.sync(('lstm', slot), w, lambda hd: hd.load_lstm_weights(slot, *w))

    def init_biases(self, types=('bias_ih',)):
        for name, param in self.rnn.named_parameters():
            if any(k in name for k in types):
                with torch.no_grad():
                    param.set_(0.5 * truncated_normal(param.shape, dtype=param.dtype, device=param.device))

    def init_orthogonal(self, types=True):
        if not types:
            return
        if types is True:
            types = ('weight_ih', 'weight_hh')
        for name, x in self.rnn.named_parameters():
            if any(k in name for k in types):
                for i in range(0, x.size(0), self.rnn.hidden_size):
                    orthogonal_(x[i:i + self.rnn.hidden_size])

    def disable_state_bias(self):
        for name, x in self.rnn.named_parameters():
            if 'bias_hh' in name:
                x.requires_grad = False
                x.zero_()


@register
class LSTM(RNNWrapper):

    def __init__(self, size, insize, bias=True, reverse=False):
        super().__init__(torch.nn.LSTM, size, insize, bias=bias, reverse=reverse)

    def to_dict(self, include_weights=False):
        res = {
            'size': self.rnn.hidden_size,
            'insize': self.rnn.input_size,
            'bias': self.rnn.bias,
            'reverse': self.reverse,
        }
        if include_weights:
            H, I = self.rnn.hidden_size, self.rnn.input_size
            res['params'] = {
                'iW': self.rnn.weight_ih_l0.reshape(4, H, I),
                'sW': self.rnn.weight_hh_l0.reshape(4, H, H),
                'b': self.rnn.bias_ih_l0.reshape(4, H),
            }
        return res


def to_dict(layer, include_weights=False):
    if hasattr(layer, 'to_dict'):
        return {'type': layer.name, **layer.to_dict(include_weights)}
    return {'type': layer.name}


def from_dict(model_dict, layer_types=None):
    model_dict = model_dict.copy()
    if layer_types is None:
        layer_types = layers
    type_name = model_dict.pop('type')
    typ = layer_types[type_name]
    if 'sublayers' in model_dict:
        sub = model_dict['sublayers']
        model_dict['sublayers'] = [from_dict(x, layer_types) for x in sub] if isinstance(sub, list) \
            else from_dict(sub, layer_types)
    try:
        return typ(**model_dict)
    except Exception as e:
        raise Exception(f'Failed to build layer of type {typ} with args {model_dict}') from e
