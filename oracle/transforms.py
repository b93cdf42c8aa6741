"""Oracle (test infrastructure): CPU restatement of reference src/model_transforms.py.

PARITY UNPINNED (see oracle/__init__.py).  Everything here is torch-CPU; nothing in the product
imports it.  Tensors are channels_first (N, C, D, H, W) like the reference's models
(src/model_types.py:180); kernels are kept in the Keras layouts:

  Conv3D.kernel           (kd, kh, kw, C_in,  C_out)
  Conv3DTranspose.kernel  (kd, kh, kw, C_out, C_in)

A transform is described by a small spec tree:
  ('conv',  dict(filters, k, s, bias, relu))       Keras Conv3D, padding='same'
  ('convT', dict(filters, k, s, bias, relu))       Keras Conv3DTranspose, padding='same'
  ('seq',   [children])                            SequentialLayer   model_transforms.py:11-19
  ('res',   [children], mode)                      ResidualLayer     model_transforms.py:22-38
and its weights are a flat list of {'kernel', 'bias'} in layer-creation order.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def same_pads(n, k, s):
    """TensorFlow 'SAME' padding of a forward conv on size n: (out, pad_before, pad_after)."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv3d_same(x, kernel, bias, stride, relu):
    """Keras Conv3D(padding='same'): cross-correlation, asymmetric SAME padding (more at the end)."""
    k = kernel.shape[0]
    pads = []
    for n in reversed(x.shape[2:]):  # F.pad takes the last dim first
        _, pb, pa = same_pads(n, k, stride)
        pads += [pb, pa]
    w = kernel.permute(4, 3, 0, 1, 2).contiguous()  # -> (C_out, C_in, kd, kh, kw)
    y = F.conv3d(F.pad(x, pads), w, bias, stride=stride)
    return torch.relu(y) if relu else y


def conv3d_transpose_same(x, kernel, bias, stride, relu):
    """Keras Conv3DTranspose(padding='same', output_padding=None): the adjoint of the forward SAME conv
    whose input has size in*stride.  Full transposed conv, then crop pad_before at the front."""
    k = kernel.shape[0]
    w = kernel.permute(4, 3, 0, 1, 2).contiguous()  # (kd,kh,kw,C_out,C_in) -> (C_in, C_out, kd, kh, kw)
    y = F.conv_transpose3d(x, w, None, stride=stride)
    sl = [slice(None), slice(None)]
    for n in x.shape[2:]:
        _, pb, _ = same_pads(n * stride, k, stride)
        sl.append(slice(pb, pb + n * stride))
    y = y[tuple(sl)]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    return torch.relu(y) if relu else y


# ---------------------------------------------------------------------------------------------
# spec builders, one per reference class
# ---------------------------------------------------------------------------------------------
def _conv(filters, k=3, s=1, bias=True, relu=True):
    return ('conv', dict(filters=filters, k=k, s=s, bias=bias, relu=relu))


def _convT(filters, k=3, s=1, bias=True, relu=True):
    return ('convT', dict(filters=filters, k=k, s=s, bias=bias, relu=relu))


def analysis_block(filters, mode='add'):  # model_transforms.py:62-70
    return ('res', [_conv(filters, s=2), _conv(filters), _conv(filters)], mode)


def synthesis_block(filters, mode='add'):  # model_transforms.py:73-81
    return ('res', [_convT(filters, s=2), _convT(filters), _convT(filters)], mode)


def build_transform(name, filters, residual_mode='add'):
    f, m = filters, residual_mode
    if name == 'AnalysisTransformV1':  # :41-48
        return ('seq', [_conv(f, 9, 2), _conv(f, 5, 2), _conv(f, 5, 2, bias=False, relu=False)])
    if name == 'SynthesisTransformV1':  # :51-59
        return ('seq', [_convT(f, 5, 2), _convT(f, 5, 2), _convT(1, 9, 2)])
    if name == 'AnalysisTransformV2':  # :84-95
        return ('seq', [analysis_block(f // 2, m), analysis_block(f, m), analysis_block(f, m),
                        _conv(f, bias=False, relu=False)])
    if name == 'SynthesisTransformV2':  # :98-109
        return ('seq', [synthesis_block(f, m), synthesis_block(f, m), synthesis_block(f // 2, m), _convT(1)])
    if name == 'AnalysisTransformProgressiveV2':  # :112-123
        return ('seq', [analysis_block(f // 4, m), analysis_block(f // 2, m), analysis_block(f, m),
                        _conv(f, bias=False, relu=False)])
    if name == 'SynthesisTransformProgressiveV2':  # :126-137
        return ('seq', [synthesis_block(f, m), synthesis_block(f // 2, m), synthesis_block(f // 4, m), _convT(1)])
    if name == 'HyperAnalysisTransform':  # :140-147
        return ('seq', [_conv(f), _conv(f, s=2), _conv(f, bias=False, relu=False)])
    if name == 'HyperSynthesisTransform':  # :150-158
        return ('seq', [_convT(f), _convT(f, s=2), _convT(f)])
    if name == 'AnalysisBlock':
        return analysis_block(f, m)
    if name == 'SynthesisBlock':
        return synthesis_block(f, m)
    raise KeyError(name)


def _walk(spec, x, weights, it):
    kind = spec[0]
    if kind in ('conv', 'convT'):
        p = spec[1]
        w = weights[next(it)]
        fn = conv3d_same if kind == 'conv' else conv3d_transpose_same
        return fn(x, w['kernel'], w.get('bias'), p['s'], p['relu'])
    if kind == 'seq':
        for child in spec[1]:
            x = _walk(child, x, weights, it)
        return x
    if kind == 'res':  # model_transforms.py:30-38
        x = _walk(spec[1][0], x, weights, it)
        t1 = x
        for child in spec[1][1:]:
            x = _walk(child, x, weights, it)
        return t1 + x if spec[2] == 'add' else torch.cat((x, t1), 1)
    raise ValueError(kind)


def apply_transform(spec, weights, x):
    """x: (N,C,D,H,W) torch tensor; weights: flat list of {'kernel': tensor, 'bias': tensor|None}."""
    return _walk(spec, x, weights, iter(range(len(weights))))


def leaf_layers(spec):
    if spec[0] in ('conv', 'convT'):
        return [spec]
    return [l for child in spec[1] for l in leaf_layers(child)]


def init_weights(spec, in_channels, rng, bias_scale=0.0, dtype=torch.float32):
    """Keras defaults: Glorot-uniform kernels, zero biases (bias_scale>0 draws small random biases so
    tests exercise the bias path).  Tracks channel counts through residual concat."""
    out = []

    def rec(s, c):
        kind = s[0]
        if kind in ('conv', 'convT'):
            p = s[1]
            k, f = p['k'], p['filters']
            limit = math.sqrt(6.0 / (k ** 3 * (c + f)))
            shape = (k, k, k, c, f) if kind == 'conv' else (k, k, k, f, c)
            kern = torch.from_numpy(rng.uniform(-limit, limit, size=shape)).to(dtype)
            b = None
            if p['bias']:
                b = torch.from_numpy(rng.uniform(-bias_scale, bias_scale, size=(f,))).to(dtype)
            out.append({'kernel': kern, 'bias': b})
            return f
        if kind == 'seq':
            for ch in s[1]:
                c = rec(ch, c)
            return c
        c1 = rec(s[1][0], c)
        c2 = c1
        for ch in s[1][1:]:
            c2 = rec(ch, c2)
        return c1 if s[2] == 'add' else c1 + c2

    rec(spec, in_channels)
    return out


def weights_to(weights, dtype):
    return [{'kernel': torch.as_tensor(np.asarray(w['kernel'])).to(dtype),
             'bias': None if w.get('bias') is None else torch.as_tensor(np.asarray(w['bias'])).to(dtype)}
            for w in weights]
