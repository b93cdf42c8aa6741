"""Host mirror of the reference's transform layer (src/model_transforms.py:7-169): same class names,
constructor arguments and call convention, executed by libpccgeo's CUDA kernels.

    TransformType.<Name>.value(filters, data_format=..., kernel_size=(3,3,3), activation=relu,
                               residual_mode='add'|'concat')(tensor) -> tensor

`tensor` is a torch CUDA fp32 tensor, (N,C,D,H,W) for channels_first or (N,D,H,W,C) for channels_last.
Layers are lazily built on first call (Keras semantics): Glorot-uniform kernels, zero biases, drawn from a
numpy Generator seeded by `set_seed` (default 42, like the reference's scripts, compress_octree.py:26-27).

Execution: a transform is traced into a flat list of conv steps (residual adds fused into the last conv of
their block).  Each step runs either on the fp32 CUDA-core kernel or, for 3x3x3 stride-1 layers with 16/32
channels, on the tcgen05 tensor-core kernel in the blocked bf16 layout; layout conversions are inserted only
where the representation changes.  There is no CPU path.
"""
import math
from enum import Enum

import numpy as np
import torch

from . import ops

_seed_state = {'rng': np.random.default_rng(42)}

# numerical mode of the tensor-core path: 'bf16x3' (hi/lo split, fp32-class accuracy; default),
# 'bf16' (single bf16 term) or 'fp32' (CUDA-core kernels only).
_precision = {'mode': 'bf16x3'}


def set_seed(seed):
    _seed_state['rng'] = np.random.default_rng(seed)


def set_precision(mode):
    assert mode in ('bf16x3', 'bf16', 'fp32')
    _precision['mode'] = mode


def get_precision():
    return _precision['mode']


def relu(x):  # stand-in for tf.nn.relu as the `activation` argument
    return torch.relu(x)


def get_channel_axis(data_format):  # model_transforms.py:7-8
    return 1 if data_format == 'channels_first' else -1


def normalize_data_format(value):  # keras conv_utils.normalize_data_format
    if value is None:
        value = 'channels_last'  # Keras default image_data_format
    value = str(value).lower()
    if value not in ('channels_first', 'channels_last'):
        raise ValueError(f'The `data_format` argument must be one of "channels_first", "channels_last". Received: {value}')
    return value


def _is_relu(activation):
    if activation is None:
        return False
    if activation is relu or activation == 'relu' or getattr(activation, '__name__', '') == 'relu':
        return True
    raise ValueError('only relu / None activations are supported by the CUDA kernels')


def _triple(v):
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(v)
    assert len(v) == 3
    return v


class Layer:
    def __init__(self, *args, **kwargs):
        self.name = kwargs.get('name')

    def leaf_layers(self):
        return []


# The y-stacked 16-channel kernel (conv3d_umma_ys.cu) keeps the tensor pipe busier (57 % vs 42 % active) but its heavier
# epilogue (row shuffles) makes it finish in the same time as the z-stacked kernel on B200 (0.373 vs 0.368 ms):
# parity-tested, off by default.
_ys_enabled = [False]

# The hi/lo-stacked 16-channel kernel (conv3d_umma.cu, HL mode): two N=96 MMAs per tap instead of three N=48 ones.
_hl_enabled = [True]

# The zy-ring 16-channel kernel (conv3d_umma_zy.cu): y and z taps accumulated in a 2-D TMEM ring, N = 144 MMAs.
_zy_enabled = [True]

# bumped whenever any layer's parameters change: captured CUDA graphs (model_types) hold device pointers of packed weights
params_epoch = [0]


class _ConvBase(Layer):
    transposed = False

    def __init__(self, filters, kernel_size, strides=(1, 1, 1), padding='valid', data_format=None, use_bias=True,
                 activation=None, **kwargs):
        super().__init__(**kwargs)
        ks, st = _triple(kernel_size), _triple(strides)
        if len(set(ks)) != 1 or len(set(st)) != 1:
            raise ValueError('only cubic kernels / isotropic strides are supported')
        if padding != 'same':
            raise ValueError("only padding='same' is supported (all the reference uses)")
        self.filters, self.k, self.stride = int(filters), ks[0], st[0]
        self.use_bias, self.relu = bool(use_bias), _is_relu(activation)
        self.data_format = normalize_data_format(data_format)
        self.kernel = None  # numpy, Keras layout
        self.bias = None
        self._dev = {}

    def leaf_layers(self):
        return [self]

    # -- parameters ----------------------------------------------------------------------------------
    def build(self, in_channels):
        if self.kernel is not None:
            return
        k, f, c = self.k, self.filters, int(in_channels)
        limit = math.sqrt(6.0 / (k ** 3 * (c + f)))
        shape = (k, k, k, f, c) if self.transposed else (k, k, k, c, f)
        self.kernel = _seed_state['rng'].uniform(-limit, limit, size=shape).astype(np.float32)
        self.bias = np.zeros((f,), np.float32) if self.use_bias else None
        self._dev = {}
        params_epoch[0] += 1

    def set_weights(self, kernel, bias=None):
        kernel = np.asarray(kernel, np.float32)
        assert kernel.ndim == 5 and kernel.shape[0] == self.k
        assert kernel.shape[3 if self.transposed else 4] == self.filters
        self.kernel = np.ascontiguousarray(kernel)
        if self.use_bias:
            assert bias is not None and np.asarray(bias).shape == (self.filters,)
            self.bias = np.ascontiguousarray(np.asarray(bias, np.float32))
        else:
            self.bias = None
        self._dev = {}
        params_epoch[0] += 1

    def get_weights(self):
        return {'kernel': self.kernel, 'bias': self.bias}

    @property
    def in_channels(self):
        return self.kernel.shape[4 if self.transposed else 3]

    def tap_major(self):
        """(k^3, Cin, Cout) fp32 numpy: the layout the kernels consume."""
        w = self.kernel.transpose(0, 1, 2, 4, 3) if self.transposed else self.kernel
        return np.ascontiguousarray(w.reshape(self.k ** 3, self.in_channels, self.filters))

    def dev(self, key):
        """Device-resident derived parameters (cached)."""
        if key not in self._dev:
            if key == 'w_tap':
                self._dev[key] = ops.upload(self.tap_major())
            elif key == 'bias':
                self._dev[key] = None if self.bias is None else ops.upload(self.bias)
            elif key.startswith('w_gemm'):
                terms = int(key[-1])
                self._dev[key] = ops.gemm_pack_weights(self.tap_major(), self.in_channels, self.filters, self.k, self.stride,
                                                       self.transposed, terms)
            elif key.startswith('w_ummays'):
                terms = int(key[-1])
                self._dev[key] = ops.umma_ys_pack_weights(self.tap_major(), self.in_channels, self.filters, self.transposed, terms)
            elif key.startswith('w_ummazy'):
                terms = int(key[-1])
                self._dev[key] = ops.umma_zy_pack_weights(self.tap_major(), self.in_channels, self.filters, self.transposed, terms)
            elif key == 'w_ummahl':
                self._dev[key] = ops.umma_hl_pack_weights(self.tap_major(), self.in_channels, self.filters, self.transposed)
            elif key.startswith('w_out1'):
                terms = int(key[-1])
                self._dev[key] = ops.out1_pack_weights(self.tap_major(), self.in_channels, self.transposed, terms)
            elif key.startswith('w_umma'):
                terms = int(key[-1])
                self._dev[key] = ops.umma_pack_weights(self.tap_major(), self.in_channels, self.filters, self.stride,
                                                       self.transposed, terms)
            else:
                raise KeyError(key)
        return self._dev[key]

    def umma_eligible(self, in_shape):
        """TMA halo-plane kernel: 3x3x3 stride-1 layers with <=32-wide channels, and stride-2 transposed layers with
        <=32 input / <=16 output channels, on volumes the 16x8 row tile divides."""
        n, c, d, h, w = in_shape
        cp, fp = ops.round_up(c, 16), ops.round_up(self.filters, 16)
        tiled = h % 16 == 0 and w % 8 == 0 and h * w >= 256
        if self.k == 3 and self.stride == 2 and self.transposed:
            return c >= 8 and cp <= 32 and fp <= 16 and tiled
        return self.k == 3 and self.stride == 1 and c >= 8 and cp <= 32 and fp <= 32 and tiled

    def first_eligible(self, in_shape, terms, src_is_f32):
        """fused first layer: one input channel, 3x3x3, stride 2, <= 16 filters, even dims, fp32 input -> blocked output"""
        n, c, d, h, w = in_shape
        return (terms and src_is_f32 and c == 1 and self.k == 3 and self.stride == 2 and not self.transposed and self.filters <= 16
                and d % 2 == 0 and h % 2 == 0 and w % 2 == 0)

    def zy_eligible(self, in_shape, terms):
        """zy-ring kernel: stride-1 3x3x3 layers with 9..16 channels in and out, batches that fill its M tiles (16 blocks x 8 x-voxels,
        or 8 x 16 / 4 x 32 for batches of 8 / 4 blocks -- the 128^3 blocks run at batch 4) and volumes large enough for its work items
        (one M tile x <= 10 rows) to occupy the GPU."""
        n, c, d, h, w = in_shape
        xch = 1 if n % 16 == 0 else 2 if n % 8 == 0 else 4
        return (_zy_enabled[0] and terms and self.k == 3 and self.stride == 1 and 8 <= c <= 16 and 8 < self.filters <= 16
                and n % (16 // xch) == 0 and w % (8 * xch) == 0 and (n * w // 128) * h >= 148 * 4)

    def hl_eligible(self, in_shape, terms):
        """hi/lo-stacked form of the TMA kernel: two-term precision, stride 1, <= 16 channels in and out."""
        n, c, d, h, w = in_shape
        return (_hl_enabled[0] and terms == 2 and self.k == 3 and self.stride == 1 and 8 <= c <= 16 and self.filters <= 16
                and self.umma_eligible(in_shape))

    def ys_eligible(self, in_shape):
        """y-stacked TMA kernel: 3x3x3 stride-1 layers with 9..16 input and output channels on volumes tall enough for its
        14-row tiles to pay (H >= 28)."""
        n, c, d, h, w = in_shape
        return (self.k == 3 and self.stride == 1 and 8 <= c <= 16 and 8 < self.filters <= 16 and w % 8 == 0 and h >= 28
                and _ys_enabled[0])

    def out1_eligible(self, in_shape):
        """Scatter-form single-output-channel kernel (last layer of the V2 synthesis transforms), fusable with the
        clip / threshold / bit-pack of the block loops."""
        n, c, d, h, w = in_shape
        return (self.k == 3 and self.stride == 1 and self.filters == 1 and 8 <= c <= 16 and h % 16 == 0 and w % 8 == 0
                and h * w >= 256)

    def gemm_eligible(self, in_shape):
        """gather -> tcgen05 kernel: everything else with >= 8 input channels and <= 64-wide channels."""
        n, c, d, h, w = in_shape
        cp, fp = ops.round_up(c, 16), ops.round_up(self.filters, 16)
        even = self.stride == 1 or self.transposed or (d % 2 == 0 and h % 2 == 0 and w % 2 == 0)
        return c >= 8 and cp <= 64 and fp <= 64 and even

    def __call__(self, tensor):
        return _run_transform(self, tensor, self.data_format)


class Conv3D(_ConvBase):
    transposed = False


class Conv3DTranspose(_ConvBase):
    transposed = True


class SequentialLayer(Layer):  # model_transforms.py:11-19
    def __init__(self, layers, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._layers = layers

    def leaf_layers(self):
        return [l for layer in self._layers for l in layer.leaf_layers()]

    @property
    def data_format(self):
        return self.leaf_layers()[0].data_format

    def get_weights(self):
        return [l.get_weights() for l in self.leaf_layers()]

    def set_weights(self, weights):
        leaves = self.leaf_layers()
        assert len(weights) == len(leaves)
        for l, w in zip(leaves, weights):
            l.set_weights(w['kernel'], w.get('bias'))

    def call(self, tensor, **kwargs):
        return _run_transform(self, tensor, self.data_format)

    __call__ = call

    def packed(self, tensor, thresholds, want_f32=False):
        """Synthesis + what the block loops do with x_hat (model_types.py:201-202,209,233-234): returns
        (x_hat fp32 or None, packed occupancy bits of min(x_hat,1) > thresholds[n], per-block counts).  Fused into the
        last layer's epilogue when that layer is the 3x3x3 single-channel one; otherwise threshold_pack runs after it."""
        return _run_transform(self, tensor, self.data_format, pack={'thresholds': thresholds, 'want_f32': want_f32})


class ResidualLayer(Layer):  # model_transforms.py:22-38
    def __init__(self, layers, residual_mode='add', data_format=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert residual_mode in ('add', 'concat')
        self._layers = layers
        self.residual_mode = residual_mode
        self.data_format = normalize_data_format(data_format)

    def leaf_layers(self):
        return [l for layer in self._layers for l in layer.leaf_layers()]

    get_weights = SequentialLayer.get_weights
    set_weights = SequentialLayer.set_weights

    def call(self, tensor, **kwargs):
        return _run_transform(self, tensor, self.data_format)

    __call__ = call


class AnalysisTransformV1(SequentialLayer):  # model_transforms.py:41-48
    def __init__(self, filters, data_format=None, activation=relu, *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = {'strides': (2, 2, 2), 'padding': 'same', 'data_format': data_format, 'filters': filters}
        layers = [Conv3D(kernel_size=(9, 9, 9), use_bias=True, activation=activation, **params),
                  Conv3D(kernel_size=(5, 5, 5), use_bias=True, activation=activation, **params),
                  Conv3D(kernel_size=(5, 5, 5), use_bias=False, activation=None, **params)]
        super().__init__(layers, *args, **kwargs)


class SynthesisTransformV1(SequentialLayer):  # model_transforms.py:51-59
    def __init__(self, filters, data_format=None, activation=relu, *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = {'strides': (2, 2, 2), 'padding': 'same', 'data_format': data_format, 'use_bias': True,
                  'activation': activation}
        layers = [Conv3DTranspose(filters, (5, 5, 5), **params),
                  Conv3DTranspose(filters, (5, 5, 5), **params),
                  Conv3DTranspose(1, (9, 9, 9), **params)]
        super().__init__(layers, *args, **kwargs)


class AnalysisBlock(ResidualLayer):  # model_transforms.py:62-70
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), strides=(2, 2, 2), activation=relu, *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = {'padding': 'same', 'data_format': data_format, 'use_bias': True, 'activation': activation,
                  'filters': filters, 'kernel_size': kernel_size}
        layers = [Conv3D(strides=strides, **params), Conv3D(**params), Conv3D(**params)]
        super().__init__(layers, *args, data_format=data_format, **kwargs)


class SynthesisBlock(ResidualLayer):  # model_transforms.py:73-81
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), strides=(2, 2, 2), activation=relu, *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = {'padding': 'same', 'data_format': data_format, 'use_bias': True, 'activation': activation,
                  'filters': filters, 'kernel_size': kernel_size}
        layers = [Conv3DTranspose(strides=strides, **params), Conv3DTranspose(**params), Conv3DTranspose(**params)]
        super().__init__(layers, *args, data_format=data_format, **kwargs)


def _v2_params(data_format, kernel_size, activation, residual_mode):
    return {'kernel_size': kernel_size, 'activation': activation, 'data_format': data_format,
            'residual_mode': residual_mode}


class AnalysisTransformV2(SequentialLayer):  # model_transforms.py:84-95
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), activation=relu, residual_mode='add', *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = _v2_params(data_format, kernel_size, activation, residual_mode)
        layers = [AnalysisBlock(filters // 2, **params), AnalysisBlock(filters, **params), AnalysisBlock(filters, **params),
                  Conv3D(filters, kernel_size, padding="same", use_bias=False, activation=None, data_format=data_format)]
        super().__init__(layers, *args, **kwargs)


class SynthesisTransformV2(SequentialLayer):  # model_transforms.py:98-109
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), activation=relu, residual_mode='add', *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = _v2_params(data_format, kernel_size, activation, residual_mode)
        layers = [SynthesisBlock(filters, **params), SynthesisBlock(filters, **params), SynthesisBlock(filters // 2, **params),
                  Conv3DTranspose(1, kernel_size, padding="same", use_bias=True, activation=activation, data_format=data_format)]
        super().__init__(layers, *args, **kwargs)


class AnalysisTransformProgressiveV2(SequentialLayer):  # model_transforms.py:112-123
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), activation=relu, residual_mode='add', *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = _v2_params(data_format, kernel_size, activation, residual_mode)
        layers = [AnalysisBlock(filters // 4, **params), AnalysisBlock(filters // 2, **params), AnalysisBlock(filters, **params),
                  Conv3D(filters, kernel_size, padding="same", use_bias=False, activation=None, data_format=data_format)]
        super().__init__(layers, *args, **kwargs)


class SynthesisTransformProgressiveV2(SequentialLayer):  # model_transforms.py:126-137
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), activation=relu, residual_mode='add', *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = _v2_params(data_format, kernel_size, activation, residual_mode)
        layers = [SynthesisBlock(filters, **params), SynthesisBlock(filters // 2, **params), SynthesisBlock(filters // 4, **params),
                  Conv3DTranspose(1, kernel_size, padding="same", use_bias=True, activation=activation, data_format=data_format)]
        super().__init__(layers, *args, **kwargs)


class HyperAnalysisTransform(SequentialLayer):  # model_transforms.py:140-147
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), activation=relu, *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = {'padding': 'same', 'data_format': data_format, 'filters': filters, 'kernel_size': kernel_size}
        layers = [Conv3D(use_bias=True, activation=activation, **params),
                  Conv3D(use_bias=True, activation=activation, strides=(2, 2, 2), **params),
                  Conv3D(use_bias=False, activation=None, **params)]
        super().__init__(layers, *args, **kwargs)


class HyperSynthesisTransform(SequentialLayer):  # model_transforms.py:150-158
    def __init__(self, filters, data_format=None, kernel_size=(3, 3, 3), activation=relu, *args, **kwargs):
        data_format = normalize_data_format(data_format)
        params = {'padding': 'same', 'data_format': data_format, 'activation': activation, 'use_bias': True,
                  'filters': filters, 'kernel_size': kernel_size}
        layers = [Conv3DTranspose(**params), Conv3DTranspose(strides=(2, 2, 2), **params), Conv3DTranspose(**params)]
        super().__init__(layers, *args, **kwargs)


class TransformType(Enum):  # model_transforms.py:161-169
    AnalysisTransformV1 = AnalysisTransformV1
    AnalysisTransformV2 = AnalysisTransformV2
    AnalysisTransformProgressiveV2 = AnalysisTransformProgressiveV2
    SynthesisTransformV1 = SynthesisTransformV1
    SynthesisTransformV2 = SynthesisTransformV2
    SynthesisTransformProgressiveV2 = SynthesisTransformProgressiveV2
    HyperAnalysisTransform = HyperAnalysisTransform
    HyperSynthesisTransform = HyperSynthesisTransform


# ---------------------------------------------------------------------------------------------------------
# tracing + execution
# ---------------------------------------------------------------------------------------------------------
def trace(layer, fuse_residual=True):
    """Flatten a layer tree into steps.  Value ids: 0 is the input.
    ('conv', layer, src, dst, res) with res = id of the residual operand fused into the epilogue (or None);
    ('add', a, b, dst); ('concat', a, b, dst)."""
    steps = []
    counter = [0]

    def new_id():
        counter[0] += 1
        return counter[0]

    def rec(l, src):
        if isinstance(l, _ConvBase):
            dst = new_id()
            steps.append(['conv', l, src, dst, None])
            return dst
        if isinstance(l, ResidualLayer):
            t1 = rec(l._layers[0], src)
            t = t1
            for sub in l._layers[1:]:
                t = rec(sub, t)
            if l.residual_mode == 'add':
                last = steps[-1]
                if fuse_residual and len(l._layers) > 1 and last[0] == 'conv' and last[3] == t and last[4] is None:
                    last[4] = t1  # fuse: out = t1 + relu(conv(...))
                    return t
                dst = new_id()
                steps.append(['add', t1, t, dst])
                return dst
            dst = new_id()
            steps.append(['concat', t, t1, dst])
            return dst
        if isinstance(l, SequentialLayer):
            for sub in l._layers:
                src = rec(sub, src)
            return src
        raise TypeError(f'cannot trace {type(l).__name__}')

    out = rec(layer, 0)
    return [tuple(s) for s in steps], out


class _Val:
    """A traced value in one of two representations: fp32 (N,C,D,H,W) or blocked bf16 (+ logical shape)."""
    __slots__ = ('f32', 'blk', 'shape', 'terms')

    def __init__(self, f32=None, blk=None, shape=None, terms=0):
        self.f32, self.blk, self.shape, self.terms = f32, blk, shape, terms

    def as_f32(self):
        if self.f32 is None:
            self.f32 = ops.blocked_to_f32(self.blk, self.shape, self.terms)
        return self.f32

    def as_blk(self, terms):
        if self.blk is None or self.terms != terms:
            self.blk = ops.f32_to_blocked(self.as_f32(), terms)
            self.terms = terms
        return self.blk


def run_steps(steps, out_id, x, pack=None, keep=None, extra=None, keep_vals=None, return_val=False):
    """Execute traced steps on a channels_first fp32 CUDA tensor.  pack = {'thresholds', 'want_f32'}: also return the
    packed thresholded occupancy of the (single-channel) output -> (y or None, bits, counts).  keep: dict that receives every
    value (id -> fp32 tensor; the training path saves activations this way; keep_vals: the same ids -> _Val with the blocked form
    when one exists); extra: {id: fp32 tensor} of additional inputs (e.g. a residual operand).  x may be a _Val."""
    mode = _precision['mode']
    terms = {'bf16x3': 2, 'bf16': 1, 'fp32': 0}[mode]
    vals = {0: x if isinstance(x, _Val) else _Val(f32=x, shape=tuple(x.shape))}   # a _Val input brings its blocked form along
    for vid, t in (extra or {}).items():
        vals[vid] = t if isinstance(t, _Val) else _Val(f32=t, shape=tuple(t.shape))
    last_use = {}
    for i, s in enumerate(steps):
        for vid in ((s[2], s[4]) if s[0] == 'conv' else (s[1], s[2])):
            if vid is not None:
                last_use[vid] = i
    for i, s in enumerate(steps):
        if s[0] == 'conv':
            _, layer, src, dst, res = s
            v = vals[src]
            layer.build(v.shape[1])
            if v.shape[1] != layer.in_channels:
                raise ValueError(f'layer expects {layer.in_channels} input channels, got {v.shape[1]}')
            if terms and res is None and dst == out_id and layer.out1_eligible(v.shape):
                thr = pack['thresholds'] if pack else None
                xh, bits, counts = ops.conv3d_out1(v.as_blk(terms), v.shape, layer.dev(f'w_out1{terms}'), layer.dev('bias'),
                                                   layer.relu, terms, (not pack) or pack['want_f32'], thr)
                if pack:
                    return xh, bits, counts
                vals[dst] = _Val(f32=xh, shape=tuple(xh.shape))
            elif res is None and dst != out_id and layer.first_eligible(v.shape, terms, v.f32 is not None):
                yb, shp = ops.conv3d_first(v.f32, layer.dev('w_tap'), layer.dev('bias'), layer.filters, layer.relu, terms)
                vals[dst] = _Val(blk=yb, shape=shp, terms=terms)
            elif terms and layer.ys_eligible(v.shape):
                rb = vals[res].as_blk(terms) if res is not None else None
                yb, shp = ops.conv3d_umma_ys(v.as_blk(terms), v.shape, layer.dev(f'w_ummays{terms}'), layer.dev('bias'),
                                             layer.filters, layer.relu, terms, rb)
                vals[dst] = _Val(blk=yb, shape=shp, terms=terms)
            elif layer.zy_eligible(v.shape, terms):
                rb = vals[res].as_blk(terms) if res is not None else None
                yb, shp = ops.conv3d_umma_zy(v.as_blk(terms), v.shape, layer.dev(f'w_ummazy{terms}'), layer.dev('bias'), layer.filters,
                                             layer.relu, terms, rb)
                vals[dst] = _Val(blk=yb, shape=shp, terms=terms)
            elif layer.hl_eligible(v.shape, terms):
                rb = vals[res].as_blk(terms) if res is not None else None
                yb, shp = ops.conv3d_umma_hl(v.as_blk(terms), v.shape, layer.dev('w_ummahl'), layer.dev('bias'), layer.filters,
                                             layer.transposed, layer.relu, rb)
                vals[dst] = _Val(blk=yb, shape=shp, terms=terms)
            elif terms and layer.umma_eligible(v.shape):
                rb = vals[res].as_blk(terms) if res is not None else None
                yb, shp = ops.conv3d_umma(v.as_blk(terms), v.shape, layer.dev(f'w_umma{terms}'), layer.dev('bias'),
                                          layer.filters, layer.stride, layer.transposed, layer.relu, terms, rb)
                vals[dst] = _Val(blk=yb, shape=shp, terms=terms)
            elif terms and layer.gemm_eligible(v.shape):
                rb = vals[res].as_blk(terms) if res is not None else None
                yb, shp = ops.conv3d_gemm(v.as_blk(terms), v.shape, layer.dev(f'w_gemm{terms}'), layer.dev('bias'),
                                          layer.filters, layer.stride, layer.transposed, layer.relu, terms, rb)
                vals[dst] = _Val(blk=yb, shape=shp, terms=terms)
            else:
                rf = vals[res].as_f32() if res is not None else None
                y = ops.conv3d_f32(v.as_f32(), layer.dev('w_tap'), layer.dev('bias'), layer.filters, layer.k, layer.stride,
                                   layer.transposed, layer.relu, rf)
                vals[dst] = _Val(f32=y, shape=tuple(y.shape))
        elif s[0] == 'add':
            y = ops.axpby(vals[s[1]].as_f32(), vals[s[2]].as_f32(), 1.0, 1.0)  # un-fused residual add (training traces)
            vals[s[3]] = _Val(f32=y, shape=tuple(y.shape))
        else:
            y = torch.cat((vals[s[1]].as_f32(), vals[s[2]].as_f32()), 1)
            vals[s[3]] = _Val(f32=y, shape=tuple(y.shape))
        if keep is not None:
            keep[s[3]] = vals[s[3]].as_f32()
            if keep_vals is not None:
                keep_vals[s[3]] = vals[s[3]]      # both representations (the tcgen05 weight gradient reads the blocked one)
            continue
        for vid in [k for k, li in last_use.items() if li == i and k != out_id]:
            vals.pop(vid, None)
    if return_val:          # the value in whatever representation the last kernel produced (the training path's backward chain)
        return vals[out_id]
    y = vals[out_id].as_f32()
    if pack:
        bits, counts = ops.threshold_pack(y, pack['thresholds'])
        return y, bits, counts
    return y


def layer_runner(layer, xb, in_shape, terms, residual_b=None):
    """(callable, kernel name) running ONE conv layer on a blocked bf16 input into a preallocated blocked output through the same
    kernel dispatch as run_steps -- what bench.py times alone for the roofline of the dominant layer."""
    if layer.ys_eligible(in_shape):
        out, _ = ops.conv3d_umma_ys(xb, in_shape, layer.dev(f'w_ummays{terms}'), layer.dev('bias'), layer.filters, layer.relu, terms, residual_b)
        return (lambda: ops.conv3d_umma_ys(xb, in_shape, layer.dev(f'w_ummays{terms}'), layer.dev('bias'), layer.filters, layer.relu, terms,
                                           residual_b, out)), 'conv3d_umma_ys_kernel'
    if layer.zy_eligible(in_shape, terms):
        out, _ = ops.conv3d_umma_zy(xb, in_shape, layer.dev(f'w_ummazy{terms}'), layer.dev('bias'), layer.filters, layer.relu, terms, residual_b)
        return (lambda: ops.conv3d_umma_zy(xb, in_shape, layer.dev(f'w_ummazy{terms}'), layer.dev('bias'), layer.filters, layer.relu, terms,
                                           residual_b, out)), 'conv3d_umma_zy_kernel'
    if layer.hl_eligible(in_shape, terms):
        out, _ = ops.conv3d_umma_hl(xb, in_shape, layer.dev('w_ummahl'), layer.dev('bias'), layer.filters, layer.transposed, layer.relu, residual_b)
        return (lambda: ops.conv3d_umma_hl(xb, in_shape, layer.dev('w_ummahl'), layer.dev('bias'), layer.filters, layer.transposed,
                                           layer.relu, residual_b, out)), 'conv3d_umma_kernel<16,hl>'
    if layer.umma_eligible(in_shape):
        out, _ = ops.conv3d_umma(xb, in_shape, layer.dev(f'w_umma{terms}'), layer.dev('bias'), layer.filters, layer.stride, layer.transposed,
                                 layer.relu, terms, residual_b)
        return (lambda: ops.conv3d_umma(xb, in_shape, layer.dev(f'w_umma{terms}'), layer.dev('bias'), layer.filters, layer.stride,
                                        layer.transposed, layer.relu, terms, residual_b, out)), f'conv3d_umma_kernel<{ops.round_up(layer.filters, 16)}>'
    if layer.gemm_eligible(in_shape):
        out, _ = ops.conv3d_gemm(xb, in_shape, layer.dev(f'w_gemm{terms}'), layer.dev('bias'), layer.filters, layer.stride, layer.transposed,
                                 layer.relu, terms, residual_b)
        return (lambda: ops.conv3d_gemm(xb, in_shape, layer.dev(f'w_gemm{terms}'), layer.dev('bias'), layer.filters, layer.stride,
                                        layer.transposed, layer.relu, terms, residual_b, out)), 'conv3d_gemm_kernel'
    raise ValueError('layer is not served by a tensor-core kernel')


def run_layer(layer, x, residual=None, return_val=False):
    """One conv layer (fp32 or _Val in / fp32 out, or a _Val with return_val) through the same kernel dispatch as a transform; residual
    is added in the epilogue."""
    steps = [('conv', layer, 0, 1, 2 if residual is not None else None)]
    return run_steps(steps, 1, x, extra={2: residual} if residual is not None else None, return_val=return_val)


def _run_transform(layer, tensor, data_format, pack=None):
    if not (torch.is_tensor(tensor) and tensor.is_cuda):
        raise TypeError('transforms run on CUDA tensors only (no CPU fallback); move the input to the GPU')
    x = tensor.to(torch.float32)
    if data_format == 'channels_last':
        x = x.permute(0, 4, 1, 2, 3)
    x = x.contiguous()
    if not hasattr(layer, '_trace'):
        layer._trace = trace(layer)
    steps, out_id = layer._trace
    y = run_steps(steps, out_id, x, pack)
    if pack:
        xh, bits, counts = y
        if xh is not None and data_format == 'channels_last':
            xh = xh.permute(0, 2, 3, 4, 1).contiguous()
        return xh, bits, counts
    if data_format == 'channels_last':
        y = y.permute(0, 2, 3, 4, 1).contiguous()
    return y
