"""Host mirror of the tensorflow-compression 1.3 surface the reference uses (src/model_types.py:7,20,254,287,
300,333,340,377,385,397,406 and src/utils/patch_gaussian_conditional.py):

    EntropyBottleneck(data_format=...)      eb(x, training=True) -> (x_tilde, likelihoods)
                                            eb.compress(x) -> [bytes]*N ; eb.decompress(strings, shape, channels=C)
                                            eb.losses[0] ; eb.updates[0]
    GaussianConditional(scale, scale_table) gc(y, training=True) -> (y_tilde, likelihoods)
                                            gc.compress(y) -> [bytes]*N ; gc.decompress(strings) ; gc.dbg_dec

Per-element arithmetic (quantisation, likelihoods, scale->index search) runs in libpccgeo CUDA kernels; the
16-bit CDF tables are built once on the host in float64 (numpy) and quantised by the C++
pmf_to_quantized_cdf; range coding is the C++ host coder (one independent stream per sample).
Tensors are torch CUDA fp32, channels_first (N,C,D,H,W) or channels_last.
"""
import math

import numpy as np
import torch

from . import ops

LIKELIHOOD_BOUND = 1e-9
PRECISION = 16


def _softplus(a):
    return np.logaddexp(0.0, a)


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


class EntropyBottleneck:
    """tfc.EntropyBottleneck (factorized prior), tfc 1.3 defaults."""

    def __init__(self, init_scale=10, filters=(3, 3, 3), tail_mass=1e-9, likelihood_bound=1e-9,
                 range_coder_precision=16, data_format='channels_last', seed=None, **kwargs):
        assert tuple(filters) == (3, 3, 3), 'the CUDA kernels implement the default (3,3,3) bottleneck'
        assert likelihood_bound == LIKELIHOOD_BOUND and range_coder_precision == PRECISION
        self.init_scale, self.filters, self.tail_mass = float(init_scale), tuple(filters), float(tail_mass)
        self.data_format = data_format
        self._rng = np.random.default_rng(42 if seed is None else seed)
        self.matrices = self.biases = self.factors = self.quantiles = None
        self._tables = None
        self._dev_params = None

    # -- variables -----------------------------------------------------------------------------------
    def build(self, channels):
        if self.matrices is not None:
            assert self.matrices[0].shape[0] == channels
            return
        r = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1.0 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = [], [], []
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1.0 / scale / r[i + 1]))
            self.matrices.append(np.full((channels, r[i + 1], r[i]), init, np.float32))
            self.biases.append(self._rng.uniform(-0.5, 0.5, size=(channels, r[i + 1], 1)).astype(np.float32))
            if i < len(self.filters):
                self.factors.append(np.zeros((channels, r[i + 1], 1), np.float32))
        self.quantiles = np.tile(np.array([[[-self.init_scale, 0.0, self.init_scale]]], np.float32), (channels, 1, 1))
        self._invalidate()

    def get_weights(self):
        return {'matrices': self.matrices, 'biases': self.biases, 'factors': self.factors, 'quantiles': self.quantiles}

    def set_weights(self, w):
        self.matrices = [np.asarray(a, np.float32) for a in w['matrices']]
        self.biases = [np.asarray(a, np.float32) for a in w['biases']]
        self.factors = [np.asarray(a, np.float32) for a in w['factors']]
        self.quantiles = np.asarray(w['quantiles'], np.float32)
        self._invalidate()

    def _invalidate(self):
        self._tables = None
        self._dev_params = None
        from .model_transforms import params_epoch
        params_epoch[0] += 1

    @property
    def channels(self):
        return self.matrices[0].shape[0]

    @property
    def medians(self):
        return self.quantiles[:, 0, 1]

    def device_params(self):
        """(C, 58) fp32 block consumed by the kernels (include/pccgeo.h, PCCGEO_EB_PARAM_STRIDE)."""
        if self._dev_params is None:
            C = self.channels
            p = np.zeros((C, 58), np.float32)
            sp = [_softplus(m.astype(np.float64)).astype(np.float32) for m in self.matrices]
            p[:, 0:3] = sp[0][:, :, 0]
            p[:, 3:12] = sp[1].reshape(C, 9)
            p[:, 12:21] = sp[2].reshape(C, 9)
            p[:, 21:24] = sp[3][:, 0, :]
            p[:, 24:27] = self.biases[0][:, :, 0]
            p[:, 27:30] = self.biases[1][:, :, 0]
            p[:, 30:33] = self.biases[2][:, :, 0]
            p[:, 33] = self.biases[3][:, 0, 0]
            for i in range(3):
                p[:, 34 + 3 * i:37 + 3 * i] = np.tanh(self.factors[i].astype(np.float64)).astype(np.float32)[:, :, 0]
            p[:, 43] = self.medians
            self._dev_params = ops.upload(p)
        return self._dev_params

    # -- host-side table construction (float64) ------------------------------------------------------
    def _logits_host(self, v):
        """v: (C,1,M) float64."""
        logits = v
        for i in range(len(self.matrices)):
            logits = np.matmul(_softplus(self.matrices[i].astype(np.float64)), logits) + self.biases[i].astype(np.float64)
            if i < len(self.factors):
                logits = logits + np.tanh(self.factors[i].astype(np.float64)) * np.tanh(logits)
        return logits

    @property
    def losses(self):
        """[auxiliary quantile loss] = sum |logits(quantiles) - (-T, 0, T)| (host float64 -> python float)."""
        target = math.log(2.0 / self.tail_mass - 1.0)
        logits = self._logits_host(self.quantiles.astype(np.float64))
        return [float(np.abs(logits - np.array([-target, 0.0, target])).sum())]

    @property
    def updates(self):
        """[callable] refreshing the quantised CDF tables from the current variables (tfc's update op)."""
        return [self._refresh_tables]

    def _refresh_tables(self):
        self._tables = None
        return self.tables

    @property
    def tables(self):
        if self._tables is None:
            q = self.quantiles.astype(np.float64)
            med = q[:, 0, 1]
            minima = np.maximum(np.ceil(med - q[:, 0, 0]).astype(np.int64), 0)
            maxima = np.maximum(np.ceil(q[:, 0, 2] - med).astype(np.int64), 0)
            pmf_start = med - minima
            pmf_length = (maxima + minima + 1).astype(np.int32)
            max_length = int(pmf_length.max())
            samples = np.arange(max_length, dtype=np.float64)[None, None, :] + pmf_start[:, None, None]
            lower, upper = self._logits_host(samples - 0.5), self._logits_host(samples + 0.5)
            sign = -np.sign(lower + upper)
            pmf = np.abs(_sigmoid(sign * upper) - _sigmoid(sign * lower))[:, 0, :]
            tail = (_sigmoid(lower[:, 0, :1]) + _sigmoid(-upper[:, 0, -1:]))[:, 0]
            C = q.shape[0]
            cdf = np.zeros((C, max_length + 2), np.int32)
            for c in range(C):
                L = int(pmf_length[c])
                cdf[c, :L + 2] = ops.pmf_to_quantized_cdf(np.concatenate([pmf[c, :L], tail[c:c + 1]]), PRECISION)
            self._tables = {'cdf': cdf, 'cdf_length': (pmf_length + 2).astype(np.int32), 'offset': (-minima).astype(np.int32)}
        return self._tables

    # -- layout helpers ------------------------------------------------------------------------------
    def _to_cf(self, x):
        x = x.to(torch.float32)
        if self.data_format == 'channels_last':
            x = x.permute(0, 4, 1, 2, 3)
        return x.contiguous()

    def _from_cf(self, x):
        return x.permute(0, 2, 3, 4, 1).contiguous() if self.data_format == 'channels_last' else x

    # -- forward -------------------------------------------------------------------------------------
    def __call__(self, inputs, training=True, noise=None):
        x = self._to_cf(inputs)
        self.build(x.shape[1])
        if training:
            if noise is None:
                noise = torch.rand_like(x) - 0.5
            values = x + self._to_cf(noise)
        else:
            _, values = ops.eb_quantize(x, self.device_params(), want_symbols=False)
        lik, _ = ops.eb_likelihood(values, self.device_params(), want_sum=False)
        return self._from_cf(values), self._from_cf(lik)

    def log_likelihood_sum(self, values):
        """sum(ln p(values)) as a device double[1] (mbpov numerator), deterministic reduction."""
        _, s = ops.eb_likelihood(self._to_cf(values), self.device_params(), want_likelihood=False)
        return s

    def quantize(self, inputs):
        """-> (int32 symbols, dequantised values) in channels_first."""
        x = self._to_cf(inputs)
        self.build(x.shape[1])
        return ops.eb_quantize(x, self.device_params())

    def compress(self, inputs, threads=0):
        sym, _ = self.quantize(inputs)
        return self.encode_symbols(sym.cpu().numpy(), threads)

    def encode_symbols(self, sym_host, threads=0):
        n = sym_host.shape[0]
        per = int(np.prod(sym_host.shape[1:]))
        spatial = per // self.channels
        offs = np.arange(n + 1, dtype=np.int64) * per
        return ops.range_encode(sym_host.reshape(-1), offs, self.tables, channel_stride=spatial, threads=threads)

    def decode_symbols(self, strings, shape_cf, threads=0):
        """strings: list of bytes; shape_cf = (C, D, H, W) -> int32 numpy (N, C, D, H, W)."""
        n = len(strings)
        per = int(np.prod(shape_cf))
        offs = np.arange(n + 1, dtype=np.int64) * per
        sym = ops.range_decode(strings, offs, self.tables, channel_stride=per // shape_cf[0], threads=threads)
        return sym.reshape((n,) + tuple(int(s) for s in shape_cf))

    def decompress(self, strings, shape, channels=None, threads=0):
        """strings [bytes]*N, shape = per-sample shape in the model's data_format (without batch)."""
        shape = tuple(int(s) for s in shape)
        shape_cf = shape if self.data_format == 'channels_first' else (shape[-1],) + shape[:-1]
        if channels is not None:
            assert shape_cf[0] == channels
        self.build(shape_cf[0])
        sym = self.decode_symbols(strings, shape_cf, threads)
        out = ops.eb_dequantize(torch.from_numpy(sym).cuda(), self.device_params())
        return self._from_cf(out)


def make_scale_table(scales_min=0.11, scales_max=256, scales_levels=64):
    return np.exp(np.linspace(np.log(scales_min), np.log(scales_max), scales_levels))


_gc_table_cache = {}


def gaussian_tables(scale_table, tail_mass=2 ** -8):
    """src/utils/patch_gaussian_conditional.py:62-97,118 in float64: pmf_center = ceil(scale * -Phi^-1(tail/2)),
    pmf over |j - center|, tail = 2*lower[:, 0], 16-bit quantised CDFs, offset = -center."""
    from scipy.special import erfc, ndtri
    key = (tuple(np.asarray(scale_table, np.float64).tolist()), tail_mass)
    if key in _gc_table_cache:
        return _gc_table_cache[key]
    st = np.asarray(scale_table, np.float64)
    multiplier = -float(ndtri(tail_mass / 2))
    center = np.ceil(st * multiplier).astype(np.int64)
    length = 2 * center + 1
    max_length = int(length.max())
    samples = np.abs(np.arange(max_length, dtype=np.int64)[None, :] - center[:, None]).astype(np.float64)

    def phi(x):
        return 0.5 * erfc(-(2 ** -0.5) * x)

    upper = phi((0.5 - samples) / st[:, None])
    lower = phi((-0.5 - samples) / st[:, None])
    pmf = upper - lower
    tail = 2 * lower[:, 0]
    cdf = np.zeros((len(st), max_length + 2), np.int32)
    for i in range(len(st)):
        L = int(length[i])
        cdf[i, :L + 2] = ops.pmf_to_quantized_cdf(np.concatenate([pmf[i, :L], tail[i:i + 1]]), PRECISION)
    t = {'cdf': cdf, 'cdf_length': (length + 2).astype(np.int32), 'offset': (-center).astype(np.int32)}
    _gc_table_cache[key] = t
    return t


_scale_table_dev_cache = {}


def _scale_table_dev(scale_table):
    """fp32 device copy of a scale table, uploaded once per table and device (a per-call upload from pageable memory
    would also be illegal inside CUDA-graph capture)."""
    key = (scale_table.tobytes(), torch.cuda.current_device())
    t = _scale_table_dev_cache.get(key)
    if t is None:
        t = _scale_table_dev_cache[key] = torch.from_numpy(scale_table.astype(np.float32)).cuda()
    return t


class GaussianConditional:
    """tfc.GaussianConditional(scale, scale_table) with the reference's patch (zero mean, scale_bound=None ->
    scales lower-bounded at scale_table[0], indexes = table search)."""

    def __init__(self, scale, scale_table, scale_bound=None, mean=None, dtype=None, tail_mass=2 ** -8,
                 likelihood_bound=1e-9, range_coder_precision=16, data_format='channels_first', **kwargs):
        assert mean is None and scale_bound is None, 'only the zero-mean, table-bounded configuration is implemented'
        assert likelihood_bound == LIKELIHOOD_BOUND and range_coder_precision == PRECISION
        self.scale_table = np.asarray(scale_table, np.float64)
        self.tail_mass = tail_mass
        self.data_format = data_format  # elementwise: only matters for the symbol order of the bitstream
        self._scale_in = scale
        self._scale = scale.to(torch.float32).contiguous()
        self._table_dev = _scale_table_dev(self.scale_table)
        self._indexes = None
        self.dbg_dec = {}

    @property
    def scale(self):
        return self._scale

    @property
    def tables(self):
        return gaussian_tables(self.scale_table, self.tail_mass)

    def _cf(self, t):
        """bitstream order is the C-order flatten of the per-sample tensor in the model's layout:
        elementwise kernels don't care, so tensors are used as they are."""
        return t.to(torch.float32).contiguous()

    def __call__(self, inputs, training=True, noise=None):
        y = self._cf(inputs)
        if training:
            if noise is None:
                noise = torch.rand_like(y) - 0.5
            values = y + self._cf(noise)
        else:
            _, values, _ = ops.gc_quantize(y, None, self._table_dev, want_symbols=False, want_indexes=False)
        lik, _ = ops.gc_likelihood(values, self._scale, float(np.float32(self.scale_table[0])), want_sum=False)
        return values, lik

    def log_likelihood_sum(self, values):
        _, s = ops.gc_likelihood(self._cf(values), self._scale, float(np.float32(self.scale_table[0])), want_likelihood=False)
        return s

    def indexes(self):
        if self._indexes is None:
            _, _, self._indexes = ops.gc_quantize(None, self._scale, self._table_dev, want_symbols=False, want_yhat=False)
        return self._indexes

    def quantize(self, inputs):
        """-> (symbols int32, y_hat fp32, indexes int32), all on the device."""
        sym, yh, idx = ops.gc_quantize(self._cf(inputs), self._scale, self._table_dev)
        self._indexes = idx
        return sym, yh, idx

    def compress(self, inputs, threads=0):
        sym, _, idx = self.quantize(inputs)
        return self.encode_symbols(sym.cpu().numpy(), idx.cpu().numpy(), threads)

    def encode_symbols(self, sym_host, idx_host, threads=0):
        n = sym_host.shape[0]
        per = int(np.prod(sym_host.shape[1:]))
        offs = np.arange(n + 1, dtype=np.int64) * per
        return ops.range_encode(sym_host.reshape(-1), offs, self.tables, indexes=idx_host.reshape(-1), threads=threads)

    def decode_symbols(self, strings, idx_host, threads=0):
        n = len(strings)
        per = int(np.prod(idx_host.shape[1:]))
        offs = np.arange(n + 1, dtype=np.int64) * per
        sym = ops.range_decode(strings, offs, self.tables, indexes=idx_host.reshape(-1), threads=threads)
        return sym.reshape(idx_host.shape)

    def decompress(self, strings, threads=0):
        idx = self.indexes()
        idx_host = idx.cpu().numpy()
        sym = self.decode_symbols(strings, idx_host, threads)
        sym_dev = torch.from_numpy(sym).cuda()
        out = ops.i32_to_f32(sym_dev)
        self.dbg_dec = {'decompress/strings': list(strings), 'decompress/build/scale_table': self.scale_table,
                        'decompress/build/_scale': self._scale, 'decompress/indexes': idx,
                        'decompress/quantized_cdf': self.tables['cdf'], 'decompress/symbols': sym_dev,
                        'decompress/outputs': out}
        return out
