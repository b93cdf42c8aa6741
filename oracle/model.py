"""Oracle (test infrastructure): CPU restatement of reference src/model_types.py (graphs `train`,
`compress`, `decompress` of CompressionModelV1 :241-309 and CompressionModelV2 :312-411), the config
table src/model_configs.py:16-42 and src/utils/focal_loss.py:5-12.

PARITY UNPINNED (see oracle/__init__.py).  torch-CPU fp32 by default; pass dtype=torch.float64 to bound
the oracle's own rounding.
"""
import math

import numpy as np
import torch

from . import entropy as E
from . import range_coder as RC
from . import transforms as T

# src/model_configs.py:16-42
CONFIGS = {
    'c1': dict(version=1, num_filters=32, analysis='AnalysisTransformV1', synthesis='SynthesisTransformV1'),
    'c2': dict(version=2, num_filters=32, analysis='AnalysisTransformV1', synthesis='SynthesisTransformV1',
               hyper_analysis='HyperAnalysisTransform', hyper_synthesis='HyperSynthesisTransform'),
    'c3': dict(version=2, num_filters=32, analysis='AnalysisTransformV2', synthesis='SynthesisTransformV2',
               hyper_analysis='HyperAnalysisTransform', hyper_synthesis='HyperSynthesisTransform'),
    'c3p': dict(version=2, num_filters=64, analysis='AnalysisTransformProgressiveV2',
                synthesis='SynthesisTransformProgressiveV2',
                hyper_analysis='HyperAnalysisTransform', hyper_synthesis='HyperSynthesisTransform'),
}


def focal_loss(y_true, y_pred, gamma=2, alpha=0.9):
    """src/utils/focal_loss.py:5-12 -- sum-reduced, clip [1e-3, .999]."""
    pt_1 = torch.where(y_true == 1, y_pred, torch.ones_like(y_pred))
    pt_0 = torch.where(y_true == 0, y_pred, torch.zeros_like(y_pred))
    pt_1 = torch.clamp(pt_1, 1e-3, .999)
    pt_0 = torch.clamp(pt_0, 1e-3, .999)
    return -torch.sum(alpha * torch.pow(1. - pt_1, gamma) * torch.log(pt_1)) \
           - torch.sum((1 - alpha) * torch.pow(pt_0, gamma) * torch.log(1. - pt_0))


def sparse_to_dense(block, x_shape):
    """src/model_types.py:108-114 (channels_first)."""
    x = np.zeros(x_shape, np.float32)
    b = np.asarray(block).astype(np.uint32)
    x[0, 0, b[:, 0], b[:, 1], b[:, 2]] = 1.0
    return x


class OracleModel:
    """Holds the spec trees + a flat parameter set; mirrors what the reference's graphs compute."""

    def __init__(self, config, dtype=torch.float32):
        cfg = CONFIGS[config]
        self.cfg, self.dtype = cfg, dtype
        f = cfg['num_filters']
        self.num_filters = f
        self.spec = {'analysis': T.build_transform(cfg['analysis'], f),
                     'synthesis': T.build_transform(cfg['synthesis'], f)}
        if cfg['version'] == 2:
            self.spec['hyper_analysis'] = T.build_transform(cfg['hyper_analysis'], f)
            self.spec['hyper_synthesis'] = T.build_transform(cfg['hyper_synthesis'], f)
            self.scale_table = E.make_scale_table()
        self.thresholds = np.linspace(0, 1.0, 256)  # model_types.py:181
        self.weights = None
        self.eb = None
        self._eb_tab = None
        self._gc_tab = None

    # -- parameters ---------------------------------------------------------------------------
    def init_params(self, seed=42, bias_scale=0.0):
        rng = np.random.default_rng(seed)
        f = self.num_filters
        in_ch = {'analysis': 1, 'synthesis': f, 'hyper_analysis': f, 'hyper_synthesis': f}
        self.weights = {k: T.init_weights(s, in_ch[k], rng, bias_scale, self.dtype) for k, s in self.spec.items()}
        self.eb = E.eb_init(f, rng)
        self._eb_tab = self._gc_tab = None

    def set_params(self, weights, eb):
        """weights: {transform: [ {'kernel','bias'} ... ]} (numpy/torch, Keras layouts); eb: eb_init-like dict."""
        self.weights = {k: T.weights_to(v, self.dtype) for k, v in weights.items()}
        self.eb = {k: ([np.asarray(a) for a in v] if isinstance(v, (list, tuple)) else np.asarray(v))
                   for k, v in eb.items()}
        self._eb_tab = self._gc_tab = None

    @property
    def eb_tab(self):
        if self._eb_tab is None:
            self._eb_tab = E.eb_tables(self.eb)
        return self._eb_tab

    @property
    def gc_tab(self):
        if self._gc_tab is None:
            self._gc_tab = E.gc_tables(self.scale_table)
        return self._gc_tab

    def _tf(self, name, x):
        return T.apply_transform(self.spec[name], self.weights[name], x)

    # -- graphs -------------------------------------------------------------------------------
    def analyse(self, x):
        """x (N,1,D,H,W) -> dict of the symbol-level intermediates of the compress graph (no range coding)."""
        x = torch.as_tensor(x).to(self.dtype)
        y = self._tf('analysis', x)
        if self.cfg['version'] == 1:  # model_types.py:283-295
            y_sym = E.eb_symbols(self.eb, y)
            y_hat = E.eb_dequantize(self.eb, y_sym, self.dtype)
            return {'y': y, 'y_symbols': y_sym, 'y_hat': y_hat}
        z = self._tf('hyper_analysis', y)  # :371-391
        z_sym = E.eb_symbols(self.eb, z)
        z_hat = E.eb_dequantize(self.eb, z_sym, self.dtype)
        sigma_hat = self._tf('hyper_synthesis', z_hat)
        idx = E.gc_indexes(sigma_hat, self.scale_table)
        y_sym = E.gc_symbols(y)
        y_hat = y_sym.to(self.dtype)
        return {'y': y, 'z': z, 'z_symbols': z_sym, 'z_hat': z_hat, 'sigma_hat': sigma_hat,
                'indexes': idx, 'y_symbols': y_sym, 'y_hat': y_hat}

    def synthesise(self, y_hat):
        return self._tf('synthesis', torch.as_tensor(y_hat).to(self.dtype))

    def _channel_indexes(self, shape):
        C = shape[0]
        return np.broadcast_to(np.arange(C, dtype=np.int32).reshape(C, 1, 1, 1), shape)

    def compress(self, x):
        """One sample x (1,1,D,H,W) -> (strings tuple, x_hat, debug)."""
        t = self.analyse(x)
        x_hat = self.synthesise(t['y_hat'])
        et = self.eb_tab
        if self.cfg['version'] == 1:
            sym = t['y_symbols'][0].numpy()
            s = RC.unbounded_index_range_encode(sym, self._channel_indexes(sym.shape), et['cdf'], et['cdf_length'],
                                                et['offset'])
            return (s,), x_hat, t
        zs = t['z_symbols'][0].numpy()
        z_string = RC.unbounded_index_range_encode(zs, self._channel_indexes(zs.shape), et['cdf'],
                                                   et['cdf_length'], et['offset'])
        gt = self.gc_tab
        y_string = RC.unbounded_index_range_encode(t['y_symbols'][0].numpy(), t['indexes'][0].numpy(), gt['cdf'],
                                                   gt['cdf_length'], gt['offset'])
        return (y_string, z_string), x_hat, t  # order: model_types.py:389

    def decompress(self, strings, x_shape):
        """strings tuple + spatial shape (D,H,W) -> x_hat (1,1,D,H,W).  model_types.py:297-309 / 393-411."""
        f = self.num_filters
        et = self.eb_tab
        if self.cfg['version'] == 1:
            shp = (f,) + tuple(int(s) // 8 for s in x_shape)
            sym = RC.unbounded_index_range_decode(strings[0], self._channel_indexes(shp), et['cdf'],
                                                  et['cdf_length'], et['offset'])
            y_hat = E.eb_dequantize(self.eb, torch.from_numpy(sym)[None], self.dtype)
            return self.synthesise(y_hat), {'y_hat': y_hat}
        zshp = (f,) + tuple(int(s) // 16 for s in x_shape)
        zs = RC.unbounded_index_range_decode(strings[1], self._channel_indexes(zshp), et['cdf'], et['cdf_length'],
                                             et['offset'])
        z_hat = E.eb_dequantize(self.eb, torch.from_numpy(zs)[None], self.dtype)
        sigma_hat = self._tf('hyper_synthesis', z_hat)
        idx = E.gc_indexes(sigma_hat, self.scale_table)
        gt = self.gc_tab
        ys = RC.unbounded_index_range_decode(strings[0], idx[0].numpy(), gt['cdf'], gt['cdf_length'], gt['offset'])
        y_hat = torch.from_numpy(ys)[None].to(self.dtype)
        return self.synthesise(y_hat), {'z_hat': z_hat, 'sigma_hat': sigma_hat, 'indexes': idx, 'y_hat': y_hat}

    def train_forward(self, x, gamma, alpha, lmbda, noise_y=None, noise_z=None):
        """model_types.py:250-274 / 327-355: returns dict(loss, fl, mbpov, x_tilde, ...)."""
        x = torch.as_tensor(x).to(self.dtype)
        y = self._tf('analysis', x)
        n_occ = torch.sum(x)
        denom = -math.log(2) * n_occ
        out = {}
        if self.cfg['version'] == 1:
            y_tilde, y_lik = E.eb_forward(self.eb, y, True, noise_y, self.dtype)
            mbpov = torch.sum(torch.log(y_lik)) / denom
        else:
            z = self._tf('hyper_analysis', y)
            z_tilde, z_lik = E.eb_forward(self.eb, z, True, noise_z, self.dtype)
            sigma = self._tf('hyper_synthesis', z_tilde)
            y_tilde, y_lik = E.gc_forward(y, sigma, self.scale_table, True, noise_y, self.dtype)
            mb_y = torch.sum(torch.log(y_lik)) / denom
            mb_z = torch.sum(torch.log(z_lik)) / denom
            mbpov = mb_y + mb_z
            out.update(z=z, z_tilde=z_tilde, z_likelihoods=z_lik, sigma_tilde=sigma, mbpov_y=mb_y, mbpov_z=mb_z)
        x_tilde = self._tf('synthesis', y_tilde)
        fl = focal_loss(x, x_tilde, gamma=gamma, alpha=alpha)
        out.update(y=y, y_tilde=y_tilde, y_likelihoods=y_lik, x_tilde=x_tilde, fl=fl, mbpov=mbpov,
                   loss=lmbda * fl + mbpov, num_occupied_voxels=n_occ)
        return out
