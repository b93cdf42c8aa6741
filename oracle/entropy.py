"""Oracle (test infrastructure): CPU restatement of the tensorflow-compression 1.3 entropy models the
reference uses (reference requirements.txt:7; call sites src/model_types.py:254,287,300,333,340,377,385,
397,406; Gaussian table math restated in-repo at src/utils/patch_gaussian_conditional.py:49-125).

PARITY UNPINNED: tfc 1.3 is an un-vendored dependency that cannot be installed here; the algorithm below
is restated from its published source as recalled.  Deliberate, documented
deviations: CDF tables are computed in float64 (tfc: float32 TF ops) so that every implementation
(numpy / torch / C++) produces bit-identical tables.

Layout convention: tensors are channels_first (N, C, D, H, W).  tfc transposes to (C, 1, N*D*H*W)
internally for the factorized prior; we do the same.
"""
import heapq
import math

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------
# EntropyBottleneck (factorized prior) -- tfc 1.3 defaults
# ---------------------------------------------------------------------------------------------
EB_INIT_SCALE = 10.0
EB_FILTERS = (3, 3, 3)
EB_TAIL_MASS = 1e-9
LIKELIHOOD_BOUND = 1e-9
RANGE_CODER_PRECISION = 16


class _LowerBound(torch.autograd.Function):
    """tfc.math_ops.lower_bound: max(x, bound) with the 'identity_if_towards' gradient -- the gradient passes when
    x >= bound or when it is negative (a descent step would move x up, towards the bound)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x)
        ctx.bound = bound
        return torch.clamp(x, min=bound)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return torch.where((x >= ctx.bound) | (g < 0), g, torch.zeros_like(g)), None


def lower_bound(x, bound):
    return _LowerBound.apply(x, float(bound))


def eb_init(channels, rng, dtype=np.float32):
    """Variables as tfc creates them: matrix_i (C, r[i+1], r[i]) = ln(expm1(1/scale/r[i+1])),
    bias_i ~ U(-.5,.5), factor_i = 0, quantiles = (-init_scale, 0, init_scale)."""
    r = (1,) + EB_FILTERS + (1,)
    scale = EB_INIT_SCALE ** (1.0 / (len(EB_FILTERS) + 1))
    p = {'matrices': [], 'biases': [], 'factors': []}
    for i in range(len(EB_FILTERS) + 1):
        init = math.log(math.expm1(1.0 / scale / r[i + 1]))
        p['matrices'].append(np.full((channels, r[i + 1], r[i]), init, dtype))
        p['biases'].append(rng.uniform(-0.5, 0.5, size=(channels, r[i + 1], 1)).astype(dtype))
        if i < len(EB_FILTERS):
            p['factors'].append(np.zeros((channels, r[i + 1], 1), dtype))
    q = np.tile(np.array([[[-EB_INIT_SCALE, 0.0, EB_INIT_SCALE]]], dtype), (channels, 1, 1))
    p['quantiles'] = q
    return p


def _t(a, dtype):
    if torch.is_tensor(a):  # keeps autograd leaves intact (gradient tests)
        return a.to(dtype)
    return torch.as_tensor(np.asarray(a)).to(dtype)


def eb_logits_cumulative(p, v, dtype=torch.float32):
    """v: (C, 1, M).  logits = chain of  softplus(matrix) @ v + bias ; v += tanh(factor) * tanh(v)."""
    logits = v
    n = len(p['matrices'])
    for i in range(n):
        m = torch.nn.functional.softplus(_t(p['matrices'][i], dtype))
        logits = torch.matmul(m, logits) + _t(p['biases'][i], dtype)
        if i < len(p['factors']):
            logits = logits + torch.tanh(_t(p['factors'][i], dtype)) * torch.tanh(logits)
    return logits


def eb_likelihood_c1m(p, values, dtype=torch.float32):
    """values: (C,1,M) already noised / dequantised.  |sigmoid(s*u) - sigmoid(s*l)|, floored at 1e-9."""
    lower = eb_logits_cumulative(p, values - 0.5, dtype)
    upper = eb_logits_cumulative(p, values + 0.5, dtype)
    sign = -torch.sign(lower + upper).detach()
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    return lower_bound(lik, LIKELIHOOD_BOUND)


def eb_medians(p):
    return np.asarray(p['quantiles'])[:, 0, 1].copy()


def eb_forward(p, x, training, noise=None, dtype=torch.float32):
    """EntropyBottleneck.__call__: returns (x_tilde, likelihoods), both shaped like x (N,C,D,H,W)."""
    x = x.to(dtype)
    C = x.shape[1]
    med = _t(eb_medians(p), dtype).view(1, C, 1, 1, 1)
    if training:
        if noise is None:
            noise = torch.rand_like(x) - 0.5
        values = x + noise.to(dtype)
    else:
        values = torch.floor(x + 0.5 - med) + med
    v = values.transpose(0, 1).reshape(C, 1, -1)
    lik = eb_likelihood_c1m(p, v, dtype).reshape(C, x.shape[0], *x.shape[2:]).transpose(0, 1)
    return values, lik


def eb_aux_loss(p, dtype=torch.float32):
    """EntropyBottleneck.losses[0]: sum |logits(quantiles) - (-T, 0, T)|, T = ln(2/tail_mass - 1)."""
    target = math.log(2.0 / EB_TAIL_MASS - 1.0)
    logits = eb_logits_cumulative(p, _t(p['quantiles'], dtype), dtype)
    tgt = torch.tensor([-target, 0.0, target], dtype=dtype)
    return torch.sum(torch.abs(logits - tgt))


def eb_symbols(p, x):
    """int32 symbols floor(x + .5 - median); x channels_first."""
    med = torch.as_tensor(eb_medians(p)).to(x.dtype).view(1, -1, 1, 1, 1)
    return torch.floor(x + 0.5 - med).to(torch.int32)


def eb_dequantize(p, symbols, dtype=torch.float32):
    med = torch.as_tensor(eb_medians(p)).to(dtype).view(1, -1, 1, 1, 1)
    return symbols.to(dtype) + med


def eb_tables(p):
    """Codec tables (tfc EntropyBottleneck.build): offset = -minima, pmf over median-minima..median+maxima,
    tail mass as the escape symbol, quantised to 16 bits.  Computed in float64 (documented deviation)."""
    q = np.asarray(p['quantiles'], np.float64)
    med = q[:, 0, 1]
    minima = np.maximum(np.ceil(med - q[:, 0, 0]).astype(np.int64), 0)
    maxima = np.maximum(np.ceil(q[:, 0, 2] - med).astype(np.int64), 0)
    pmf_start = med - minima
    pmf_length = (maxima + minima + 1).astype(np.int32)
    max_length = int(pmf_length.max())
    samples = np.arange(max_length, dtype=np.float64)[None, None, :] + pmf_start[:, None, None]
    st = torch.from_numpy(samples)
    lower = eb_logits_cumulative(p, st - 0.5, torch.float64)
    upper = eb_logits_cumulative(p, st + 0.5, torch.float64)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :].numpy()
    tail = (torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])).numpy()[:, 0]
    C = q.shape[0]
    cdf = np.zeros((C, max_length + 2), np.int32)
    for c in range(C):
        L = int(pmf_length[c])
        row = pmf_to_quantized_cdf(np.concatenate([pmf[c, :L], tail[c:c + 1]]), RANGE_CODER_PRECISION)
        cdf[c, :L + 2] = row
    return {'cdf': cdf, 'cdf_length': (pmf_length + 2).astype(np.int32),
            'offset': (-minima).astype(np.int32), 'medians': med.astype(np.float32)}


# ---------------------------------------------------------------------------------------------
# pmf_to_quantized_cdf  (tfc range_coding_ops kernel, restated)
# ---------------------------------------------------------------------------------------------
def pmf_to_quantized_cdf(pmf, precision=16):
    """Quantise a (not re-normalised) pmf to integers summing to exactly 2**precision and return the CDF
    with a leading 0 (len(pmf)+1 entries).  value = max(1, rint(p * 2**precision)); the sum is then
    repaired greedily: while too large, decrement the entry (>1) whose code-length penalty
    p*(log2 v - log2(v-1)) is smallest; while too small, increment the entry whose gain
    p*(log2(v+1) - log2 v) is largest.  Ties break towards the lower index."""
    pmf = np.asarray(pmf, np.float64)
    target = 1 << precision
    v = np.maximum(1, np.rint(pmf * target)).astype(np.int64)
    total = int(v.sum())
    if total > target:
        heap = [(pmf[i] * (math.log2(v[i]) - math.log2(v[i] - 1)), i) for i in range(len(v)) if v[i] > 1]
        heapq.heapify(heap)
        while total > target:
            _, i = heapq.heappop(heap)
            v[i] -= 1
            total -= 1
            if v[i] > 1:
                heapq.heappush(heap, (pmf[i] * (math.log2(v[i]) - math.log2(v[i] - 1)), i))
    elif total < target:
        heap = [(-pmf[i] * (math.log2(v[i] + 1) - math.log2(v[i])), i) for i in range(len(v))]
        heapq.heapify(heap)
        while total < target:
            _, i = heapq.heappop(heap)
            v[i] += 1
            total += 1
            heapq.heappush(heap, (-pmf[i] * (math.log2(v[i] + 1) - math.log2(v[i])), i))
    cdf = np.zeros(len(v) + 1, np.int32)
    cdf[1:] = np.cumsum(v)
    return cdf


# ---------------------------------------------------------------------------------------------
# GaussianConditional (scale hyperprior) -- tfc 1.3 + reference patch
# ---------------------------------------------------------------------------------------------
GC_TAIL_MASS = 2.0 ** -8


def make_scale_table(scales_min=0.11, scales_max=256, scales_levels=64):
    """src/model_types.py:324"""
    return np.exp(np.linspace(np.log(scales_min), np.log(scales_max), scales_levels))


def _phi(x):
    """standardized cumulative, 0.5*erfc(-x/sqrt(2)) (patch_gaussian_conditional.py:72-73 -> tfc)."""
    return 0.5 * torch.special.erfc(-(2 ** -0.5) * x)


def _std_quantile(q):
    """Phi^-1(q) in float64 (tfc: scipy.stats.norm.ppf)."""
    from scipy.special import ndtri
    return float(ndtri(q))


def gc_tables(scale_table):
    """patch_gaussian_conditional.py:62-97,118.  float64 (documented deviation: reference float32)."""
    st = np.asarray(scale_table, np.float64)
    multiplier = -_std_quantile(GC_TAIL_MASS / 2)
    pmf_center = np.ceil(st * multiplier).astype(np.int64)
    pmf_length = 2 * pmf_center + 1
    max_length = int(pmf_length.max())
    samples = np.abs(np.arange(max_length, dtype=np.int64)[None, :] - pmf_center[:, None]).astype(np.float64)
    s = torch.from_numpy(samples)
    sc = torch.from_numpy(st)[:, None]
    upper = _phi((0.5 - s) / sc)
    lower = _phi((-0.5 - s) / sc)
    pmf = (upper - lower).numpy()
    tail = (2 * lower[:, :1]).numpy()[:, 0]
    cdf = np.zeros((len(st), max_length + 2), np.int32)
    for i in range(len(st)):
        L = int(pmf_length[i])
        cdf[i, :L + 2] = pmf_to_quantized_cdf(np.concatenate([pmf[i, :L], tail[i:i + 1]]), RANGE_CODER_PRECISION)
    return {'cdf': cdf, 'cdf_length': (pmf_length + 2).astype(np.int32),
            'offset': (-pmf_center).astype(np.int32)}


def gc_bound_scale(sigma, scale_table):
    """scale_bound=None -> lower-bound the scale at scale_table[0] (patch :57-60)."""
    lo = np.float32(scale_table[0])
    return lower_bound(sigma, float(lo))


def gc_indexes(sigma, scale_table):
    """patch_gaussian_conditional.py:106-116: idx = (L-1) - #{t in table[:-1] : scale <= t}, float32 compare."""
    st = torch.as_tensor(np.asarray(scale_table, np.float32))
    s = gc_bound_scale(sigma.to(torch.float32), scale_table)
    cnt = (s.unsqueeze(-1) <= st[:-1]).sum(-1)
    return (len(st) - 1 - cnt).to(torch.int32)


def gc_likelihood(values, sigma, scale_table, dtype=torch.float32):
    """Phi((.5-|v|)/s) - Phi((-.5-|v|)/s), floored at 1e-9."""
    s = gc_bound_scale(sigma.to(dtype), scale_table).to(dtype)
    v = torch.abs(values.to(dtype))
    lik = _phi((0.5 - v) / s) - _phi((-0.5 - v) / s)
    return lower_bound(lik, LIKELIHOOD_BOUND)


def gc_forward(y, sigma, scale_table, training, noise=None, dtype=torch.float32):
    y = y.to(dtype)
    if training:
        if noise is None:
            noise = torch.rand_like(y) - 0.5
        values = y + noise.to(dtype)
    else:
        values = torch.round(y)
    return values, gc_likelihood(values, sigma, scale_table, dtype)


def gc_symbols(y):
    return torch.round(y).to(torch.int32)
