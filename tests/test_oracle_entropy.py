"""Pins the oracle's entropy-model restatement with closed-form identities (the reference has no value tests
for this path): table validity, likelihood bounds, index search, quantisation round trips, focal loss."""
import math

import numpy as np
import torch

from oracle import entropy as E
from oracle.model import focal_loss


def test_pmf_to_quantized_cdf_sums_and_floor():
    rng = np.random.default_rng(0)
    for n in (2, 5, 100, 1481):
        p = rng.random(n) ** 4
        p /= p.sum()
        p[0] = 1e-12  # would round to zero: must still get a nonzero slot
        cdf = E.pmf_to_quantized_cdf(p, 16)
        assert cdf[0] == 0 and cdf[-1] == 65536 and len(cdf) == n + 1
        assert np.all(np.diff(cdf) >= 1)


def test_pmf_to_quantized_cdf_known_answer():
    assert E.pmf_to_quantized_cdf([0.5, 0.25, 0.25], 4).tolist() == [0, 8, 12, 16]
    # un-normalised input: values rint(p*16) = 3,3,3 -> 9, padded greedily where the gain is largest
    out = E.pmf_to_quantized_cdf([0.2, 0.2, 0.2], 4)
    assert out[-1] == 16 and np.all(np.diff(out) >= 5)


def test_gc_tables_shape_and_constants():
    st = E.make_scale_table()
    assert len(st) == 64 and abs(st[0] - 0.11) < 1e-12 and abs(st[-1] - 256) < 1e-9
    t = E.gc_tables(st)
    mult = -E._std_quantile(2 ** -9)
    assert abs(mult - 2.88563) < 1e-4
    center = -t['offset']
    assert center[0] == 1 and center[-1] == 739
    assert t['cdf'].shape == (64, 1481)
    for i in range(64):
        L = t['cdf_length'][i]
        row = t['cdf'][i, :L]
        assert row[0] == 0 and row[-1] == 65536 and np.all(np.diff(row) >= 1)
        assert np.all(t['cdf'][i, L:] == 0)
    # symmetric pmf
    row = np.diff(t['cdf'][10, :t['cdf_length'][10] - 1])
    assert np.abs(row - row[::-1]).max() <= 1


def test_gc_indexes_known_answers():
    st = E.make_scale_table()
    s = torch.tensor([0.0, 0.05, 0.11, float(np.float32(st[1])), float(np.float32(st[1])) * 1.0001, 1e6])
    idx = E.gc_indexes(s, st).tolist()
    assert idx[0] == 0 and idx[1] == 0 and idx[2] == 0      # lower-bounded at table[0]
    assert idx[3] == 1 and idx[4] == 2 and idx[5] == 63


def test_gc_likelihood_is_a_pmf():
    st = E.make_scale_table()
    for sigma in (0.05, 0.5, 3.0, 40.0):
        v = torch.arange(-400, 401, dtype=torch.float64)
        p = E.gc_likelihood(v, torch.full_like(v, sigma), st, torch.float64)
        assert abs(float(p.sum()) - 1.0) < 1e-6
        assert float(p.min()) >= 1e-9


def test_eb_init_constants_and_likelihood():
    p = E.eb_init(4, np.random.default_rng(0))
    assert abs(float(p['matrices'][0][0, 0, 0]) - (-1.5791)) < 1e-3
    assert abs(float(p['matrices'][3][0, 0, 0]) - (-0.2813)) < 1e-3
    v = torch.arange(-60, 61, dtype=torch.float64).view(1, 1, -1).repeat(4, 1, 1)
    lik = E.eb_likelihood_c1m(p, v, torch.float64)
    s = lik.sum(-1)
    assert torch.all((s > 0.99) & (s < 1.0 + 1e-6)) and float(lik.min()) >= 1e-9
    assert abs(E.eb_aux_loss(p).item() - 0) > 1.0            # quantiles are not yet at the targets
    T = math.log(2 / 1e-9 - 1)
    assert abs(T - 21.4164) < 1e-3


def test_eb_tables_and_quantise_round_trip():
    p = E.eb_init(3, np.random.default_rng(1))
    p['quantiles'][:, 0, 1] = np.array([0.3, -0.2, 0.0], np.float32)
    t = E.eb_tables(p)
    assert t['offset'].tolist() == [-11, -10, -10] or t['offset'].tolist() == [-11, -10, -10]
    for c in range(3):
        L = t['cdf_length'][c]
        assert t['cdf'][c, 0] == 0 and t['cdf'][c, L - 1] == 65536
    x = torch.randn(2, 3, 2, 2, 2) * 5
    sym = E.eb_symbols(p, x)
    xh = E.eb_dequantize(p, sym)
    med = torch.tensor(p['quantiles'][:, 0, 1]).view(1, 3, 1, 1, 1)
    assert torch.all((xh - x).abs() <= 0.5 + 1e-6)
    assert torch.allclose(xh, torch.floor(x + 0.5 - med) + med)


def test_focal_loss_known_answer():
    yt = torch.tensor([1.0, 1.0, 0.0, 0.0])
    yp = torch.tensor([0.5, 5.0, 0.5, 0.0])   # 5.0 clips to .999 ; 0.0 clips to 1e-3
    a, g = 0.75, 2
    expect = -(a * 0.25 * math.log(0.5) + a * (1e-3) ** 2 * math.log(0.999)) \
             - ((1 - a) * 0.25 * math.log(0.5) + (1 - a) * (1e-3) ** 2 * math.log(1 - 1e-3))
    assert abs(float(focal_loss(yt, yp, g, a)) - expect) < 1e-6
