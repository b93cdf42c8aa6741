"""Entropy-model / voxel kernels against the oracle, through the C ABI.  Integer outputs (symbols, indexes, packed
bits, counts) are bit-exact; likelihoods within 1e-5 relative (north star: 1e-4); sums within 1e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import entropy as E
from oracle.model import focal_loss as oracle_focal_loss
from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200 import entropy_models as EM
from pcc_geo_cnn_v2_b200.focal_loss import focal_loss

pytestmark = pytest.mark.gpu


def _eb(channels, seed):
    rng = np.random.default_rng(seed)
    eb = EM.EntropyBottleneck(data_format='channels_first', seed=seed)
    eb.build(channels)
    w = eb.get_weights()
    w['factors'] = [rng.uniform(-0.3, 0.3, size=f.shape).astype(np.float32) for f in w['factors']]
    w['quantiles'][:, 0, 1] = rng.uniform(-0.4, 0.4, size=channels).astype(np.float32)
    eb.set_weights(w)
    return eb, w


def test_eb_quantize_bit_exact_and_likelihood():
    eb, w = _eb(64, 0)
    x = torch.randn(3, 64, 4, 4, 4) * 6
    sym, xh = eb.quantize(x.cuda())
    assert torch.equal(sym.cpu(), E.eb_symbols(w, x))
    assert torch.equal(xh.cpu(), E.eb_dequantize(w, E.eb_symbols(w, x)))
    assert torch.equal(ops.eb_dequantize(sym, eb.device_params()), xh)
    for training in (False, True):
        noise = torch.rand_like(x) - 0.5
        vt, lik = eb(x.cuda(), training=training, noise=noise.cuda())
        ov, ol = E.eb_forward(w, x, training, noise, torch.float64)
        assert float((vt.cpu().double() - ov).abs().max()) < 1e-5
        rel = ((lik.cpu().double() - ol).abs() / ol).max()
        assert float(rel) < 1e-4, float(rel)
        s = eb.log_likelihood_sum(vt)
        assert abs(float(s[0]) - float(torch.log(ol).sum())) < 1e-5 * abs(float(torch.log(ol).sum()))


def test_eb_likelihood_sum_at_training_batch_size():
    """tr_train.py's batch 32 x 64 channels (model_types.py:333-353) exceeds one partial-sum slot per (sample, channel): the
    kernel then walks several samples per block.  Sum of ln p against the oracle in float64."""
    eb, w = _eb(64, 3)
    x = torch.randn(40, 64, 4, 4, 4) * 4
    s = eb.log_likelihood_sum(x.cuda())
    _, ol = E.eb_forward(w, x, True, torch.zeros_like(x), torch.float64)
    want = float(torch.log(ol).sum())
    assert abs(float(s[0]) - want) < 1e-5 * abs(want), (float(s[0]), want)


def test_gc_quantize_bit_exact_and_likelihood():
    st = EM.make_scale_table()
    rng = np.random.default_rng(1)
    y = torch.from_numpy((rng.normal(size=(2, 64, 8, 8, 8)) * 8).astype(np.float32))
    y.view(-1)[:8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 3.4999, -2.5, 0.0])   # ties: round half to even
    sigma = torch.from_numpy(np.exp(rng.uniform(np.log(0.01), np.log(400), size=y.shape)).astype(np.float32))
    sigma.view(-1)[:64] = torch.from_numpy(st.astype(np.float32))                     # exactly on the table entries
    gc = EM.GaussianConditional(sigma.cuda(), st)
    sym, yh, idx = gc.quantize(y.cuda())
    assert torch.equal(sym.cpu(), E.gc_symbols(y))
    assert torch.equal(yh.cpu(), torch.round(y))
    assert torch.equal(idx.cpu(), E.gc_indexes(sigma, st))
    assert torch.equal(gc.indexes(), idx)
    noise = torch.rand_like(y) - 0.5
    for training in (False, True):
        v, lik = gc(y.cuda(), training=training, noise=noise.cuda())
        ov, ol = E.gc_forward(y, sigma, st, training, noise, torch.float64)
        assert float((v.cpu().double() - ov).abs().max()) < 1e-5
        got = lik.cpu().double()
        rel = ((got - ol).abs() / ol).max()       # north star: within 1e-4 relative (kernel: double erfc, rounded once)
        assert float(rel) < 1e-5, float(rel)
        tot = float(torch.log(ol).sum())
        assert abs(float(gc.log_likelihood_sum(v)[0]) - tot) < 2e-5 * abs(tot)


def test_codec_round_trip_through_entropy_models():
    eb, w = _eb(32, 3)
    z = torch.randn(5, 32, 2, 2, 2) * 9      # beyond the +-10 table: escape codes
    strings = eb.compress(z.cuda())
    assert len(strings) == 5
    zh = eb.decompress(strings, (32, 2, 2, 2), channels=32)
    assert torch.equal(zh.cpu(), E.eb_dequantize(w, E.eb_symbols(w, z)))
    st = EM.make_scale_table()
    sigma = (torch.rand(5, 32, 4, 4, 4) * 20).cuda()
    y = (torch.randn(5, 32, 4, 4, 4) * 10).cuda()
    gc = EM.GaussianConditional(sigma, st)
    ys = gc.compress(y)
    gc2 = EM.GaussianConditional(sigma, st)
    assert torch.equal(gc2.decompress(ys), torch.round(y))
    assert 'decompress/indexes' in gc2.dbg_dec


def test_densify_threshold_pack_and_focal_loss():
    rng = np.random.default_rng(5)
    n, d = 3, 32
    pts = [np.unique(rng.integers(0, d, size=(500, 3)), axis=0) for _ in range(n)]
    coords = np.concatenate([np.concatenate([np.full((len(p), 1), j), p], 1) for j, p in enumerate(pts)]).astype(np.int16)
    x = ops.densify(torch.from_numpy(coords).cuda(), n, d, d, d)
    want = np.zeros((n, 1, d, d, d), np.float32)
    for j, p in enumerate(pts):
        want[j, 0, p[:, 0], p[:, 1], p[:, 2]] = 1
    assert np.array_equal(x.cpu().numpy(), want)
    assert float(ops.densify(None, 2, 8, 8, 8).sum()) == 0.0

    xh = torch.from_numpy(rng.uniform(-0.2, 1.6, size=(n, 1, d, d, d)).astype(np.float32))
    thr = np.array([0.3, 0.50196, 1.0], np.float32)
    bits, counts = ops.threshold_pack(xh.cuda(), torch.from_numpy(thr).cuda())
    occ = np.minimum(xh.numpy()[:, 0], 1.0) > thr[:, None, None, None]
    got = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder='little').reshape(n, d, d, d).astype(bool)
    assert np.array_equal(got, occ) and counts.cpu().tolist() == occ.reshape(n, -1).sum(1).tolist()
    assert counts[2] == 0   # threshold 1.0 = "emit nothing" sentinel after the clip

    xt = torch.from_numpy(want)
    xp = torch.from_numpy(rng.uniform(-0.1, 1.3, size=want.shape).astype(np.float32))
    for gamma, alpha in ((2, 0.75), (2, 0.9), (1.5, 0.5)):
        o = float(oracle_focal_loss(xt.double(), xp.double(), gamma, alpha))
        g = float(focal_loss(xt.cuda(), xp.cuda(), gamma, alpha))
        assert abs(g - o) < 2e-5 * abs(o)
