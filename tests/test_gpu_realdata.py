"""Codec path on the reference's own data (a 24-block sample of ModelNet40_200_pc512_oct3_4k, tests/golden/
modelnet_blocks.npz) and on the edge cases of the block loops: ragged batches, empty inputs, 128^3 blocks (1024-res clouds
at octree level 3), coded size against the entropy estimate."""
import os

import numpy as np
import pytest
import torch

from oracle.model import OracleModel, sparse_to_dense
from pcc_geo_cnn_v2_b200 import ops, synthetic
from pcc_geo_cnn_v2_b200.entropy_models import GaussianConditional
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(params=['host', 'device'], autouse=True)
def _entropy_coder(request, monkeypatch):
    """every test of this file runs with the range coder in the host workers and on the GPU (models read the switch when built)"""
    monkeypatch.setenv('PCCGEO_DEVICE_CODER', '1' if request.param == 'device' else '0')
    yield request.param


def _modelnet_blocks():
    g = np.load(os.path.join(GOLDEN, 'modelnet_blocks.npz'))
    return [g[f'block{i}'].astype(np.float32) for i in range(len(g['names']))]


def _model(config='c3p', seed=42, **kw):
    m = ModelConfigType[config].build(**kw)
    m.set_weights(synthetic.trained_like_weights(m, seed=seed))
    return m


def _ideal_bits(sym, idx, tab):
    """Shannon code length of the symbols under the coder's 16-bit tables, with the escape mechanism of
    unbounded_index_range_encode (escape slot + 4-bit uniform chunks)."""
    cdf, cl, off = tab['cdf'], tab['cdf_length'], tab['offset']
    maxv = cl[idx] - 2
    value = sym.astype(np.int64) - off[idx]
    esc = (value < 0) | (value >= maxv)
    v = np.where(esc, maxv, value)
    freq = cdf[idx, v + 1] - cdf[idx, v]
    bits = np.log2(65536.0 / freq).sum()
    for val, mx in zip(value[esc], maxv[esc]):
        overflow = -2 * val - 1 if val < 0 else 2 * (val - mx)
        widths = 0
        while (int(overflow) >> (4 * widths)) != 0:
            widths += 1
        bits += 4 * (widths // 15 + 1 + widths)
    return float(bits)


def test_modelnet_blocks_round_trip_and_rate():
    blocks = _modelnet_blocks()
    assert len(blocks) == 24 and min(map(len, blocks)) == 3571 and max(map(len, blocks)) == 38933
    m = _model(batch_size=8)          # 3 pipelined batches
    assert m.device_coder == (os.environ['PCCGEO_DEVICE_CODER'] == '1')
    m.coder_group_blocks = 16         # 2 coder groups
    m.compress((1, 1, 64, 64, 64))
    data, meta, _ = m.compress_blocks(None, blocks, None, None, 512, 3, fixed_threshold=True)
    dec, _ = m.decompress_blocks(None, data[0], (64, 64, 64))
    assert len(dec) == 24
    for a, b in zip(meta[0]['x_hat_list'], dec):
        assert np.array_equal(a, b)           # decoder reproduces the encoder's point sets exactly
    # batch composition must not matter (kernels are deterministic per sample): re-encode with another batch size
    m2 = _model(batch_size=5)
    m2.compress((1, 1, 64, 64, 64))
    data2, _, _ = m2.compress_blocks(None, blocks, None, None, 512, 3, fixed_threshold=True)
    assert [d[0] for d in data2[0]] == [d[0] for d in data[0]]
    # coded size vs the ideal code length of the quantised tables the coder uses (sum log2(2^16 / freq), escapes included):
    # the arithmetic coder must be within 0.1 % + termination bytes of it
    x = ops.densify(torch.from_numpy(blocks_to_coords(blocks[:8])).cuda(), 8, 64, 64, 64)
    dev = m._encode_device(x)
    cb = GaussianConditional(dev['sigma_hat'], m.scale_table)
    ysym, idx, zsym = dev['y_sym'].cpu().numpy(), dev['indexes'].cpu().numpy(), dev['z_sym'].cpu().numpy()
    zidx = np.broadcast_to(np.arange(64).reshape(1, 64, 1, 1, 1), zsym.shape)
    for j in range(8):
        for sym, ind, tab, coded in ((ysym[j], idx[j], cb.tables, data[0][j][0][0]), (zsym[j], zidx[j], m.entropy_bottleneck.tables, data[0][j][0][1])):
            ideal = _ideal_bits(sym.reshape(-1), ind.reshape(-1), tab)
            assert 8 * len(coded) <= 1.001 * ideal + 48, (8 * len(coded), ideal)
            assert 8 * len(coded) >= ideal - 40, (8 * len(coded), ideal)      # trailing zero bytes are stripped
    # and the likelihood model agrees with the tables to first order (scales are quantised to 64 levels -> a few percent)
    _, lik_y = cb(dev['y'], training=False)
    est = float(-torch.log2(lik_y[:8]).sum())
    coded = sum(8 * len(data[0][j][0][0]) for j in range(8))
    assert 0.8 * est < coded < 1.25 * est, (coded, est)


def test_modelnet_block_against_oracle():
    blocks = _modelnet_blocks()[2:4]
    m = _model()
    o = OracleModel('c3p')
    w = m.get_weights()
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, w['entropy_bottleneck'])
    x = ops.densify(torch.from_numpy(blocks_to_coords(blocks)).cuda(), 2, 64, 64, 64)
    dev = m._encode_device(x)
    for j, b in enumerate(blocks):
        xo = sparse_to_dense(b, (1, 1, 64, 64, 64))
        assert np.array_equal(x[j:j + 1].cpu().numpy(), xo)
        t = o.analyse(xo)
        rel = float((dev['y'][j].cpu() - t['y'][0]).abs().max() / t['y'].abs().max())
        assert rel < 3e-5, rel                                      # bf16x3: fp32-class latents
        flips = int((dev['y_sym'][j].cpu() != t['y_symbols'][0]).sum())
        assert flips <= 8, flips                                    # of 32 768 symbols, each by one step
        zflip = int((dev['z_sym'][j].cpu() != t['z_symbols'][0]).sum())
        assert zflip <= 2
        if zflip == 0:
            xh = o.synthesise(t['y_hat'])
            mism = int(((dev['x_hat'][j, 0].cpu() > 0.50196) != (xh[0, 0] > 0.50196)).sum())
            assert mism <= 4 * max(flips, 1) + 8, mism              # occupancy differs only around flipped latents


def test_empty_and_degenerate_inputs():
    m = _model(batch_size=4)
    m.compress((1, 1, 64, 64, 64))
    data, meta, dbg = m.compress_blocks(None, [], None, None, 64, 0, fixed_threshold=True)
    assert data == [[]] and dbg == []
    assert m.decompress_blocks(None, [], (64, 64, 64)) == ([], [])
    empty = np.zeros((0, 3), np.float32)
    full_row = np.stack([np.zeros(64), np.zeros(64), np.arange(64)], 1).astype(np.float32)
    data, meta, _ = m.compress_blocks(None, [empty, full_row, empty], None, None, 64, 0, fixed_threshold=True)
    dec, _ = m.decompress_blocks(None, data[0], (64, 64, 64))
    assert len(dec) == 3
    for a, b in zip(meta[0]['x_hat_list'], dec):
        assert np.array_equal(a, b)
    assert data[0][0][0] == data[0][2][0]      # the two empty blocks code identically


def test_128_cube_blocks():
    """1024-resolution clouds at octree level 3 give 128^3 blocks (latents 16^3 / 8^3): the nets are fully convolutional."""
    m = _model(batch_size=2)
    blocks = synthetic.surface_blocks(3, size=128, seed=4)
    m.compress((1, 1, 128, 128, 128))
    data, meta, _ = m.compress_blocks(None, blocks, None, None, 1024, 3, fixed_threshold=True)
    dec, _ = m.decompress_blocks(None, data[0], (128, 128, 128))
    for a, b in zip(meta[0]['x_hat_list'], dec):
        assert np.array_equal(a, b)
        assert a.shape[1] == 3 and (len(a) == 0 or a.max() <= 127)
