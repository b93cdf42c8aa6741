"""Per-block threshold optimisation on the GPU (csrc/threshold_opt.cu + model_opt.py) against brute-force numpy sums (exact
integers) and against the oracle's kd-tree restatement of the reference (src/model_opt.py:9-77, pc_metric.py:76-108)."""
import numpy as np
import pytest
import torch

from oracle import model_opt as OMO
from pcc_geo_cnn_v2_b200 import ModelConfigType, synthetic
from pcc_geo_cnn_v2_b200 import model_opt as MO
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords, threshold_f32

pytestmark = pytest.mark.gpu


def _brute(block, xh, t32):
    a = np.asarray(block, np.int64)[:, :3]
    sab, sba, cb = [], [], []
    for t in t32:
        b = np.argwhere(xh > t).astype(np.int64)
        cb.append(len(b))
        if len(b) == 0:
            sab.append(-1); sba.append(0)
            continue
        d = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
        sab.append(int(d.min(1).sum())); sba.append(int(d.min(0).sum()))
    return np.array(sab), np.array(sba), np.array(cb)


@pytest.mark.parametrize('shape', [(16, 16, 16), (8, 24, 16)])
def test_threshold_sums_are_exact(shape):
    rng = np.random.default_rng(7)
    n, t = 3, 24
    t32 = threshold_f32(np.linspace(0, 1.0, t), np.arange(t))
    blocks, xs = [], []
    for j in range(n):
        occ = rng.random(shape) < 0.03
        occ[tuple(s // 2 for s in shape)] = True
        blocks.append(np.argwhere(occ).astype(np.float32))
        field = rng.random(shape).astype(np.float32) ** (2 + j)          # few voxels near 1: high thresholds empty out
        if j == 2:
            field = np.minimum(field, 0.6)                                # B_i empty from some threshold on
        xs.append(field)
    x_hat = torch.from_numpy(np.stack(xs)[:, None]).cuda()
    coords = torch.from_numpy(blocks_to_coords(blocks)).cuda()
    offsets = np.concatenate([[0], np.cumsum([len(b) for b in blocks])]).astype(np.int64)
    sab, sba, cb = MO.threshold_sums(x_hat, t32, coords, offsets)
    for j in range(n):
        want = _brute(blocks[j], xs[j], t32)
        assert np.array_equal(cb[j], want[2]) and np.array_equal(sba[j], want[1]) and np.array_equal(sab[j], want[0]), j


def test_optimal_thresholds_match_the_oracle_kdtree_search():
    size = 32
    m = ModelConfigType['c3p'].build(batch_size=2)
    m.set_weights(synthetic.trained_like_weights(m, seed=5, output_bias=-0.45))
    m.compress((1, 1, size, size, size))
    blocks = synthetic.surface_blocks(3, size=size, seed=31)
    _, x_hat, _ = m.encode_blocks(blocks, keep_x_hat=True)
    opt_metrics, max_deltas = ('d1_mse', 'd1_sum_mean', 'd1_mse_AB'), (np.inf, 1.5)
    idx, names = m._optimal_thresholds(blocks, x_hat, size, False, opt_metrics, max_deltas)
    xh = np.clip(x_hat[:, 0].cpu().numpy(), 0, 1)
    t32 = threshold_f32(m.thresholds, np.arange(len(m.thresholds)))
    for j, b in enumerate(blocks):
        want_names, want = OMO.compute_optimal_thresholds(b, xh[j], t32, size, opt_metrics=opt_metrics, max_deltas=max_deltas)
        assert names == want_names
        assert list(idx[j]) == list(want), (j, list(idx[j]), want)
    assert len({int(v) for v in idx.ravel()}) > 1   # not a degenerate case
    # the public entry point with adaptive thresholds: the decoder reproduces the encoder's chosen point sets
    dl, meta, _ = m.compress_blocks(None, blocks, None, None, size, 0, opt_metrics=('d1_mse',), max_deltas=(np.inf,), fixed_threshold=False)
    assert [t for _, t in dl[0]] == [int(v) for v in idx[:, 0]]
    dec, _ = m.decompress_blocks(None, dl[0], (size, size, size))
    for a, b in zip(meta[0]['x_hat_list'], dec):
        assert np.array_equal(a, b)
