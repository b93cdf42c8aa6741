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


def test_gpu_threshold_search_equals_the_reference_model_opt():
    """The GPU path against choices made by the REFERENCE's own model_opt.py / pc_metric.py on the same blocks and x_hat
    fields (tests/golden/ref_model_opt.npz): this row's parity is pinned by the reference itself."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_model_opt.npz'))
    size, thr = int(g['size']), g['thresholds']
    opt_metrics, max_deltas = [str(m) for m in g['opt_metrics']], [float(d) for d in g['max_deltas']]
    n = int(g['n_blocks'])
    blocks = [g[f'block{j}'].astype(np.float32) for j in range(n)]
    x_hat = torch.from_numpy(np.stack([g[f'x_hat{j}'] for j in range(n)])[:, None]).cuda()
    coords = torch.from_numpy(blocks_to_coords(blocks)).cuda()
    offsets = np.concatenate([[0], np.cumsum([len(b) for b in blocks])]).astype(np.int64)
    t32 = threshold_f32(thr, np.arange(len(thr)))
    names, best = MO.compute_optimal_thresholds_batch(blocks, x_hat, t32, coords, offsets, opt_metrics=opt_metrics, max_deltas=max_deltas)
    assert names == [str(v) for v in g['names']]
    for j in range(n):
        assert list(best[j]) == list(g[f'best{j}']), (j, list(best[j]), list(g[f'best{j}']))
    # and the sums behind the first choice equal the reference's compute_metrics
    sab, sba, cb = MO.threshold_sums(x_hat, t32, coords, offsets)
    for j in range(n):
        i = int(g[f'best{j}'][0])
        assert [float(sab[j, i]), float(sba[j, i])] == list(g[f'metrics{j}'][:2])
