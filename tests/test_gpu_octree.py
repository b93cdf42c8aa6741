"""Octree partition on the GPU (csrc/octree.cu) against the golden vectors produced by the REFERENCE's own
utils/octree_coding.py (tests/golden/ref_host_fixtures.npz, make_reference_host_fixtures.py): same blocks, block order,
in-block point order, local coordinates and occupancy bytes; and the device-resident form drives compress_blocks to the same
bitstream as host-partitioned blocks."""
import os

import numpy as np
import pytest
import torch

from pcc_geo_cnn_v2_b200 import ModelConfigType, codec, synthetic
from pcc_geo_cnn_v2_b200 import octree_coding as OC

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_host_fixtures.npz'))


@pytest.mark.parametrize('ci', range(int(G['oct_cases'])))
def test_gpu_partition_equals_the_reference(ci):
    res, level, cols = (int(v) for v in G[f'oct{ci}_meta'])
    pts = G[f'oct{ci}_points']
    if level < 1 or level > 6:
        pytest.skip('GPU partition covers 1 <= level <= 6')
    blocks, binstr = OC.partition_octree_gpu(pts, [0, 0, 0], [res] * 3, level)
    assert [len(b) for b in blocks] == list(G[f'oct{ci}_block_len'])
    assert np.array_equal(np.vstack(blocks), G[f'oct{ci}_blocks']) and blocks[0].dtype == np.float64
    assert list(binstr) == list(G[f'oct{ci}_binstr'])
    dev, binstr2 = OC.partition_octree_gpu(pts, [0, 0, 0], [res] * 3, level, device=True)
    assert list(binstr2) == list(binstr) and len(dev) == len(blocks) and list(dev.counts) == [len(b) for b in blocks]
    c = dev.coords.cpu().numpy()
    assert np.array_equal(c[:, 1:].astype(np.float64), np.vstack(blocks)[:, :3])
    assert np.array_equal(c[:, 0], np.repeat(np.arange(len(blocks)), [len(b) for b in blocks]))


def test_gpu_partition_large_cloud_matches_the_host_path():
    rng = np.random.default_rng(1)
    res, level = 1024, 4
    u = rng.random((400000, 2))
    pts = np.stack([u[:, 0] * (res - 1), (np.sin(u[:, 0] * 7) * 0.25 + 0.5) * (res - 1), u[:, 1] * (res - 1)], 1)
    pts = np.unique(pts.astype(np.int64), axis=0).astype(np.float64)
    pts = pts[rng.permutation(len(pts))]                  # unordered input: the sort has to be stable
    want, wb = OC.partition_octree(pts, [0, 0, 0], [res] * 3, level)
    got, gb = OC.partition_octree_gpu(pts, [0, 0, 0], [res] * 3, level)
    assert list(gb) == list(wb) and len(got) == len(want)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    with pytest.raises(ValueError):
        OC.partition_octree_gpu(pts * 2.0, [0, 0, 0], [res] * 3, level)


def test_device_blocks_give_the_same_bitstream():
    res, level, bs = 256, 2, 64
    rng = np.random.default_rng(4)
    u = rng.random((60000, 2))
    pts = np.stack([u[:, 0] * (res - 1), (np.sin(u[:, 0] * 5) * 0.3 + 0.5) * (res - 1) * (0.3 + 0.7 * u[:, 1]), u[:, 1] * (res - 1)], 1)
    pts = np.unique(pts.astype(np.int64), axis=0).astype(np.float64)
    m = ModelConfigType['c3p'].build(batch_size=8)
    m.set_weights(synthetic.trained_like_weights(m, seed=3, output_bias=-0.45))
    for coder in (False, True):
        m.device_coder = coder
        blobs_gpu, _ = codec.compress_point_cloud(m, pts, res, level, fixed_threshold=True, partition='gpu')
        blobs_host, _ = codec.compress_point_cloud(m, pts, res, level, fixed_threshold=True, partition='host')
        assert blobs_gpu == blobs_host
    blobs_a, meta_a = codec.compress_point_cloud(m, pts, res, level, fixed_threshold=False, partition='gpu')
    blobs_b, meta_b = codec.compress_point_cloud(m, pts, res, level, fixed_threshold=False, partition='host')
    assert blobs_a == blobs_b
    dec = codec.decompress_point_cloud(m, blobs_gpu[0])
    assert dec.shape[1] == 3 and len(dec) > 1000
