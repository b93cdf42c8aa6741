"""Container and octree partitioning against golden vectors produced by the REFERENCE's own modules
(tests/golden/make_reference_host_fixtures.py imports /root/reference/src/model_syntax.py and utils/octree_coding.py): the
two rows of this repo whose parity is pinned by the reference itself.  CPU only."""
import io
import os

import numpy as np
import pytest

from pcc_geo_cnn_v2_b200 import model_syntax as MS
from pcc_geo_cnn_v2_b200 import octree_coding as OC

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_host_fixtures.npz'))


@pytest.mark.parametrize('ci', range(int(G['syntax_cases'])))
def test_container_bytes_equal_the_reference(ci):
    res, lvl, n_blocks, n_strings = (int(v) for v in G[f'syntax{ci}_meta'])
    lens, blob = G[f'syntax{ci}_lens'], G[f'syntax{ci}_strings'].tobytes()
    strings, pos = [], 0
    for ln in lens:
        strings.append(blob[pos:pos + int(ln)])
        pos += int(ln)
    data = [(tuple(strings[b * n_strings:(b + 1) * n_strings]), int(G[f'syntax{ci}_thr'][b])) for b in range(n_blocks)]
    got = MS.save_compressed_file(list(G[f'syntax{ci}_binstr']), data, res, lvl)
    assert got == G[f'syntax{ci}_blob'].tobytes()
    r2, l2, binstr2, blocks2 = MS.load_compressed_file(io.BytesIO(got))
    assert (int(r2), int(l2)) == (res, lvl) and np.array_equal(binstr2, G[f'syntax{ci}_binstr'])
    assert [(tuple(s), int(t)) for s, t in blocks2] == data


def test_container_limits_raise_like_the_reference():
    with pytest.raises(AssertionError):
        MS.save_compressed_file([1], [((b'x' * 70000,), 0)], 64, 1)
    with pytest.raises(AssertionError):
        MS.load_compressed_file(io.BytesIO(MS.save_compressed_file([1], [((b'ab',), 3)], 64, 1) + b'!'))


@pytest.mark.parametrize('ci', range(int(G['oct_cases'])))
def test_octree_partition_equals_the_reference(ci):
    res, level, cols = (int(v) for v in G[f'oct{ci}_meta'])
    pts = G[f'oct{ci}_points']
    blocks, binstr = OC.partition_octree(pts, [0, 0, 0], [res] * 3, level)
    assert [len(b) for b in blocks] == list(G[f'oct{ci}_block_len'])
    assert np.array_equal(np.vstack(blocks), G[f'oct{ci}_blocks']) and blocks[0].dtype == np.float64
    assert list(binstr) == list(G[f'oct{ci}_binstr'])
    dep = OC.departition_octree(blocks, binstr, [0, 0, 0], [res] * 3, level)
    if len(G[f'oct{ci}_departitioned']):       # (the reference's departition raises on one-level trees)
        assert np.array_equal(np.vstack(dep), G[f'oct{ci}_departitioned'])
    assert sorted(map(tuple, np.vstack(dep))) == sorted(map(tuple, pts))


def test_octree_deep_levels_round_trip():
    """geo_level < 2*level: the reference's truncated sort key breaks its own departition; Morton order round-trips."""
    rng = np.random.default_rng(0)
    pts = np.unique(rng.integers(0, 256, size=(3000, 3)), axis=0).astype(np.float64)
    for level in (5, 8):
        blocks, binstr = OC.partition_octree(pts, [0, 0, 0], [256] * 3, level)
        dep = OC.departition_octree(blocks, binstr, [0, 0, 0], [256] * 3, level)
        assert sorted(map(tuple, np.vstack(dep))) == sorted(map(tuple, pts))
        assert all((b[:, :3] >= 0).all() and (b[:, :3] < 256 // 2 ** level).all() for b in blocks)


def test_oracle_threshold_search_equals_the_reference_model_opt():
    """oracle/model_opt.py (kd-tree restatement) against choices made by the reference's own model_opt.py / pc_metric.py
    (tests/golden/make_reference_model_opt_fixture.py): pins the oracle that the GPU path is tested against."""
    from oracle import model_opt as OMO
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_model_opt.npz'))
    size, thr = int(g['size']), g['thresholds']
    opt_metrics, max_deltas = [str(m) for m in g['opt_metrics']], [float(d) for d in g['max_deltas']]
    for j in range(int(g['n_blocks'])):
        names, best = OMO.compute_optimal_thresholds(g[f'block{j}'].astype(np.float32), g[f'x_hat{j}'], thr, size,
                                                     opt_metrics=opt_metrics, max_deltas=max_deltas)
        assert names == [str(n) for n in g['names']]
        assert list(best) == list(g[f'best{j}']), j
        m = OMO.compute_metrics(g[f'block{j}'].astype(np.float64), np.argwhere(g[f'x_hat{j}'] > thr[best[0]]), size - 1)
        assert np.allclose([m['d1_sum_AB'], m['d1_sum_BA'], m['d1_mse'], m['d1_psnr']], g[f'metrics{j}'], rtol=1e-6)


def test_scale_table_and_thresholds_equal_the_reference_constructors():
    """CompressionModel.thresholds / CompressionModelV2.scale_table (model_types.py:181,324) as the reference builds them."""
    from pcc_geo_cnn_v2_b200 import ModelConfigType
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_block_loops.npz'))
    m = ModelConfigType['c3p'].build()
    assert np.array_equal(np.asarray(m.thresholds), g['thresholds'])
    assert np.array_equal(np.asarray(m.scale_table, np.float64), g['scale_table'])


def test_ply_reader_and_writer(tmp_path):
    """pc_io: binary / ASCII PLY vertex elements (the reference's dataset layout: binary float x y z) and the writer's layout
    (float32 x y z + uint8 colours, pc_io.py:16-26,49-51)."""
    from pcc_geo_cnn_v2_b200 import pc_io
    rng = np.random.default_rng(0)
    pts = rng.integers(0, 64, size=(100, 3)).astype(np.float32)
    p = tmp_path / 'a.ply'
    pc_io.write_pc(str(p), pts)
    assert np.array_equal(pc_io.load_pc(str(p)), pts)
    cols = np.concatenate([pts, rng.integers(0, 256, size=(100, 3))], axis=1)
    pc_io.write_pc(str(p), cols)
    c = pc_io.read_ply(str(p))
    assert list(c) == ['x', 'y', 'z', 'red', 'green', 'blue'] and c['red'].dtype == np.uint8 and np.array_equal(c['blue'], cols[:, 5])
    ascii_ply = ('ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n'
                 'property float nx\nproperty float ny\nproperty float nz\nelement face 0\nproperty list uchar int vertex_indices\n'
                 'end_header\n1 2 3 0 0 1\n4 5 6 0 1 0\n7 8 9 1 0 0\n').encode()
    q = tmp_path / 'b.ply'
    q.write_bytes(ascii_ply)
    assert np.array_equal(pc_io.load_pc(str(q)), [[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    assert np.array_equal(pc_io.load_normals(str(q)), [[0, 0, 1], [0, 1, 0], [1, 0, 0]])
    blocks = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'modelnet_blocks.npz'))
    header = b'ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n'
    b0 = blocks['block0'].astype('<f4')
    assert np.array_equal(pc_io.load_pc(header % len(b0) + b0.tobytes()) if False else pc_io.read_ply(header % len(b0) + b0.tobytes())['x'], b0[:, 0])
    p_min, p_max, shape = pc_io.get_shape_data(64, 'channels_first')
    assert list(shape) == [1, 64, 64, 64] and list(pc_io.get_shape_data(64, 'channels_last')[2]) == [64, 64, 64, 1]


def test_d2_metrics_and_threshold_search_equal_the_reference():
    """pc_metric.compute_metrics incl. the D2 branch / assign_attr, and model_opt.compute_optimal_thresholds with normals,
    against the reference's own pc_metric.py (numba assign_attr) and model_opt.py (tests/golden/make_reference_d2_fixture.py)."""
    from pcc_geo_cnn_v2_b200 import model_opt as MO
    from pcc_geo_cnn_v2_b200 import pc_metric as PM
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_d2.npz'))
    size, thr = int(g['size']), g['thresholds']
    opt_metrics, max_deltas = [str(m) for m in g['opt_metrics']], [float(d) for d in g['max_deltas']]
    for j in range(int(g['n_blocks'])):
        block, x_hat = g[f'block{j}'], g[f'x_hat{j}']
        names, best = MO.compute_optimal_thresholds(block, x_hat, thr, size, normals=block[:, 3:], opt_metrics=opt_metrics,
                                                    max_deltas=max_deltas)
        assert names == [str(n) for n in g['names']] and list(best) == list(g[f'best{j}'])
        m = PM.compute_metrics(block[:, :3], np.argwhere(x_hat > thr[20]).astype('float32'), size - 1, p1_n=block[:, 3:])
        got = np.array([m[str(k)] for k in g['metric_keys']], np.float64)
        assert np.allclose(got, g[f'metrics{j}'], rtol=1e-9, atol=0), (got, g[f'metrics{j}'])
