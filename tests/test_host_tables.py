"""Host-side logic of the product (table construction, threshold rounding, block packing, tracing) against the
oracle -- CPU only, no kernels."""
import numpy as np
import pytest
import torch

from oracle import entropy as E
from pcc_geo_cnn_v2_b200 import entropy_models as EM
from pcc_geo_cnn_v2_b200 import model_transforms as MT
from pcc_geo_cnn_v2_b200 import model_types as MTY
from pcc_geo_cnn_v2_b200 import synthetic
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType, PAPER_CONFIGS


def test_gaussian_tables_match_oracle():
    st = EM.make_scale_table()
    assert np.array_equal(st, E.make_scale_table())
    a, b = EM.gaussian_tables(st), E.gc_tables(st)
    for k in ('cdf', 'cdf_length', 'offset'):
        assert np.array_equal(a[k], b[k]), k


def test_entropy_bottleneck_tables_and_aux_loss_match_oracle():
    m = ModelConfigType['c2'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=5))
    eb = m.entropy_bottleneck
    w = eb.get_weights()
    b = E.eb_tables(w)
    for k in ('cdf', 'cdf_length', 'offset'):
        assert np.array_equal(eb.tables[k], b[k]), k
    assert abs(eb.losses[0] - float(E.eb_aux_loss(w, torch.float64))) < 1e-6
    assert eb.updates[0]() is eb.tables


def test_config_table():
    assert list(ModelConfigType.keys()) == ['c1', 'c2', 'c3', 'c3p']        # model_configs.py:16-42
    m = ModelConfigType['c3p'].build()
    assert m.num_filters == 64 and type(m).__name__ == 'CompressionModelV2'
    assert type(m.analysis_transform).__name__ == 'AnalysisTransformProgressiveV2'
    assert ModelConfigType['c1'].build().num_filters == 32
    assert PAPER_CONFIGS['c4'] == dict(model_config='c3p', alpha=0.75, fixed_threshold=True, train_mode='independent')
    assert set(PAPER_CONFIGS) == {'c1', 'c2', 'c3', 'c4', 'c5', 'c6'}
    assert np.array_equal(m.thresholds, np.linspace(0, 1, 256))
    assert np.allclose(m.scale_table, np.exp(np.linspace(np.log(0.11), np.log(256), 64)))


def test_trace_fuses_residual_adds():
    t = MT.AnalysisTransformProgressiveV2(64, data_format='channels_first')
    steps, out = MT.trace(t)
    assert len(steps) == 10 and all(s[0] == 'conv' for s in steps)
    assert [s[4] is not None for s in steps] == [False, False, True] * 3 + [False]
    assert steps[2][4] == steps[0][3]           # residual operand = output of the block's first conv
    blk = MT.AnalysisBlock(4, data_format='channels_first', residual_mode='concat')
    steps, out = MT.trace(blk)
    assert steps[-1][0] == 'concat'


def test_threshold_rounding_is_exact():
    th = np.linspace(0, 1.0, 256)
    t32 = MTY.threshold_f32(th, np.arange(256))
    assert np.all(t32.astype(np.float64) <= th)
    up = np.nextafter(t32, np.float32(np.inf))
    assert np.all(up.astype(np.float64)[:-1] > th[:-1])
    # fp32 x > t32  <=>  x > t64 for the floats around the threshold
    for i in (1, 128, 254):
        for x in (t32[i], up[i], np.nextafter(t32[i], np.float32(-np.inf))):
            assert (x > t32[i]) == (np.float64(x) > th[i])


def test_blocks_to_coords_and_bits_to_points():
    blocks = [np.array([[1, 2, 3], [4, 5, 6]], np.float32), np.zeros((0, 3), np.float32), np.array([[7, 0, 63]], np.float64)]
    c = MTY.blocks_to_coords(blocks)
    assert c.dtype == np.int16 and c.tolist() == [[0, 1, 2, 3], [0, 4, 5, 6], [2, 7, 0, 63]]
    occ = np.zeros((4, 4, 8), bool)
    occ[1, 2, 3] = occ[3, 3, 7] = occ[0, 0, 0] = True
    words = np.packbits(occ.reshape(-1), bitorder='little').view(np.uint32)
    pts = MTY.bits_to_points(words, (4, 4, 8))
    assert pts.dtype == np.float32 and pts.tolist() == [[0, 0, 0], [1, 2, 3], [3, 3, 7]]


def test_synthetic_blocks_are_valid():
    blocks = synthetic.surface_blocks(4, size=32, seed=1)
    for b in blocks:
        assert b.dtype == np.float32 and b.min() >= 0 and b.max() <= 31
        assert len(np.unique(b, axis=0)) == len(b)
        assert 0.005 < len(b) / 32 ** 3 < 0.12


def test_bits_to_points_cpp_matches_numpy():
    from pcc_geo_cnn_v2_b200 import ops
    rng = np.random.default_rng(2)
    occ = rng.random((5, 8, 16, 32)) < 0.07
    occ[3] = False                                            # empty block
    occ[4] = True                                             # full block
    words = np.packbits(occ.reshape(5, -1), axis=1, bitorder='little').view(np.uint32)
    got = ops.bits_to_points(words, (8, 16, 32), threads=3)
    for j in range(5):
        want = np.argwhere(occ[j]).astype(np.float32)
        assert got[j].dtype == np.float32 and np.array_equal(got[j], want)
        assert np.array_equal(MTY.bits_to_points(words[j], (8, 16, 32)), want)


def test_blocks_to_coords_host_helper_matches_numpy():
    """C++ coordinate packer (host half of sparse_to_dense, model_types.py:108-114) against the numpy statement."""
    from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords
    rng = np.random.default_rng(5)

    def ref(blocks):
        rows = [np.concatenate([np.full((len(b), 1), i, np.int16), np.asarray(b)[:, :3].astype(np.int16)], 1) for i, b in enumerate(blocks)]
        return np.concatenate(rows) if rows else np.zeros((0, 4), np.int16)

    b32 = [rng.integers(0, 64, size=(n, 3)).astype(np.float32) for n in (5, 0, 1000, 1)]
    assert np.array_equal(blocks_to_coords(b32, threads=3), ref(b32))
    with_normals = [np.concatenate([b, rng.normal(size=(len(b), 3)).astype(np.float32)], 1) for b in b32]
    assert np.array_equal(blocks_to_coords(with_normals), ref(b32))
    b64 = [b.astype(np.float64) for b in b32]
    assert np.array_equal(blocks_to_coords(b64, threads=2), ref(b32))
    mixed = [b32[0], b64[2], np.asfortranarray(with_normals[2]), rng.integers(0, 64, size=(7, 3))]  # dtype / layout mix
    assert np.array_equal(blocks_to_coords(mixed), ref(mixed))
    assert blocks_to_coords([]).shape == (0, 4)
    with pytest.raises(ValueError):
        blocks_to_coords([np.zeros((4, 2), np.float32)])
