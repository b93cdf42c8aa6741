"""Weight import / export (the role of tf.train.Saver.restore in compress_octree.py:82-92): Keras-style variable names,
suffix matching under arbitrary scopes, shape checks, round trip.  Host-only (no kernels are launched)."""
import numpy as np
import pytest

from pcc_geo_cnn_v2_b200 import synthetic, weights_io
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType


def _same(a, b):
    for k in a:
        if k == 'entropy_bottleneck':
            for kk in a[k]:
                va, vb = a[k][kk], b[k][kk]
                if isinstance(va, list):
                    assert all(np.array_equal(x, y) for x, y in zip(va, vb))
                else:
                    assert np.array_equal(va, vb)
        else:
            for la, lb in zip(a[k], b[k]):
                assert np.array_equal(la['kernel'], lb['kernel'])
                assert (la['bias'] is None) == (lb['bias'] is None) and (la['bias'] is None or np.array_equal(la['bias'], lb['bias']))


@pytest.mark.parametrize('config', ['c1', 'c2', 'c3', 'c3p'])
def test_save_load_round_trip(tmp_path, config):
    m = ModelConfigType[config].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=3))
    path = str(tmp_path / 'w.npz')
    m.save_weights(path)
    m2 = ModelConfigType[config].build()           # fresh model: layers not built yet
    names = m2.load_weights(path)
    _same(m.get_weights(), m2.get_weights())
    assert len(names) == len(set(names)) == len(np.load(path).files)


def test_keras_auto_names_and_scope_insensitive_matching():
    m = ModelConfigType['c3p'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=4))
    names = [n for n, *_ in weights_io.variable_names(m)]
    # creation order analysis -> synthesis -> hyper-analysis -> hyper-synthesis (model_types.py:329-333): 10 + 3 Conv3D, 10 + 3 Conv3DTranspose
    assert names[0] == 'analysis_transform_progressive_v2/analysis_block/conv3d/kernel'
    assert 'analysis_transform_progressive_v2/analysis_block_2/conv3d_8/bias' in names
    assert 'analysis_transform_progressive_v2/conv3d_9/kernel' in names and 'analysis_transform_progressive_v2/conv3d_9/bias' not in names
    assert 'synthesis_transform_progressive_v2/conv3d_transpose_9/bias' in names
    assert 'hyper_analysis_transform/conv3d_12/kernel' in names and 'hyper_synthesis_transform/conv3d_transpose_12/bias' in names
    assert 'entropy_bottleneck/matrix_3' in names and 'entropy_bottleneck/factor_2' in names and 'entropy_bottleneck/quantiles' in names
    sd = weights_io.state_dict(m)
    # a TF1 training checkpoint: other scopes, ':0' suffixes, optimizer slots and counters mixed in
    ckpt = {}
    for k, v in sd.items():
        ckpt['model/' + k.split('/', 1)[1] + ':0' if not k.startswith('entropy') else k] = v
        ckpt[k + '/Adam'] = np.zeros_like(v)
        ckpt[k + '/Adam_1'] = np.zeros_like(v)
    ckpt['beta1_power'] = np.float32(0.5)
    ckpt['global_step'] = np.int64(7)
    m2 = ModelConfigType['c3p'].build()
    m2.load_weights(ckpt)
    _same(m.get_weights(), m2.get_weights())
    bad = dict(ckpt)
    key = next(k for k in bad if k.endswith('conv3d_3/kernel:0'))
    bad[key] = bad[key][..., :-1]
    with pytest.raises(ValueError, match='shape'):
        ModelConfigType['c3p'].build().load_weights(bad)
    del bad[key]
    with pytest.raises(KeyError, match='conv3d_3/kernel'):
        ModelConfigType['c3p'].build().load_weights(bad)
