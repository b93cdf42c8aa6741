"""Golden vectors for the HOST logic of the block loops, produced by the REFERENCE's own CompressionModelV2.compress_blocks
/ decompress_blocks / select_best_per_opt_metric / scale table (src/model_types.py:128-238,313-325) in the build container:

    python tests/golden/make_reference_block_loop_fixture.py   ->  tests/golden/ref_block_loops.npz

TensorFlow / tensorflow-compression / pyntcloud are absent: their modules are stubbed (nothing of them is executed by these
functions -- the networks are replaced by a fake `sess.run` that returns prescribed x_hat volumes and strings), and
cKDTree.query's old `n_jobs=` keyword is mapped to `workers=`."""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import scipy.spatial
from scipy.ndimage import gaussian_filter


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in ('tensorflow', 'tensorflow_core', 'tensorflow_compression', 'pyntcloud'):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = MagicMock()
        m.__name__, m.__path__, m.__spec__ = spec.name, [], spec
        return m

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Finder())


class _Tree(scipy.spatial.cKDTree):
    def query(self, x, k=1, eps=0, p=2, distance_upper_bound=np.inf, n_jobs=None, workers=1):
        return super().query(x, k=k, eps=eps, p=p, distance_upper_bound=distance_upper_bound, workers=workers if n_jobs is None else n_jobs)


scipy.spatial.cKDTree = _Tree
import scipy.spatial.ckdtree as _legacy  # noqa: E402
_legacy.cKDTree = _Tree
sys.path.insert(0, '/root/reference/src')
import model_types as RMT  # noqa: E402
import model_opt as RMO  # noqa: E402
from utils import octree_coding as RO  # noqa: E402
from utils import pc_metric as RPM  # noqa: E402
RPM.cKDTree = RMT.cKDTree = RMO.cKDTree = _Tree

res, level, bs = 128, 2, 32
rng = np.random.default_rng(2024)
u = rng.random((9000, 2))
pts = np.stack([u[:, 0] * (res - 1), (np.sin(u[:, 0] * 4) * 0.3 + 0.5) * (res - 1) * (0.4 + 0.6 * u[:, 1]), u[:, 1] * (res - 1)], 1)
pts = np.unique(pts.astype(np.int64), axis=0).astype(np.float64)
pts = pts[(pts[:, 0] < 64) & (pts[:, 2] < 64)]   # at most 16 occupied blocks: keeps the fixture small
blocks, binstr = RO.partition_octree(pts, [0, 0, 0], [res] * 3, level)

# prescribed network outputs per block: blurred occupancy + noise (unclipped, as the graph returns it) and two strings
x_hats, strings = [], []
for j, b in enumerate(blocks):
    occ = np.zeros((bs,) * 3, np.float32)
    occ[tuple(b[:, :3].astype(int).T)] = 1
    f = gaussian_filter(occ, 0.7 + 0.05 * (j % 5)) * (2.2 + 0.1 * (j % 3)) + rng.normal(size=occ.shape).astype(np.float32) * 0.04 - 0.02
    x_hats.append(f.astype(np.float16).astype(np.float32))   # float16-representable values: stored losslessly as float16
    strings.append((rng.integers(0, 256, size=int(rng.integers(1, 40))).astype(np.uint8).tobytes(),
                    rng.integers(0, 256, size=int(rng.integers(1, 12))).astype(np.uint8).tobytes()))


class FakeSession:
    def __init__(self):
        self.i = 0

    def run(self, fetches, feed_dict=None):
        j = self.i
        self.i += 1
        if len(fetches) == 3:   # compress: [strings, x_hat, debug]
            return [np.array([s]) for s in strings[j]], x_hats[j][None, None], None
        return x_hats[j][None, None], None   # decompress: [x_hat, debug]


m = RMT.CompressionModelV2(num_filters=64)
class _Placeholder:   # stands for the tf.placeholder self.x (hashable, has .shape)
    shape = (1, 1, bs, bs, bs)


m.x = _Placeholder()
m.strings, m.x_hat, m.debug_tensors = 'strings', 'x_hat', None
m.x_shape_t, m.strings_t = 'x_shape', ['y_string', 'z_string']
out = {'res': np.array(res), 'level': np.array(level), 'bs': np.array(bs), 'points': pts, 'binstr': np.array(binstr, np.int64),
       'block_len': np.array([len(b) for b in blocks], np.int64), 'blocks': np.vstack(blocks),
       'x_hats': np.stack(x_hats).astype(np.float16), 'scale_table': m.scale_table, 'thresholds': m.thresholds,
       'str_lens': np.array([len(s) for st in strings for s in st], np.int64),
       'str_bytes': np.frombuffer(b''.join(s for st in strings for s in st), np.uint8)}
for tag, kw in (('fixed', dict(fixed_threshold=True)),
                ('adaptive', dict(fixed_threshold=False, opt_metrics=['d1_mse', 'd1_sum_mean'], max_deltas=[np.inf, 1.3]))):
    data_list, metadata, _ = m.compress_blocks(FakeSession(), blocks, binstr, pts, res, level, **kw)
    assert len(metadata) == 1
    md = metadata[0]
    out[f'{tag}_idx'] = np.array(md['idx'])
    out[f'{tag}_thr'] = np.array([int(t) for _, t in data_list[0]], np.int64)
    out[f'{tag}_metrics'] = np.array([md['metrics'][k] for k in ('d1_sum_AB', 'd1_sum_BA', 'd1_mse', 'd1_psnr')], np.float64)
    out[f'{tag}_pts_len'] = np.array([len(p) for p in md['x_hat_list']], np.int64)
    out[f'{tag}_pts'] = np.vstack(md['x_hat_list']).astype(np.float32)
    out[f'{tag}_full'] = md['blocks_full'].astype(np.float32)
    assert [s for s, _ in data_list[0]] == [list(s) for s in strings] or [tuple(s) for s, _ in data_list[0]] == strings
    dec, _ = m.decompress_blocks(FakeSession(), [(s, t) for s, t in data_list[0]], [bs] * 3)
    assert all(np.array_equal(a, b) for a, b in zip(dec, md['x_hat_list']))
    print(tag, 'idx', md['idx'], 'thresholds', sorted(set(out[f'{tag}_thr'].tolist()))[:8], 'psnr %.3f' % md['metrics']['d1_psnr'])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_block_loops.npz'), **out)
