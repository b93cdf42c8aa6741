"""Golden vectors for the D2 (point-to-plane) metrics and the threshold search with normals, produced by the REFERENCE's own
src/utils/pc_metric.py (incl. its numba `assign_attr`) and src/model_opt.py in the build container:

    python tests/golden/make_reference_d2_fixture.py   ->  tests/golden/ref_d2.npz

Shims as in make_reference_model_opt_fixture.py (pyntcloud stub, cKDTree n_jobs -> workers)."""
import os
import sys
import types

import numpy as np
import scipy.spatial
from scipy.ndimage import gaussian_filter

sys.modules['pyntcloud'] = types.SimpleNamespace(PyntCloud=None)
sys.path.insert(0, '/root/reference/src')


class _Tree(scipy.spatial.cKDTree):
    def query(self, x, k=1, eps=0, p=2, distance_upper_bound=np.inf, n_jobs=None, workers=1):
        return super().query(x, k=k, eps=eps, p=p, distance_upper_bound=distance_upper_bound, workers=workers if n_jobs is None else n_jobs)


scipy.spatial.cKDTree = _Tree
import scipy.spatial.ckdtree as _legacy  # noqa: E402
_legacy.cKDTree = _Tree
import model_opt as RMO  # noqa: E402
from utils import pc_metric as RPM  # noqa: E402
RPM.cKDTree = RMO.cKDTree = _Tree

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402

size = 24
rng = np.random.default_rng(5)
thresholds = np.linspace(0, 1.0, 64)
opt_metrics = ['d2_mse', 'd1_mse', 'd2_sum_mean']
max_deltas = [np.inf, 1.4]
out = {'size': np.array(size), 'thresholds': thresholds, 'opt_metrics': np.array(opt_metrics), 'max_deltas': np.array(max_deltas)}
blocks = synthetic.surface_blocks(2, size=size, seed=3)
for j, b in enumerate(blocks):
    normals = rng.normal(size=(len(b), 3))
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    block = np.concatenate([b.astype(np.float64), normals], axis=1)
    occ = np.zeros((size,) * 3, np.float32)
    occ[tuple(b.astype(int).T)] = 1
    x_hat = np.clip(gaussian_filter(occ, 0.9) * 1.5 + rng.normal(size=occ.shape).astype(np.float32) * 0.03, 0, 1).astype(np.float32)
    names, best = RMO.compute_optimal_thresholds(block, x_hat, thresholds, size, normals=block[:, 3:], opt_metrics=opt_metrics,
                                                 max_deltas=max_deltas, fixed_threshold=False)
    p2 = np.argwhere(x_hat > thresholds[20]).astype('float32')   # metrics at a fixed, non-empty threshold
    m = RPM.compute_metrics(block[:, :3], p2, size - 1, p1_n=block[:, 3:])
    keys = sorted(m)
    out[f'block{j}'], out[f'x_hat{j}'], out[f'best{j}'] = block, x_hat, np.array(best, np.int64)
    out[f'metrics{j}'] = np.array([m[k] for k in keys], np.float64)
    print(j, best, {k: round(float(m[k]), 4) for k in ('d1_psnr', 'd2_psnr')})
out['metric_keys'], out['names'], out['n_blocks'] = np.array(keys), np.array(names), np.array(len(blocks))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_d2.npz'), **out)
