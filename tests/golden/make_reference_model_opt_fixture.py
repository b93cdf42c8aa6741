"""Golden vectors for the per-block threshold optimisation, produced by the REFERENCE's own src/model_opt.py and
src/utils/pc_metric.py in the build container:

    python tests/golden/make_reference_model_opt_fixture.py   ->  tests/golden/ref_model_opt.npz

Two environment shims, neither touching the algorithm: `pyntcloud` (imported at the top of pc_metric.py, used only under
__main__) is stubbed, and cKDTree.query's old `n_jobs=` keyword is mapped to today's `workers=`."""
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
import scipy.spatial.ckdtree as _legacy  # noqa: E402  (model_opt.py imports the class from the legacy module path)
_legacy.cKDTree = _Tree
import model_opt as RMO  # noqa: E402
from utils import pc_metric as RPM  # noqa: E402
RPM.cKDTree = _Tree
RMO.cKDTree = _Tree

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402

size = 32
rng = np.random.default_rng(77)
thresholds = np.linspace(0, 1.0, 256)     # CompressionModel.thresholds, model_types.py:181
opt_metrics = ['d1_mse', 'd1_sum_mean', 'd1_mse_AB', 'd1_sum_BA']
max_deltas = [np.inf, 1.5, 1.1]
blocks = synthetic.surface_blocks(4, size=size, seed=12)
out = {'size': np.array(size), 'thresholds': thresholds, 'opt_metrics': np.array(opt_metrics), 'max_deltas': np.array(max_deltas),
       'n_blocks': np.array(len(blocks))}
for j, b in enumerate(blocks):
    occ = np.zeros((size,) * 3, np.float32)
    occ[tuple(b.astype(int).T)] = 1
    # decoder-like field: blurred occupancy + noise; block 3 is a failure case (field unrelated to the block)
    field = gaussian_filter(occ, 0.8 + 0.3 * j) * (2.5 - 0.4 * j) + rng.normal(size=occ.shape).astype(np.float32) * 0.03
    if j == 3:
        field = rng.random(occ.shape).astype(np.float32) ** 6
    x_hat = np.clip(field, 0.0, 1.0).astype(np.float32)                      # model_types.py:202
    names, best = RMO.compute_optimal_thresholds(b, x_hat, thresholds, size, normals=None, opt_metrics=opt_metrics,
                                                 max_deltas=max_deltas, fixed_threshold=False)
    m = RPM.compute_metrics(b[:, :3], np.argwhere(x_hat > thresholds[best[0]]).astype('float32'), size - 1)
    out[f'block{j}'] = b.astype(np.int16)
    out[f'x_hat{j}'] = x_hat.astype(np.float16).astype(np.float32) if False else x_hat
    out[f'best{j}'] = np.array(best, np.int64)
    out[f'metrics{j}'] = np.array([m['d1_sum_AB'], m['d1_sum_BA'], m['d1_mse'], m['d1_psnr']], np.float64)
    print(j, names[:2], best)
out['names'] = np.array(names)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_model_opt.npz'), **out)
