"""Whole-cloud D1 (point-to-point) metrics, reference src/utils/pc_metric.py:76-108 (host code outside the hot path: one
kd-tree pair per candidate reconstruction of the whole cloud; scipy's cKDTree like the reference).  The D2 (point-to-plane)
branch needs normals and is not restated."""
import numpy as np
from scipy.spatial import cKDTree


def psnr(x, max_energy):  # pc_metric.py:52-53
    with np.errstate(divide='ignore'):
        return 10 * np.log10(np.float64(max_energy) / np.float64(x))


def compute_metrics(p1, p2, r, p1_n=None, t1=None):
    if p1_n is not None:
        raise NotImplementedError('D2 metrics (normals) are not implemented; use the reference module for them')
    p1, p2 = np.asarray(p1), np.asarray(p2)
    if t1 is None:
        t1 = cKDTree(p1, balanced_tree=False)
    t2 = cKDTree(p2, balanced_tree=False)
    _, idx2 = t2.query(p1, workers=-1)
    _, idx1 = t1.query(p2, workers=-1)
    max_energy = 3 * r * r
    d1_sum_AB = np.sum(np.sum((p1 - p2[idx2]) ** 2, axis=1))
    d1_sum_BA = np.sum(np.sum((p2 - p1[idx1]) ** 2, axis=1))
    d1_mse_AB, d1_mse_BA = d1_sum_AB / p1.shape[0], d1_sum_BA / p2.shape[0]
    return {'d1_sum_AB': d1_sum_AB, 'd1_sum_BA': d1_sum_BA, 'd1_sum_max': max(d1_sum_AB, d1_sum_BA),
            'd1_sum_mean': (d1_sum_AB + d1_sum_BA) / 2, 'd1_mse_AB': d1_mse_AB, 'd1_mse_BA': d1_mse_BA,
            'd1_mse': max(d1_mse_AB, d1_mse_BA), 'd1_psnr_AB': psnr(d1_mse_AB, max_energy), 'd1_psnr_BA': psnr(d1_mse_BA, max_energy),
            'd1_psnr': min(psnr(d1_mse_AB, max_energy), psnr(d1_mse_BA, max_energy))}
