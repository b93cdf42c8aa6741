"""Point-cloud distortion metrics, reference src/utils/pc_metric.py:8-138 (host code outside the hot path: one kd-tree pair
per candidate reconstruction; scipy's cKDTree like the reference): D1 (point-to-point) and, when normals of the original
cloud are given, D2 (point-to-plane) with the reference's normal transfer (`assign_attr`)."""
import numpy as np
from scipy.spatial import cKDTree


def psnr(x, max_energy):  # pc_metric.py:52-53
    with np.errstate(divide='ignore'):
        return 10 * np.log10(np.float64(max_energy) / np.float64(x))


avail_opt_metrics = [y for x in zip(*[(f'd1_{x}', f'd2_{x}') for x in ['sum_AB', 'sum_BA', 'sum_max', 'sum_mean',
                                                                       'mse_AB', 'mse_BA', 'mse']]) for y in x]


def validate_opt_metrics(opt_metrics, with_normals=False):  # pc_metric.py:59-63
    for opt_metric in opt_metrics:
        assert opt_metric in avail_opt_metrics, f'{opt_metric} not found in {avail_opt_metrics}'
        if not with_normals:
            assert not opt_metric.startswith('d2'), f'{opt_metric} not available without normals'


def assign_attr(attr1, idx1, idx2):
    """pc_metric.py:8-27: transfer attributes of x1 to x2.  idx1: (N2,) nearest neighbours of x2 in x1; idx2: (N1,) nearest
    neighbours of x1 in x2.  A point of x2 gets the mean attribute of the x1 points whose nearest neighbour it is, or, if
    there is none, the attribute of its own nearest neighbour in x1."""
    attr1 = np.asarray(attr1, np.float64)
    n2 = idx1.shape[0]
    counts = np.bincount(idx2, minlength=n2).astype(np.float64)
    sums = np.zeros((n2, attr1.shape[1]))
    np.add.at(sums, idx2, attr1)
    empty = counts == 0
    sums[empty] = attr1[idx1[empty]]
    counts[empty] = 1
    return sums / counts[:, None]


def compute_metrics(p1, p2, r, p1_n=None, t1=None):
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
    metrics = {'d1_sum_AB': d1_sum_AB, 'd1_sum_BA': d1_sum_BA, 'd1_sum_max': max(d1_sum_AB, d1_sum_BA),
            'd1_sum_mean': (d1_sum_AB + d1_sum_BA) / 2, 'd1_mse_AB': d1_mse_AB, 'd1_mse_BA': d1_mse_BA,
            'd1_mse': max(d1_mse_AB, d1_mse_BA), 'd1_psnr_AB': psnr(d1_mse_AB, max_energy), 'd1_psnr_BA': psnr(d1_mse_BA, max_energy),
            'd1_psnr': min(psnr(d1_mse_AB, max_energy), psnr(d1_mse_BA, max_energy))}
    if p1_n is not None:   # pc_metric.py:110-137
        p2_n = assign_attr(p1_n, idx1, idx2)
        d2_sum_AB = np.sum(np.sum((p1 - p2[idx2]) * p2_n[idx2], axis=1) ** 2)
        d2_sum_BA = np.sum(np.sum((p2 - p1[idx1]) * np.asarray(p1_n)[idx1], axis=1) ** 2)
        d2_mse_AB, d2_mse_BA = d2_sum_AB / p1.shape[0], d2_sum_BA / p2.shape[0]
        metrics.update({'d2_sum_AB': d2_sum_AB, 'd2_sum_BA': d2_sum_BA, 'd2_sum_max': max(d2_sum_AB, d2_sum_BA),
                        'd2_sum_mean': (d2_sum_AB + d2_sum_BA) / 2, 'd2_mse_AB': d2_mse_AB, 'd2_mse_BA': d2_mse_BA,
                        'd2_mse': max(d2_mse_AB, d2_mse_BA), 'd2_psnr_AB': psnr(d2_mse_AB, max_energy),
                        'd2_psnr_BA': psnr(d2_mse_BA, max_energy),
                        'd2_psnr': min(psnr(d2_mse_AB, max_energy), psnr(d2_mse_BA, max_energy))})
    return metrics
