"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's per-block
threshold optimisation, src/model_opt.py:9-77, and of the D1 part of src/utils/pc_metric.py:76-108 (nearest neighbours with
scipy's cKDTree exactly as the reference; the D2 / normals branch is not restated).  PARITY UNPINNED for the same reason as
the rest of oracle/: the reference module cannot be imported here (pyntcloud is absent) and ships no fixtures; the
restatement follows the source line by line."""
import numpy as np
from scipy.spatial import cKDTree

D1_METRICS = ('d1_sum_AB', 'd1_sum_BA', 'd1_sum_max', 'd1_sum_mean', 'd1_mse_AB', 'd1_mse_BA', 'd1_mse')


def compute_metrics(p1, p2, r, t1=None):  # pc_metric.py:76-108 (D1 only)
    p1, p2 = np.asarray(p1, np.float64), np.asarray(p2, np.float64)
    if t1 is None:
        t1 = cKDTree(p1, balanced_tree=False)
    t2 = cKDTree(p2, balanced_tree=False)
    _, idx2 = t2.query(p1)
    _, idx1 = t1.query(p2)
    max_energy = 3 * r * r
    d1_sum_AB = float(np.sum(np.sum((p1 - p2[idx2]) ** 2, axis=1)))
    d1_sum_BA = float(np.sum(np.sum((p2 - p1[idx1]) ** 2, axis=1)))
    d1_mse_AB, d1_mse_BA = d1_sum_AB / p1.shape[0], d1_sum_BA / p2.shape[0]
    with np.errstate(divide='ignore'):
        psnr = lambda x: 10 * np.log10(np.float64(max_energy) / np.float64(x))
        return {'d1_sum_AB': d1_sum_AB, 'd1_sum_BA': d1_sum_BA, 'd1_sum_max': max(d1_sum_AB, d1_sum_BA),
                'd1_sum_mean': (d1_sum_AB + d1_sum_BA) / 2, 'd1_mse_AB': d1_mse_AB, 'd1_mse_BA': d1_mse_BA,
                'd1_mse': max(d1_mse_AB, d1_mse_BA), 'd1_psnr_AB': psnr(d1_mse_AB), 'd1_psnr_BA': psnr(d1_mse_BA),
                'd1_psnr': min(psnr(d1_mse_AB), psnr(d1_mse_BA))}


def build_points_threshold(x_hat, thresholds, len_block, max_delta=np.inf):  # model_opt.py:9-18
    pa_list = []
    for i, t in enumerate(thresholds):
        pa = np.argwhere(x_hat > t).astype('float32')
        if len(pa) == 0:
            break
        len_ratio = len(pa) / len_block
        if (1 / max_delta) < len_ratio < max_delta:
            pa_list.append((i, pa))
    return pa_list


def compute_optimal_thresholds(block, x_hat, thresholds, resolution, opt_metrics=('d1_mse',), max_deltas=(np.inf,),
                               fixed_threshold=False):  # model_opt.py:21-77 (normals=None)
    for m in opt_metrics:
        assert m in D1_METRICS, m
    assert len(max_deltas) > 0
    best_thresholds = []
    ret_opt_metrics = [f'{opt_metric}_{max_delta}' for max_delta in max_deltas for opt_metric in opt_metrics]
    if fixed_threshold:
        return ret_opt_metrics, [len(thresholds) // 2] * len(max_deltas) * len(opt_metrics)
    block = np.asarray(block)[:, :3]
    pa_list = build_points_threshold(x_hat, thresholds, len(block))
    max_threshold_idx = len(thresholds) - 1
    if len(pa_list) == 0:
        return ret_opt_metrics, [max_threshold_idx] * len(opt_metrics)
    t1 = cKDTree(block, balanced_tree=False)
    pa_metrics = [compute_metrics(block, pa, resolution - 1, t1=t1) for _, pa in pa_list]
    for max_delta in max_deltas:
        cur_pa_list, cur_pa_metrics = pa_list, pa_metrics
        if max_delta is not None:
            cand = build_points_threshold(x_hat, thresholds, len(block), max_delta)
            if len(cand) > 0:
                cur_pa_list = cand
                cur_pa_metrics = [pa_metrics[i] for i in [x[0] for x in cand]]
        for opt_metric in opt_metrics:
            best = int(np.argmin([x[opt_metric] for x in cur_pa_metrics]))
            cur_best_metric = cur_pa_metrics[best][opt_metric]
            mean_point = np.round(np.mean(block, axis=0))[np.newaxis, :]
            mean_point_metric = compute_metrics(block, mean_point, resolution - 1, t1=t1)[opt_metric]
            best_thresholds.append(max_threshold_idx if cur_best_metric > mean_point_metric else cur_pa_list[best][0])
    assert len(ret_opt_metrics) == len(best_thresholds)
    return ret_opt_metrics, best_thresholds
