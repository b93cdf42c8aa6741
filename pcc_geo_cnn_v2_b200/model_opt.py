"""Per-block threshold optimisation (reference src/model_opt.py:9-77, D1 metrics of src/utils/pc_metric.py:76-108) with the
nearest-neighbour sums computed on the GPU for every threshold at once (csrc/threshold_opt.cu) -- SURVEY.md section 8f
"next" #1.  The selection logic (eligibility by max_delta, argmin, the mean-point failure rule, the order of the returned
lists) restates the reference; the D2 metrics need per-point normals and nearest-neighbour indices and are not covered."""
import numpy as np
import torch

from . import _lib as L

D1_METRICS = ('d1_sum_AB', 'd1_sum_BA', 'd1_sum_max', 'd1_sum_mean', 'd1_mse_AB', 'd1_mse_BA', 'd1_mse')


def validate_opt_metrics(opt_metrics, with_normals=False):  # the GPU path: D1 only
    for m in opt_metrics:
        if m.startswith('d2'):
            raise NotImplementedError(f'{m}: the D2 (point-to-plane) metrics run on the host path (compute_optimal_thresholds)')
        assert m in D1_METRICS, f'{m} not found in {D1_METRICS}'


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


def compute_optimal_thresholds(block, x_hat, thresholds, resolution, normals=None, opt_metrics=('d1_mse',), max_deltas=(np.inf,),
                               fixed_threshold=False):
    """model_opt.py:21-77 on the host with kd-trees, any metric incl. D2 (the GPU path, compute_optimal_thresholds_batch,
    covers the D1 metrics and is what compress_blocks uses when no normals are involved)."""
    from scipy.spatial import cKDTree
    from .pc_metric import compute_metrics, validate_opt_metrics as validate
    validate(opt_metrics, with_normals=normals is not None)
    assert len(max_deltas) > 0
    best_thresholds = []
    ret_opt_metrics = [f'{opt_metric}_{max_delta}' for max_delta in max_deltas for opt_metric in opt_metrics]
    if fixed_threshold:
        return ret_opt_metrics, [len(thresholds) // 2] * len(max_deltas) * len(opt_metrics)
    block = np.asarray(block)
    pa_list = build_points_threshold(x_hat, thresholds, len(block))
    max_threshold_idx = len(thresholds) - 1
    if len(pa_list) == 0:
        return ret_opt_metrics, [max_threshold_idx] * len(opt_metrics)
    t1 = cKDTree(block[:, :3], balanced_tree=False)
    pa_metrics = [compute_metrics(block[:, :3], pa, resolution - 1, p1_n=normals, t1=t1) for _, pa in pa_list]
    for max_delta in max_deltas:
        cur_pa_list, cur_pa_metrics = pa_list, pa_metrics
        if max_delta is not None:
            cand = build_points_threshold(x_hat, thresholds, len(block), max_delta)
            if len(cand) > 0:
                cur_pa_list, cur_pa_metrics = cand, [pa_metrics[i] for i in [x[0] for x in cand]]
        for opt_metric in opt_metrics:
            best = int(np.argmin([x[opt_metric] for x in cur_pa_metrics]))
            mean_point = np.round(np.mean(block[:, :3], axis=0))[np.newaxis, :]
            mean_point_metric = compute_metrics(block[:, :3], mean_point, resolution - 1, p1_n=normals, t1=t1)[opt_metric]
            best_thresholds.append(max_threshold_idx if cur_pa_metrics[best][opt_metric] > mean_point_metric else cur_pa_list[best][0])
    assert len(ret_opt_metrics) == len(best_thresholds)
    return ret_opt_metrics, best_thresholds


def threshold_sums(x_hat, thresholds_f32, coords, offsets):
    """x_hat: CUDA fp32 (N,1,D,H,W); thresholds_f32: host float32 (T,) ascending; coords: CUDA int16 (P,4) rows
    (block,z,y,x) sorted by block; offsets: host int64 (N+1,).
    -> (sum_AB, sum_BA, count_B) host int64 (N,T); sum_AB is -1 where B_i is empty."""
    L.require_cuda()
    n, _, d, h, w = x_hat.shape
    t = len(thresholds_f32)
    dev = x_hat.device
    thr = torch.from_numpy(np.ascontiguousarray(thresholds_f32, np.float32)).to(dev)
    offs = torch.from_numpy(np.ascontiguousarray(offsets, np.int64)).to(dev)
    ws = torch.empty(int(L.lib().pccgeo_threshold_opt_ws_bytes(n, d, h, w)), device=dev, dtype=torch.uint8)
    hist = torch.empty((n, t + 1), device=dev, dtype=torch.int64)
    cnt = torch.empty((n, t + 1), device=dev, dtype=torch.int64)
    L.check(L.lib().pccgeo_threshold_hist(L.ptr(x_hat), L.ptr(thr), t, L.ptr(coords), L.ptr(offs), L.ptr(ws), L.ptr(hist), L.ptr(cnt),
                                          n, d, h, w, L.stream_ptr()), 'threshold_hist')
    hist_h, cnt_h = hist.cpu().numpy(), cnt.cpu().numpy()
    # B_i = {rank > i}: exclusive suffix sums over the rank histograms
    sum_ba = np.cumsum(hist_h[:, ::-1], axis=1)[:, ::-1][:, 1:]
    count_b = np.ascontiguousarray(np.cumsum(cnt_h[:, ::-1], axis=1)[:, ::-1][:, 1:])
    cb = torch.from_numpy(count_b).to(dev)
    sab = torch.empty((n, t), device=dev, dtype=torch.int64)
    max_pts = int(np.max(np.diff(offsets))) if n else 0
    L.check(L.lib().pccgeo_threshold_sum_ab(L.ptr(ws), L.ptr(coords), L.ptr(offs), L.ptr(cb), L.ptr(sab), n, t, d, h, w, max_pts,
                                            L.stream_ptr()), 'threshold_sum_ab')
    sum_ab = sab.cpu().numpy()
    for i in range(1, t):  # -2: same point set as the previous threshold
        same = sum_ab[:, i] == -2
        sum_ab[same, i] = sum_ab[same, i - 1]
    return sum_ab, np.ascontiguousarray(sum_ba), count_b


def _metrics(sum_ab, sum_ba, n_a, n_b):
    """the D1 entries of compute_metrics (pc_metric.py:85-97) from the exact integer sums"""
    mse_ab, mse_ba = sum_ab / n_a, sum_ba / n_b
    return {'d1_sum_AB': float(sum_ab), 'd1_sum_BA': float(sum_ba), 'd1_sum_max': float(max(sum_ab, sum_ba)),
            'd1_sum_mean': (sum_ab + sum_ba) / 2, 'd1_mse_AB': mse_ab, 'd1_mse_BA': mse_ba, 'd1_mse': max(mse_ab, mse_ba)}


def select_thresholds(block, sum_ab, sum_ba, count_b, n_thresholds, opt_metrics, max_deltas):
    """model_opt.py:33-77 for one block, given the per-threshold sums."""
    pts = np.asarray(block, np.float64)[:, :3]
    n_a = len(pts)
    ret = []
    max_threshold_idx = n_thresholds - 1
    n_valid = int(np.argmax(count_b == 0)) if (count_b == 0).any() else n_thresholds   # build_points_threshold breaks at the first empty set
    if n_valid == 0:
        return [max_threshold_idx] * len(opt_metrics)
    pa_metrics = [_metrics(int(sum_ab[i]), int(sum_ba[i]), n_a, int(count_b[i])) for i in range(n_valid)]
    mean_point = np.round(np.mean(pts, axis=0))
    d2 = np.sum((pts - mean_point) ** 2, axis=1)
    mean_metrics = _metrics(float(d2.sum()), float(d2.min()), n_a, 1)
    for max_delta in max_deltas:
        cur = list(range(n_valid))
        if max_delta is not None:
            cand = [i for i in range(n_valid) if (1 / max_delta) < count_b[i] / n_a < max_delta]
            if cand:
                cur = cand
        for opt_metric in opt_metrics:
            best = int(np.argmin([pa_metrics[i][opt_metric] for i in cur]))
            if pa_metrics[cur[best]][opt_metric] > mean_metrics[opt_metric]:
                ret.append(max_threshold_idx)   # a single point beats the network output: emit nothing (model_opt.py:66-70)
            else:
                ret.append(cur[best])
    return ret


def compute_optimal_thresholds_batch(blocks, x_hat, thresholds_f32, coords, offsets, opt_metrics=('d1_mse',), max_deltas=(np.inf,)):
    """-> (ret_opt_metrics, (n_blocks, n_metrics) int64 best threshold indexes), as compute_optimal_thresholds per block."""
    validate_opt_metrics(opt_metrics)
    assert len(max_deltas) > 0
    ret_opt_metrics = [f'{m}_{md}' for md in max_deltas for m in opt_metrics]
    sum_ab, sum_ba, count_b = threshold_sums(x_hat, thresholds_f32, coords, offsets)
    best = []
    for j, block in enumerate(blocks):
        r = select_thresholds(block, sum_ab[j], sum_ba[j], count_b[j], len(thresholds_f32), list(opt_metrics), list(max_deltas))
        if len(r) != len(ret_opt_metrics):   # the reference's empty-output early return has one entry per opt_metric only
            r = (r * len(max_deltas))[:len(ret_opt_metrics)]
        best.append(r)
    return ret_opt_metrics, np.asarray(best, np.int64)
