// Per-block threshold optimisation on the GPU: the D1 point-to-point sums between a block's points A and the decoded point
// set B_i = { v : x_hat[v] > threshold_i } for EVERY threshold i at once.
//
// Replaces the kd-tree loop of the reference, src/model_opt.py:9-44 (build_points_threshold + one compute_metrics per
// threshold, src/utils/pc_metric.py:76-108: two cKDTree queries per threshold and block) -- SURVEY.md section 8f "next" #1.
// Coordinates are integers on the block's voxel grid, so nearest-neighbour distances are exact squared Euclidean distance
// transforms (EDT) and every sum is an integer:
//   rank[v]      = #{ i : x_hat[v] > t_i }                         (B_i = { rank > i }, nested in i)
//   sum_BA[i]    = sum_{v in B_i} EDT2_A[v]                         -> histogram of EDT2_A over rank, suffix sums (host)
//   count_B[i]   = |B_i|                                            -> histogram of rank, suffix sums (host)
//   sum_AB[i]    = sum_{a in A} EDT2_{B_i}[a]                       -> one CTA per (threshold, block): slice-wise 2-D EDT
//                  (x sweep, y min-plus) of B_i, folded along z only at A's points (never materialised in HBM)
// All arithmetic is int32 / int64: results are exact and independent of the schedule.
#include "common.cuh"

namespace pccgeo {

constexpr int TO_THREADS = 256;
constexpr int TO_INF = 1 << 28;

// rank[v] = number of thresholds strictly below x_hat[v] (thresholds ascending, T <= 65535)
__global__ void rank_kernel(const float* __restrict__ xhat, const float* __restrict__ thr, int T, uint16_t* __restrict__ rank,
                            long long count) {
  extern __shared__ float ts[];
  for (int i = threadIdx.x; i < T; i += blockDim.x) ts[i] = thr[i];
  __syncthreads();
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < count; v += (long long)gridDim.x * blockDim.x) {
    const float x = fminf(fmaxf(xhat[v], 0.f), 1.f);   // np.clip(x_hat, 0, 1), model_types.py:202
    int lo = 0, hi = T;                                 // first i with !(x > t_i)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (x > ts[mid]) lo = mid + 1; else hi = mid;
    }
    rank[v] = (uint16_t)lo;
  }
}

// One z-slice of a binary volume -> its 2-D squared EDT g[y][x] in shared memory; returns false (g untouched) when the
// slice is empty.  occ: H*W bytes; f1, g: H*W int32; rows: H+1 ints (list of the non-empty rows; the y pass only visits
// those -- decoded sets are surfaces, most rows of a slice are empty).  Called by all threads of the CTA.
__device__ __forceinline__ bool slice_edt(const uint8_t* occ, int* f1, int* g, int* rows, int H, int W) {
  if (threadIdx.x == 0) rows[H] = 0;
  __syncthreads();
  // x direction: two running-distance sweeps per row
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    int d = TO_INF;
    bool any = false;
    for (int x = 0; x < W; ++x) {
      const bool o = occ[y * W + x] != 0;
      any |= o;
      d = o ? 0 : (d >= TO_INF ? TO_INF : d + 1);
      f1[y * W + x] = d;
    }
    if (!any) continue;
    rows[atomicAdd(&rows[H], 1)] = y;   // order is irrelevant: the y pass takes a minimum
    d = TO_INF;
    for (int x = W - 1; x >= 0; --x) {
      d = occ[y * W + x] ? 0 : (d >= TO_INF ? TO_INF : d + 1);
      const int m = min(f1[y * W + x], d);
      f1[y * W + x] = m * m;
    }
  }
  __syncthreads();
  const int nr = rows[H];
  if (nr == 0) return false;
  // y direction: min-plus with the parabola (y - y')^2 over the non-empty rows
  for (int e = threadIdx.x; e < H * W; e += blockDim.x) {
    const int y = e / W, x = e - y * W;
    int best = TO_INF;
    for (int r = 0; r < nr; ++r) {
      const int yy = rows[r], dy = y - yy;
      best = min(best, f1[yy * W + x] + dy * dy);
    }
    g[e] = best;
  }
  __syncthreads();
  return true;
}

// sum_AB for one (threshold, block): grid (T, N).  points: int16 (npts_total, 4) rows (block, z, y, x) sorted by block,
// offsets[n] = first row of block n.  counts_b[n*T + i] = |B_i| (0 -> nothing decoded: out = -1); thresholds whose B equals
// the previous one (same count: the sets are nested) are skipped (out = -2, the host copies the previous value).
__global__ void __launch_bounds__(TO_THREADS) sum_ab_kernel(const uint16_t* __restrict__ rank, const int16_t* __restrict__ points,
                                                            const long long* __restrict__ offsets, const long long* __restrict__ counts_b,
                                                            long long* __restrict__ out, int T, int D, int H, int W, int pchunk) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int i = blockIdx.x, n = blockIdx.y;
  const long long cb = counts_b[(long long)n * T + i];
  if (cb == 0 || (i > 0 && counts_b[(long long)n * T + i - 1] == cb)) {
    if (threadIdx.x == 0) out[(long long)n * T + i] = cb == 0 ? -1 : -2;
    return;
  }
  int* f1 = reinterpret_cast<int*>(sm);
  int* g = f1 + H * W;
  int* rows = g + H * W;                     // H + 1 ints
  int* mp = rows + H + 1;                    // per-point running minimum (pchunk entries)
  uint8_t* occ = reinterpret_cast<uint8_t*>(mp + pchunk);
  __shared__ long long red[TO_THREADS / 32];
  const long long p0 = offsets[n], p1 = offsets[n + 1];
  const uint16_t* rk = rank + (long long)n * D * H * W;
  long long total = 0;
  for (long long c0 = p0; c0 < p1; c0 += pchunk) {
    const int np = (int)min((long long)pchunk, p1 - c0);
    for (int q = threadIdx.x; q < np; q += blockDim.x) mp[q] = TO_INF;
    for (int z = 0; z < D; ++z) {
      __syncthreads();
      for (int e = threadIdx.x; e < H * W; e += blockDim.x) occ[e] = rk[(long long)z * H * W + e] > (uint16_t)i;
      __syncthreads();
      if (!slice_edt(occ, f1, g, rows, H, W)) continue;   // empty slice: no candidate for any point
      for (int q = threadIdx.x; q < np; q += blockDim.x) {
        const short4 pt = reinterpret_cast<const short4*>(points)[c0 + q];   // (block, z, y, x)
        const int dz = pt.y - z;
        const int cand = g[pt.z * W + pt.w] + dz * dz;
        mp[q] = min(mp[q], cand);
      }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < np; q += blockDim.x) total += mp[q];
    __syncthreads();
  }
  // block sum (integers: any order gives the same result)
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int w = 0; w < TO_THREADS / 32; ++w) s += red[w];
    out[(long long)n * T + i] = s;
  }
}

// 2-D EDT of every z-slice of the block's own occupancy (points A): grid (D, N) -> gA int32 (N, D, H, W)
__global__ void __launch_bounds__(TO_THREADS) slice_edt_points_kernel(const int16_t* __restrict__ points,
                                                                      const long long* __restrict__ offsets, int* __restrict__ gA,
                                                                      int D, int H, int W) {
  extern __shared__ __align__(16) uint8_t sm[];
  int* f1 = reinterpret_cast<int*>(sm);
  int* g = f1 + H * W;
  int* rows = g + H * W;
  uint8_t* occ = reinterpret_cast<uint8_t*>(rows + H + 1);
  const int z = blockIdx.x, n = blockIdx.y;
  for (int e = threadIdx.x; e < H * W; e += blockDim.x) occ[e] = 0;
  __syncthreads();
  for (long long q = offsets[n] + threadIdx.x; q < offsets[n + 1]; q += blockDim.x) {
    const short4 pt = reinterpret_cast<const short4*>(points)[q];
    if (pt.y == z) occ[pt.z * W + pt.w] = 1;
  }
  __syncthreads();
  const bool any = slice_edt(occ, f1, g, rows, H, W);
  int* dst = gA + (((long long)n * D + z) * H) * W;
  for (int e = threadIdx.x; e < H * W; e += blockDim.x) dst[e] = any ? g[e] : TO_INF;
}

// z fold of gA + histograms over rank: hist[n][k] += EDT2_A[v], cnt[n][k] += 1 (k = rank[v], 0..T); grid (chunks, N)
__global__ void __launch_bounds__(TO_THREADS) fold_hist_kernel(const int* __restrict__ gA, const uint16_t* __restrict__ rank,
                                                               unsigned long long* __restrict__ hist, unsigned long long* __restrict__ cnt,
                                                               int T, int D, int H, int W) {
  extern __shared__ unsigned long long hs[];   // [T+1] sums, [T+1] counts
  unsigned long long* cs = hs + (T + 1);
  const int n = blockIdx.y;
  for (int k = threadIdx.x; k < 2 * (T + 1); k += blockDim.x) hs[k] = 0;
  __syncthreads();
  const long long HW = (long long)H * W, V = HW * D;
  const int* ga = gA + (long long)n * V;
  const uint16_t* rk = rank + (long long)n * V;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const int k = rk[v];
    if (k == 0) continue;                      // in no B_i: contributes to no threshold
    const int z = (int)(v / HW);
    const long long yx = v - (long long)z * HW;
    int best = TO_INF;
    for (int zz = 0; zz < D; ++zz) {
      const int dz = z - zz;
      best = min(best, ga[(long long)zz * HW + yx] + dz * dz);
    }
    atomicAdd(&hs[k], (unsigned long long)best);
    atomicAdd(&cs[k], 1ull);
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= T; k += blockDim.x) {
    if (hs[k]) atomicAdd(&hist[(long long)n * (T + 1) + k], hs[k]);
    if (cs[k]) atomicAdd(&cnt[(long long)n * (T + 1) + k], cs[k]);
  }
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" size_t pccgeo_threshold_opt_ws_bytes(int n, int d, int h, int wd) {
  return (size_t)n * d * h * wd * (sizeof(uint16_t) + sizeof(int));   // rank + per-slice EDT of A
}

// Stage 1: rank volume, histograms of EDT2_A and of the voxel counts over rank.  hist / cnt: (N, T+1) uint64, zeroed here.
extern "C" int pccgeo_threshold_hist(const float* x_hat, const float* thresholds, int t, const int16_t* points,
                                     const long long* offsets, void* ws, unsigned long long* hist, unsigned long long* cnt, int n,
                                     int d, int h, int wd, void* stream) {
  PCCGEO_REQUIRE(x_hat && thresholds && points && offsets && ws && hist && cnt, "threshold_hist: null pointer");
  PCCGEO_REQUIRE(n > 0 && t > 0 && t <= 4096 && d > 0 && h > 0 && wd > 0 && h <= 128 && wd <= 128, "threshold_hist: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)d * h * wd;
  uint16_t* rank = (uint16_t*)ws;
  int* gA = (int*)((uint8_t*)ws + (size_t)n * V * sizeof(uint16_t));
  long long b = (n * V + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  rank_kernel<<<(int)b, 256, t * sizeof(float), st>>>(x_hat, thresholds, t, rank, n * V);
  int rc = check_launch("rank_kernel");
  if (rc) return rc;
  const size_t sm2 = (size_t)h * wd * (2 * sizeof(int) + 1) + (h + 1) * sizeof(int);
  static bool attr = false;
  if (!attr) {
    PCCGEO_CUDA(cudaFuncSetAttribute(slice_edt_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PCCGEO_CUDA(cudaFuncSetAttribute(sum_ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  slice_edt_points_kernel<<<dim3(d, n), TO_THREADS, sm2, st>>>(points, offsets, gA, d, h, wd);
  rc = check_launch("slice_edt_points_kernel");
  if (rc) return rc;
  PCCGEO_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * n * (t + 1), st));
  PCCGEO_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * n * (t + 1), st));
  fold_hist_kernel<<<dim3(32, n), TO_THREADS, 2 * (t + 1) * sizeof(unsigned long long), st>>>(gA, rank, hist, cnt, t, d, h, wd);
  return check_launch("fold_hist_kernel");
}

// Stage 2: sum_AB[n][i] for every threshold (needs counts_b = |B_i| from stage 1's suffix sums; -1 where B_i is empty, -2
// where B_i equals B_{i-1}).  ws is the workspace stage 1 filled.
extern "C" int pccgeo_threshold_sum_ab(const void* ws, const int16_t* points, const long long* offsets, const long long* counts_b,
                                       long long* sum_ab, int n, int t, int d, int h, int wd, int max_points, void* stream) {
  PCCGEO_REQUIRE(ws && points && offsets && counts_b && sum_ab, "threshold_sum_ab: null pointer");
  PCCGEO_REQUIRE(n > 0 && t > 0 && d > 0 && h > 0 && wd > 0 && h <= 128 && wd <= 128, "threshold_sum_ab: bad shape");
  const size_t fixed = (size_t)h * wd * (2 * sizeof(int) + 1) + (h + 1) * sizeof(int) + 16;
  long long pchunk = max_points > 0 ? max_points : 1;
  const long long room = (long long)(200 * 1024 - fixed) / (long long)sizeof(int);
  if (pchunk > room) pchunk = room;
  pchunk = (pchunk + 3) / 4 * 4;
  const size_t smem = fixed + (size_t)pchunk * sizeof(int);
  sum_ab_kernel<<<dim3(t, n), TO_THREADS, smem, (cudaStream_t)stream>>>((const uint16_t*)ws, points, offsets, counts_b, sum_ab, t, d, h,
                                                                         wd, (int)pchunk);
  return check_launch("sum_ab_kernel");
}
