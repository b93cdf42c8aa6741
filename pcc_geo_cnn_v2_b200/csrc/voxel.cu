// Voxel-grid helpers around the transforms: densify (sparse_to_dense, reference src/model_types.py:108-114),
// clip + threshold + bit-pack of the decoded occupancy (src/model_types.py:201-202,209,233-234) and the
// focal loss (src/utils/focal_loss.py:5-12).  HBM-bound byte/elementwise work: coalesced loads, one pass.
#include "common.cuh"

namespace pccgeo {

__global__ void densify_kernel(const int16_t* __restrict__ coords, long long npts, float* __restrict__ x, int N, int D,
                               int H, int W, int* __restrict__ err, int block0) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (long long)gridDim.x * blockDim.x) {
    short4 c = reinterpret_cast<const short4*>(coords)[i];  // (block, z, y, x)
    c.x = (short)(c.x - block0);
    if (c.x < 0 || c.x >= N || c.y < 0 || c.y >= D || c.z < 0 || c.z >= H || c.w < 0 || c.w >= W) continue;
    x[(((long long)c.x * D + c.y) * H + c.z) * W + c.w] = 1.0f;
  }
}

// one warp packs 32 consecutive voxels into a word with a ballot; per-block popcounts via atomicAdd (integer:
// order-independent, deterministic).
__global__ void threshold_pack_kernel(const float* __restrict__ xhat, const float* __restrict__ thr,
                                      uint32_t* __restrict__ bits, int32_t* __restrict__ counts, int N,
                                      long long vpb) {
  const long long words_per_block = vpb >> 5;
  const long long total_words = words_per_block * N;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long wi = warp0; wi < total_words; wi += nwarps) {
    const int b = (int)(wi / words_per_block);
    const float t = thr[b];
    const float v = fminf(xhat[wi * 32 + lane], 1.0f);
    const unsigned m = __ballot_sync(0xffffffffu, v > t);
    if (lane == 0) {
      bits[wi] = m;
      if (counts && m) atomicAdd(counts + b, __popc(m));
    }
  }
}

__global__ void focal_loss_kernel(const float* __restrict__ xt, const float* __restrict__ xp, float gamma, float alpha,
                                  double* __restrict__ partials, long long count) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float t = xt[i], p = xp[i];
    float pt1 = t == 1.f ? p : 1.f, pt0 = t == 0.f ? p : 0.f;
    pt1 = fminf(fmaxf(pt1, 1e-3f), .999f);
    pt0 = fminf(fmaxf(pt0, 1e-3f), .999f);
    const float a = alpha * powf(1.f - pt1, gamma) * logf(pt1);
    const float b = (1.f - alpha) * powf(pt0, gamma) * logf(1.f - pt0);
    acc -= (double)a + (double)b;
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" int pccgeo_densify(const int16_t* coords, long long npts, float* x, int n, int d, int h, int wd, void* stream) {
  PCCGEO_REQUIRE(x && n > 0 && d > 0 && h > 0 && wd > 0 && npts >= 0, "densify: bad argument");
  if (npts == 0) return PCCGEO_OK;
  PCCGEO_REQUIRE(coords, "densify: null coords");
  long long b = (npts + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  densify_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(coords, npts, x, n, d, h, wd, nullptr, 0);
  return check_launch("densify_kernel");
}

extern "C" int pccgeo_densify_from(const int16_t* coords, long long npts, int block0, float* x, int n, int d, int h, int wd, void* stream) {
  PCCGEO_REQUIRE(x && n > 0 && d > 0 && h > 0 && wd > 0 && npts >= 0 && block0 >= 0, "densify_from: bad argument");
  if (npts == 0) return PCCGEO_OK;
  PCCGEO_REQUIRE(coords, "densify_from: null coords");
  long long b = (npts + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  densify_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(coords, npts, x, n, d, h, wd, nullptr, block0);
  return check_launch("densify_kernel");
}

extern "C" int pccgeo_threshold_pack(const float* x_hat, const float* thresholds, uint32_t* bits, int32_t* counts, int n,
                                     long long voxels_per_block, void* stream) {
  PCCGEO_REQUIRE(x_hat && thresholds && bits && n > 0, "threshold_pack: bad argument");
  PCCGEO_REQUIRE(voxels_per_block > 0 && voxels_per_block % 32 == 0, "threshold_pack: voxels per block must be a multiple of 32");
  cudaStream_t st = (cudaStream_t)stream;
  if (counts) PCCGEO_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n, st));
  long long words = (voxels_per_block >> 5) * n;
  long long b = (words + 7) / 8;  // 8 warps per block
  if (b > 148 * 16) b = 148 * 16;
  threshold_pack_kernel<<<(int)b, 256, 0, st>>>(x_hat, thresholds, bits, counts, n, voxels_per_block);
  return check_launch("threshold_pack_kernel");
}

extern "C" int pccgeo_focal_loss(const float* x_true, const float* x_pred, float gamma, float alpha, double* out,
                                 double* partials, long long count, void* stream) {
  PCCGEO_REQUIRE(x_true && x_pred && out && partials && count > 0, "focal_loss: bad argument");
  long long b = (count + 255) / 256;
  if (b > kReduceBlocks) b = kReduceBlocks;
  cudaStream_t st = (cudaStream_t)stream;
  focal_loss_kernel<<<(int)b, 256, 0, st>>>(x_true, x_pred, gamma, alpha, partials, count);
  int rc = check_launch("focal_loss_kernel");
  if (rc) return rc;
  finish_sum_kernel<<<1, 256, 0, st>>>(partials, (int)b, out);
  return check_launch("finish_sum_kernel");
}
