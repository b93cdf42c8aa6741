// First layer of the V2 / progressive analysis transforms: Conv3D(F, (3,3,3), strides 2, 'same') + BiasAdd + Relu on the
// ONE-channel occupancy volume (reference src/model_transforms.py:67 via AnalysisBlock, first block of :88-92 / :116-120),
// writing the blocked bf16 (hi[/lo]) layout the tensor-core layers consume -- no fp32 intermediate, no layout pass.
//
// HBM-bound by construction (0.45 GMAC per 32 blocks): one thread per output voxel accumulates the <= 27 taps in fp32 in a
// fixed order; the input is an occupancy grid, ~97 % zeros, so a tap costs its 16 FMAs only where a voxel is set.
// TF 'SAME' for k=3, s=2, even sizes pads (0, 1): out[o] = sum_k x[2o + k] w[k].
#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {

template <int TERMS>
__global__ void __launch_bounds__(256, 3) conv3d_first_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ y, int N, int D, int H, int W, int cout, int relu) {
  __shared__ float ws[27 * 16];
  __shared__ float bs[16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) ws[i] = (i % 16) < cout ? w[(i / 16) * cout + (i % 16)] : 0.f;
  if (threadIdx.x < 16) bs[threadIdx.x] = (bias && threadIdx.x < cout) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long HWo = (long long)Ho * Wo, DHWo = HWo * Do, total = DHWo * N;
  const long long term_stride = total * 16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), oz = (int)((i / HWo) % Do);
    const int n = (int)(i / DHWo);
    const float* xn = x + (long long)n * D * H * W;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll 1
    for (int kz = 0; kz < 3; ++kz) {   // one plane of taps at a time: 9 inputs live, not 27
      const int iz = 2 * oz + kz;
      if (iz >= D) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky;
        if (iy >= H) continue;
        const float* row = xn + ((long long)iz * H + iy) * W + 2 * ox;
        const float2 v01 = *reinterpret_cast<const float2*>(row);        // W even, 2*ox even: 8-byte aligned
        const float v2 = 2 * ox + 2 < W ? row[2] : 0.f;
        const float vv[3] = {v01.x, v01.y, v2};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if (vv[kx] == 0.f) continue;
          const float* wt = ws + ((kz * 3 + ky) * 3 + kx) * 16;
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = fmaf(vv[kx], wt[c], acc[c]);
        }
      }
    }
    const long long vox = (long long)oz * HWo + (long long)oy * Wo + ox;
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        v[c] = acc[cg * 8 + c] + bs[cg * 8 + c];
        if (relu) v[c] = fmaxf(v[c], 0.f);
      }
      __nv_bfloat16 hi[8];
      int4 qh, ql;
      uint32_t* qh32 = reinterpret_cast<uint32_t*>(&qh);
      uint32_t* ql32 = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
      for (int c = 0; c < 8; ++c) hi[c] = __float2bfloat16_rn(v[c]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __nv_bfloat162 t2 = __halves2bfloat162(hi[2 * c], hi[2 * c + 1]);
        qh32[c] = *reinterpret_cast<uint32_t*>(&t2);
        ql32[c] = pack_bf16x2(v[2 * c] - __bfloat162float(hi[2 * c]), v[2 * c + 1] - __bfloat162float(hi[2 * c + 1]));
      }
      const long long e = (((long long)n * 2 + cg) * DHWo + vox) * 8;
      *reinterpret_cast<int4*>(y + e) = qh;
      if (TERMS == 2) *reinterpret_cast<int4*>(y + term_stride + e) = ql;
    }
  }
}

}  // namespace pccgeo

using namespace pccgeo;

// x: fp32 (N,1,D,H,W), even D/H/W; w: tap-major fp32 (27, 1, cout), cout <= 16; yb: blocked bf16 (terms, N, 2, D/2, H/2, W/2, 8).
extern "C" int pccgeo_conv3d_first(const float* x, const float* w, const float* bias, void* yb, int n, int d, int h, int wd, int cout,
                                   int relu, int terms, void* stream) {
  PCCGEO_REQUIRE(x && w && yb, "conv3d_first: null pointer");
  PCCGEO_REQUIRE(n > 0 && d > 0 && h > 0 && wd > 0 && d % 2 == 0 && h % 2 == 0 && wd % 2 == 0, "conv3d_first: even dims required (got %dx%dx%d)", d, h, wd);
  PCCGEO_REQUIRE(cout > 0 && cout <= 16 && (terms == 1 || terms == 2), "conv3d_first: cout <= 16, terms 1 or 2");
  const long long total = (long long)n * (d / 2) * (h / 2) * (wd / 2);
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (terms == 2) conv3d_first_kernel<2><<<(int)b, 256, 0, st>>>(x, w, bias, (__nv_bfloat16*)yb, n, d, h, wd, cout, relu);
  else conv3d_first_kernel<1><<<(int)b, 256, 0, st>>>(x, w, bias, (__nv_bfloat16*)yb, n, d, h, wd, cout, relu);
  return check_launch("conv3d_first_kernel");
}
