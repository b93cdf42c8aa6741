// Entropy-model kernels: factorized-prior EntropyBottleneck and scale-hyperprior GaussianConditional.
//
// Replaces the TF-op compositions inside tensorflow-compression 1.3 that the reference calls at
// src/model_types.py:254,258,287,291-292,300,306,333,338-341,377,382-387,397,404-407 and the CPU-pinned
// scale->index foldr of src/utils/patch_gaussian_conditional.py:106-116.
//
// All of this is HBM-bound elementwise work: each kernel reads every input once with coalesced (float4
// where the shape allows) loads, keeps the per-channel parameters in registers, and reduces sum(ln p)
// with warp shuffles into per-block partials that a second one-block kernel adds in a fixed order
// (deterministic -- no fp atomics).
#include "common.cuh"

namespace pccgeo {

struct EbP {
  float m0[3], m1[9], m2[9], m3[3], b0[3], b1[3], b2[3], b3, f0[3], f1[3], f2[3], med;
};

__device__ __forceinline__ EbP load_eb(const float* __restrict__ q) {
  EbP p;
#pragma unroll
  for (int i = 0; i < 3; ++i) p.m0[i] = q[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) p.m1[i] = q[3 + i];
#pragma unroll
  for (int i = 0; i < 9; ++i) p.m2[i] = q[12 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) p.m3[i] = q[21 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { p.b0[i] = q[24 + i]; p.b1[i] = q[27 + i]; p.b2[i] = q[30 + i]; }
  p.b3 = q[33];
#pragma unroll
  for (int i = 0; i < 3; ++i) { p.f0[i] = q[34 + i]; p.f1[i] = q[37 + i]; p.f2[i] = q[40 + i]; }
  p.med = q[43];
  return p;
}

// logits = cumulative-logit MLP 1 -> 3 -> 3 -> 3 -> 1 (tfc EntropyBottleneck._logits_cumulative)
__device__ __forceinline__ float eb_logits(const EbP& p, float v) {
  float h[3], g[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = p.m0[i] * v + p.b0[i];
    h[i] = t + p.f0[i] * tanhf(t);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = p.m1[i * 3] * h[0] + p.m1[i * 3 + 1] * h[1] + p.m1[i * 3 + 2] * h[2] + p.b1[i];
    g[i] = t + p.f1[i] * tanhf(t);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = p.m2[i * 3] * g[0] + p.m2[i * 3 + 1] * g[1] + p.m2[i * 3 + 2] * g[2] + p.b2[i];
    h[i] = t + p.f2[i] * tanhf(t);
  }
  return p.m3[0] * h[0] + p.m3[1] * h[1] + p.m3[2] * h[2] + p.b3;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// grid: (chunks, C, N); each block walks a contiguous slice of one (n, c) plane.
__global__ void eb_quantize_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                   int32_t* __restrict__ sym, float* __restrict__ xhat, int C, int S) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float med = params[c * PCCGEO_EB_PARAM_STRIDE + 43];
  const long long base = ((long long)n * C + c) * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
    const float q = floorf(x[base + i] + (0.5f - med));
    if (sym) sym[base + i] = (int32_t)q;
    if (xhat) xhat[base + i] = q + med;
  }
}

__global__ void eb_dequantize_kernel(const int32_t* __restrict__ sym, const float* __restrict__ params,
                                     float* __restrict__ xhat, int C, int S) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float med = params[c * PCCGEO_EB_PARAM_STRIDE + 43];
  const long long base = ((long long)n * C + c) * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x)
    xhat[base + i] = (float)sym[base + i] + med;
}

__global__ void eb_likelihood_kernel(const float* __restrict__ v, const float* __restrict__ params,
                                     float* __restrict__ lik, double* __restrict__ partials, int N, int C, int S) {
  __shared__ double sm[32];
  const int c = blockIdx.y;
  const EbP p = load_eb(params + c * PCCGEO_EB_PARAM_STRIDE);
  double acc = 0.0;
  // grid.z may be smaller than the batch (the partial-sum workspace is fixed): a block then walks several samples, in order
  for (int n = blockIdx.z; n < N; n += gridDim.z) {
    const long long base = ((long long)n * C + c) * S;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
      const float val = v[base + i];
      const float lo = eb_logits(p, val - 0.5f), up = eb_logits(p, val + 0.5f);
      const float t = lo + up;
      const float s = t > 0.f ? -1.f : (t < 0.f ? 1.f : 0.f);  // -sign(lower + upper)
      float l = fabsf(sigmoidf_(s * up) - sigmoidf_(s * lo));
      l = fmaxf(l, 1e-9f);
      if (lik) lik[base + i] = l;
      acc += (double)logf(l);
    }
  }
  if (partials) {
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
  }
}

__global__ void gc_quantize_kernel(const float* __restrict__ y, const float* __restrict__ sigma,
                                   const float* __restrict__ table, int levels, int32_t* __restrict__ sym,
                                   float* __restrict__ yhat, int32_t* __restrict__ idx, long long count) {
  __shared__ float tab[256];
  for (int i = threadIdx.x; i < levels; i += blockDim.x) tab[i] = table[i];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    if (y) {
      const float q = rintf(y[i]);  // round half to even == tf.math.round
      if (sym) sym[i] = (int32_t)q;
      if (yhat) yhat[i] = q;
    }
    if (idx) {
      const float s = fmaxf(sigma[i], tab[0]);
      // table is increasing: count entries of table[:-1] that are >= s by binary search for the first one
      int lo = 0, hi = levels - 1;  // first j in [0, levels-1) with tab[j] >= s, else levels-1
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] >= s) hi = mid; else lo = mid + 1;
      }
      idx[i] = lo;  // = (levels-1) - #{j < levels-1 : s <= tab[j]}
    }
  }
}

__device__ __forceinline__ float phi_(float x) { return 0.5f * erfcf(-0.70710678118654752440f * x); }

__global__ void gc_likelihood_kernel(const float* __restrict__ v, const float* __restrict__ sigma, float smin,
                                     float* __restrict__ lik, double* __restrict__ partials, long long count) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    // The difference of two Phi values near 0.5 cancels catastrophically in fp32 for large scales; this kernel touches
    // only the latents (32 K values per block), so the difference is evaluated in double and rounded once.
    const double s = (double)fmaxf(sigma[i], smin);
    const double a = (double)fabsf(v[i]);
    const double kInvSqrt2 = 0.70710678118654752440;
    float l = (float)(0.5 * (erfc(-kInvSqrt2 * ((0.5 - a) / s)) - erfc(-kInvSqrt2 * ((-0.5 - a) / s))));
    l = fmaxf(l, 1e-9f);
    if (lik) lik[i] = l;
    acc += (double)logf(l);
  }
  if (partials) {
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
  }
}

__global__ void i32_to_f32_kernel(const int32_t* __restrict__ s, float* __restrict__ o, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    o[i] = (float)s[i];
}

static int grid_for(long long count, int threads, int cap) {
  long long b = (count + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" int pccgeo_eb_quantize(const float* x, const float* eb_params, int32_t* symbols, float* x_hat, int n, int c,
                                  int spatial, void* stream) {
  PCCGEO_REQUIRE(x && eb_params && n > 0 && c > 0 && spatial > 0, "eb_quantize: bad argument");
  PCCGEO_REQUIRE(n <= 65535 && c <= 65535, "eb_quantize: n or c too large");
  dim3 grid(grid_for(spatial, 256, 64), c, n);
  eb_quantize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, eb_params, symbols, x_hat, c, spatial);
  return check_launch("eb_quantize_kernel");
}

extern "C" int pccgeo_eb_dequantize(const int32_t* symbols, const float* eb_params, float* x_hat, int n, int c,
                                    int spatial, void* stream) {
  PCCGEO_REQUIRE(symbols && eb_params && x_hat && n > 0 && c > 0 && spatial > 0, "eb_dequantize: bad argument");
  PCCGEO_REQUIRE(n <= 65535 && c <= 65535, "eb_dequantize: n or c too large");
  dim3 grid(grid_for(spatial, 256, 64), c, n);
  eb_dequantize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(symbols, eb_params, x_hat, c, spatial);
  return check_launch("eb_dequantize_kernel");
}

extern "C" int pccgeo_eb_likelihood(const float* values, const float* eb_params, float* likelihood, double* sum_log,
                                    double* partials, int n, int c, int spatial, void* stream) {
  PCCGEO_REQUIRE(values && eb_params && n > 0 && c > 0 && spatial > 0, "eb_likelihood: bad argument");
  PCCGEO_REQUIRE(n <= 65535 && c <= 65535, "eb_likelihood: n or c too large");
  PCCGEO_REQUIRE(!sum_log || partials, "eb_likelihood: sum_log needs a partials workspace");
  PCCGEO_REQUIRE(c <= kReduceBlocks, "eb_likelihood: %d channels exceed the reduction workspace", c);
  int gx = grid_for(spatial, 128, 16), gz = n;
  while ((long long)gx * c * gz > kReduceBlocks && gx > 1) --gx;
  while ((long long)gx * c * gz > kReduceBlocks && gz > 1) --gz;   // fewer z-blocks than samples: blocks loop over the batch
  dim3 grid(gx, c, gz);
  cudaStream_t st = (cudaStream_t)stream;
  eb_likelihood_kernel<<<grid, 128, 0, st>>>(values, eb_params, likelihood, sum_log ? partials : nullptr, n, c, spatial);
  int rc = check_launch("eb_likelihood_kernel");
  if (rc || !sum_log) return rc;
  finish_sum_kernel<<<1, 256, 0, st>>>(partials, gx * c * gz, sum_log);
  return check_launch("finish_sum_kernel");
}

extern "C" int pccgeo_gc_quantize(const float* y, const float* sigma, const float* scale_table, int levels,
                                  int32_t* symbols, float* y_hat, int32_t* indexes, long long count, void* stream) {
  PCCGEO_REQUIRE(count > 0 && (y || indexes), "gc_quantize: bad argument");
  PCCGEO_REQUIRE(!indexes || (sigma && scale_table && levels >= 2 && levels <= 256), "gc_quantize: indexes need sigma and a table of 2..256 levels");
  gc_quantize_kernel<<<grid_for(count, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(y, sigma, scale_table, levels, symbols,
                                                                                       y_hat, indexes, count);
  return check_launch("gc_quantize_kernel");
}

extern "C" int pccgeo_gc_likelihood(const float* values, const float* sigma, float scale_min, float* likelihood,
                                    double* sum_log, double* partials, long long count, void* stream) {
  PCCGEO_REQUIRE(values && sigma && count > 0, "gc_likelihood: bad argument");
  PCCGEO_REQUIRE(!sum_log || partials, "gc_likelihood: sum_log needs a partials workspace");
  const int g = grid_for(count, 256, kReduceBlocks);
  cudaStream_t st = (cudaStream_t)stream;
  gc_likelihood_kernel<<<g, 256, 0, st>>>(values, sigma, scale_min, likelihood, sum_log ? partials : nullptr, count);
  int rc = check_launch("gc_likelihood_kernel");
  if (rc || !sum_log) return rc;
  finish_sum_kernel<<<1, 256, 0, st>>>(partials, g, sum_log);
  return check_launch("finish_sum_kernel");
}

extern "C" int pccgeo_i32_to_f32(const int32_t* symbols, float* out, long long count, void* stream) {
  PCCGEO_REQUIRE(symbols && out && count > 0, "i32_to_f32: bad argument");
  i32_to_f32_kernel<<<grid_for(count, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(symbols, out, count);
  return check_launch("i32_to_f32_kernel");
}
