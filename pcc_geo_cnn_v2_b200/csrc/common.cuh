// Shared helpers for libpccgeo (sm_100a).  Error text, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/pccgeo.h"

namespace pccgeo {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PCCGEO_ECUDA;
  }
  return PCCGEO_OK;
}

#define PCCGEO_REQUIRE(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      pccgeo::set_error(__VA_ARGS__);        \
      return PCCGEO_EINVAL;                  \
    }                                        \
  } while (0)

#define PCCGEO_CUDA(call)                                                         \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      pccgeo::set_error("%s: %s", #call, cudaGetErrorString(e_));                 \
      return PCCGEO_ECUDA;                                                        \
    }                                                                             \
  } while (0)

constexpr int kReduceBlocks = 1184;  // 8 x 148 SMs: partial-sum slots for the deterministic reductions

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in a fixed order (warp shuffles, then warp 0 over the per-warp partials).
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// Second stage of the two-stage reductions: one block sums `n` partials in a fixed order.
__global__ void finish_sum_kernel(const double* __restrict__ partials, int n, double* __restrict__ out);

}  // namespace pccgeo
