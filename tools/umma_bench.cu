// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode, no-swizzle K-major) as a function of N and
// of the A-operand address pattern (alignment of the 128-byte core matrices, row pitch).  Results feed DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pcc_geo_cnn_v2_b200/csrc tools/umma_bench.cu -o gpurun_out/umma_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "umma_ptx.cuh"

using namespace pccgeo;

struct Cfg {
  int n;        // MMA N
  int pitch;    // A row pitch (SBO) in bytes
  int lbo;      // A LBO in bytes
  int pattern;  // 0: fixed base; 1: cycle the 9 (ky,kx) offsets ky*pitch + kx*16; 2: cycle 3 ky offsets only; 3: fixed base + 16
  int iters;    // groups of 9 MMAs
  int ts;       // 1: A operand from TMEM
  int rot;      // number of distinct accumulators cycled over (1, 2, 4)
  int m64;      // 1: M = 64
  int smem_kb;
  int dual;     // 1: two warps of the same CTA issue concurrently (different accumulators)
  int commit_every;  // > 0: tcgen05.commit (to a scratch mbarrier, nobody waits) after this many groups of 9 MMAs
};

__global__ void __launch_bounds__(128, 1) bench_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2, scratch[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x * 16; i < c.smem_kb * 1024; i += blockDim.x * 16) *reinterpret_cast<int4*>(smem + i) = make_int4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_init(smem_u32(&scratch[0]), 1); mbar_init(smem_u32(&scratch[1]), 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(c.smem_kb >= 200 ? 512 : 256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0 || (c.dual && warp == 1)) {
    const uint32_t mybar = warp == 0 ? smem_u32(&bar) : smem_u32(&bar2);
    const uint32_t dwarp = warp == 0 ? 0u : 256u;
    const uint32_t a_base = smem_u32(smem) + 1024;            // A region: 64 KB
    const uint32_t b_base = smem_u32(smem) + (c.smem_kb >= 200 ? 96 : 40) * 1024;       // B region
    const uint64_t ad = make_smem_desc(a_base, c.lbo, c.pitch);
    const uint64_t bd = make_smem_desc(b_base, (c.n / 8) * 128, 128);
    const uint32_t a_lo0 = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32), b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
    const uint32_t idesc = c.m64 ? ((make_idesc(c.n) & ~(0x1fu << 24)) | ((64u >> 4) << 24)) : make_idesc(c.n);
    uint32_t doff[9];
    for (int k = 0; k < 9; ++k) doff[k] = (uint32_t)((k % c.rot) * (c.rot == 1 ? 0 : (c.n <= 128 ? 128 : 256)));
    uint32_t offs[9];
    for (int k = 0; k < 9; ++k) {
      int o = 0;
      if (c.pattern == 1) o = (k / 3) * c.pitch + (k % 3) * 16;
      if (c.pattern == 2) o = (k / 3) * c.pitch;
      if (c.pattern == 3) o = 16;
      offs[k] = o / 16;
    }
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      __syncwarp();
      t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            if (c.ts) {
              asm volatile(
                  "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %3};\n"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %5, p;\n}\n" ::"r"(tmem_base + doff[k]),
                  "r"(tmem_base + (c.smem_kb >= 200 ? 448 : 192) + (k % 8) * 8), "r"(b_lo), "r"(b_hi), "r"(1u), "r"(idesc)
                  : "memory");
            } else {
              umma_bf16_lh(tmem_base + dwarp + doff[k], a_lo0 + offs[k], a_hi, b_lo, b_hi, idesc, 1u);
            }
          }
          if (c.commit_every > 0 && (it % c.commit_every) == c.commit_every - 1) umma_commit(smem_u32(&scratch[warp & 1]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(mybar);
      __syncwarp();
      mbar_wait(mybar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(c.smem_kb >= 200 ? 512 : 256) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int ns[] = {16, 32, 48, 64, 96, 128, 144, 192, 256};
  struct P { const char* name; int ts, rot, m64, ctas_per_sm, dual, ce; } pats[] = {
      {"SS 2 warps of one CTA (per warp)", 0, 1, 0, 1, 1, 0},
      {"SS 2 warps, commit every 27 MMAs", 0, 1, 0, 1, 1, 3},
      {"SS 2 warps, commit every 9 MMAs", 0, 1, 0, 1, 1, 1},
      {"SS 1 warp, commit every 9 MMAs", 0, 1, 0, 1, 0, 1},
      {"SS 2 CTAs/SM, commit every 9 MMAs", 0, 1, 0, 2, 0, 1},
      {"SS 1 acc", 0, 1, 0, 1, 0, 0},
      {"SS 2 acc rotate", 0, 2, 0, 1, 0, 0},
      {"SS 3 acc rotate", 0, 3, 0, 1, 0, 0},
      {"SS M=64", 0, 1, 1, 1, 0, 0},
      {"SS 2 CTAs/SM (per CTA)", 0, 1, 0, 2, 0, 0},
      {"SS 2 CTAs/SM, 2 acc", 0, 2, 0, 2, 0, 0},
      {"TS 1 acc", 1, 1, 0, 1, 0, 0},
      {"TS 2 acc rotate", 1, 2, 0, 1, 0, 0},
      {"TS 2 CTAs/SM (per CTA)", 1, 1, 0, 2, 0, 0},
  };
  printf("%-32s", "cycles/MMA  N=");
  for (int n : ns) printf("%7d", n);
  printf("\n");
  for (auto& pt : pats) {
    printf("%-32s", pt.name);
    for (int n : ns) {
      if (pt.rot == 3 && n > 128) { printf("      -"); continue; }
      if (pt.ctas_per_sm == 2 && pt.rot * (n <= 128 ? 128 : 256) > 256 && pt.rot > 1) { printf("      -"); continue; }
      const int smem_kb = pt.ctas_per_sm == 2 ? 100 : 200;
      Cfg c{n, 160, 2880, 1, 400, pt.ts, pt.rot, pt.m64, smem_kb, pt.dual, pt.ce};
      bench_kernel<<<148 * pt.ctas_per_sm, 128, smem_kb * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  ERR %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("%7.1f", (double)cyc / (400 * 9));
    }
    printf("\n");
  }
  printf("floor (N/2)                     ");
  for (int n : ns) printf("%7.1f", n / 2.0);
  printf("\n");
  return 0;
}
