// Single-output-channel 3x3x3 stride-1 conv / transposed conv on tcgen05 in SCATTER form, fused with bias + ReLU and the
// decoder's clip / threshold / bit-pack.
//
// Replaces the last layer of the V2 synthesis transforms, Conv3DTranspose(1, (3,3,3), 'same') + BiasAdd + Relu
// (reference src/model_transforms.py:107,135), and what the block loops do with its output: clip to [0,1], compare with
// the block's threshold, argwhere (src/model_types.py:201-202,209,233-234).
//
// With one output channel the implicit GEMM of conv3d_umma.cu would spend N = 3*16 padded columns per MMA on 3 useful
// ones.  Here the roles of taps and channels are swapped: ONE un-shifted A tile (all voxels of a halo'd input plane,
// flat index f = y'*10 + x', K = 16 input channels) is multiplied with B = the 27 taps (N = 32), so
//     P[f][tap] = sum_c X[z][f][c] * W[tap][c]
// costs 2 M-halves x {1|3 precision pairs} MMAs per 128 output voxels instead of 27.  The epilogue thread that owns
// halo'd voxel f folds the z taps in registers as the planes stream by
//     S(zo)[f][dy,dx] = P(zo-1)[f][0,dy,dx] + P(zo)[f][1,dy,dx] + P(zo+1)[f][2,dy,dx]
// stores the 9 finished sums of output plane zo to shared memory, and every output voxel gathers its 9 at the tap's
// (y,x) offset:  out[zo][y][x] = b + sum_{dy,dx} S(zo)[(y+dy)*10 + (x+dx)][dy,dx]  -- all in a fixed order
// (deterministic), then ReLU -> fp32 x_hat (optional) and/or min(.,1) > threshold -> one byte of packed occupancy per
// 8-voxel row (optional; integer popcount per block via atomicAdd).
//
// Pipeline per CTA (persistent over (n, y-tile, x-tile) columns, streamed along z), same skeleton as conv3d_umma.cu:
//   warp 0     TMA producer: 4-D box {10 x * 8 ch, 18 y, 1 z, 2 channel groups} per precision term, zero-filled halo
//   warp 1     MMA issuer: per input plane 2 x npairs tcgen05.mma (M=128, N=32, K=16) into a TMEM ring of 4 plane slots
//   warps 2-9  epilogue: tcgen05.ld -> z-tap fold in registers -> 9 sums to shared memory -> bar.sync -> 9-tap gather -> outputs
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {

namespace out1 {
constexpr int TY = 16, TX = 8, PY = TY + 2, PX = TX + 2;
constexpr int PLANE_CG_BYTES = PY * PX * 16;   // 2880
constexpr int NF = PY * PX;                    // 180 halo'd voxels per plane
constexpr int NFP = 192;                       // padded row length of the P ring
constexpr int NTAP = 27, NCOL = 32;            // taps, padded to the MMA N
constexpr int NSLOT = 4;                       // TMEM ring depth (planes)
constexpr int NTHREADS = 64 + 8 * 32;
constexpr int MAX_STAGES = 8;
constexpr int HEADER_BYTES = 1024;
constexpr int A_SLACK = 2048;                  // the second M half reads 1216 B past the last stage (rows f >= 180: unused)
constexpr int P_BYTES = 2 * 9 * NFP * 4;        // two buffers (output-plane parity) of 9 (dy,dx) sums per halo'd voxel

struct Params {
  const float* bias;       // 1 value or null
  const uint8_t* wimg;     // terms x [kcore 2][ngroup 4][8 n][8 k] bf16
  float* xhat;             // (N,1,D,H,W) fp32 or null
  uint32_t* bits;          // (N, D*H*W/32) packed occupancy or null
  const float* thr;        // (N,) thresholds (with bits)
  int32_t* counts;         // (N,) popcounts or null (zeroed by the host wrapper)
  int N, D, H, W, terms, relu;
  int ytiles, xtiles, items, nstage;
};

struct __align__(8) Header {
  uint64_t in_full[MAX_STAGES], in_empty[MAX_STAGES];
  uint64_t acc_full[NSLOT], acc_empty[NSLOT];
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(Header) <= HEADER_BYTES, "header too large");

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
}  // namespace out1

template <int TERMS>
__global__ void __launch_bounds__(out1::NTHREADS, 2)
conv3d_out1_kernel(const __grid_constant__ CUtensorMap tmap_x, const out1::Params p) {
  using namespace out1;
  extern __shared__ __align__(1024) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CGI = 2;
  constexpr int WBYTES_TERM = 2 * (NCOL / 8) * 128;   // 1024
  constexpr int STAGE_BYTES = TERMS * CGI * PLANE_CG_BYTES;
  uint8_t* wsm = smem + HEADER_BYTES;
  uint8_t* stages = wsm + TERMS * WBYTES_TERM;
  float* Ps = reinterpret_cast<float*>(stages + (size_t)p.nstage * STAGE_BYTES + A_SLACK);
  constexpr uint32_t TMEM_COLS = NSLOT * 2 * NCOL;    // 256

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int i = 0; i < p.nstage; ++i) { mbar_init(smem_u32(&hdr->in_full[i]), 1); mbar_init(smem_u32(&hdr->in_empty[i]), 1); }
    for (int i = 0; i < NSLOT; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x * 16; i < TERMS * WBYTES_TERM; i += NTHREADS * 16)
    *reinterpret_cast<int4*>(wsm + i) = __ldg(reinterpret_cast<const int4*>(p.wimg + i));
  // the slack behind the last stage is read (rows never used) by the second M half: keep it finite
  for (int i = threadIdx.x * 16; i < A_SLACK; i += NTHREADS * 16)
    *reinterpret_cast<int4*>(stages + (size_t)p.nstage * STAGE_BYTES + i) = make_int4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, hdr->tmem_base, 0);
  const int D = p.D;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t s = 0, phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int xt = item % p.xtiles, yt = (item / p.xtiles) % p.ytiles, n = item / (p.xtiles * p.ytiles);
        for (int z = 0; z < D; ++z) {
          mbar_wait(smem_u32(&hdr->in_empty[s]), phase ^ 1);
          const uint32_t full = smem_u32(&hdr->in_full[s]);
          mbar_expect_tx(full, (uint32_t)STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < TERMS; ++t)
            tma_load_4d(smem_u32(stages + (size_t)s * STAGE_BYTES + (size_t)t * CGI * PLANE_CG_BYTES), &tmap_x, full,
                        (xt * TX - 1) * 8, yt * TY - 1, z, (t * p.N + n) * CGI);
          if (++s == (uint32_t)p.nstage) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint64_t adesc = make_smem_desc(0, PLANE_CG_BYTES, 128);   // K core matrices one channel group apart; 8-voxel groups contiguous
    const uint64_t bdesc = make_smem_desc(smem_u32(wsm), (NCOL / 8) * 128, 128);
    const uint32_t a_hi = (uint32_t)(adesc >> 32), a_lo_proto = (uint32_t)adesc;
    const uint32_t b_hi = (uint32_t)(bdesc >> 32), b_lo0 = (uint32_t)bdesc;
    constexpr uint32_t idesc = make_idesc(NCOL);
    constexpr int npairs = TERMS == 2 ? 3 : 1;
    const uint32_t stages16 = smem_u32(stages) / 16;
    uint32_t s = 0, in_phase = 0, g = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      for (int z = 0; z < D; ++z, ++g) {
        const uint32_t slot = g & (NSLOT - 1);
        mbar_wait(smem_u32(&hdr->in_full[s]), in_phase);
        mbar_wait(smem_u32(&hdr->acc_empty[slot]), ((g / NSLOT) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a_lo0 = a_lo_proto + stages16 + s * (STAGE_BYTES / 16);
        if (elect_one()) {
#pragma unroll
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int pr = 0; pr < npairs; ++pr) {
              const uint32_t ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
              umma_bf16_lh(tmem_base + slot * (2 * NCOL) + half * NCOL, a_lo0 + ta * (CGI * PLANE_CG_BYTES / 16) + half * (128 * 16 / 16), a_hi,
                           b_lo0 + tb * (WBYTES_TERM / 16), b_hi, idesc, pr == 0 ? 0u : 1u);
            }
          umma_commit(smem_u32(&hdr->in_empty[s]));
          umma_commit(smem_u32(&hdr->acc_full[slot]));
        }
        __syncwarp();
        if (++s == (uint32_t)p.nstage) { s = 0; in_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int grp = ew >> 2;                   // M half this warp drains; parity of the output planes it computes
    const int f = grp * 128 + quad * 32 + lane;  // flat halo'd voxel whose partial sums this thread moves
    const int o = quad * 32 + lane;            // output voxel of the tile this thread computes
    const int yl = o >> 3, xl = o & 7;
    const float bias = p.bias ? __ldg(p.bias) : 0.f;
    const long long HW = (long long)p.H * p.W;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t g = 0;
    float S0[9] = {}, S1[9] = {}, S2[9] = {};   // running (dy,dx) sums of three consecutive output planes, rotated by plane index mod 3
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int xt = item % p.xtiles, yt = (item / p.xtiles) % p.ytiles, n = item / (p.xtiles * p.ytiles);
      const int y = yt * TY + yl, x = xt * TX + xl;
      const float thr = p.bits ? __ldg(p.thr + n) : 0.f;
      int npop = 0;   // occupied voxels this warp produced in this column (one integer atomic per warp and column)
      // one input plane: Sprev / Scur / Snext are the sums of output planes z-1 / z / z+1
      auto plane = [&](int z, float (&Sprev)[9], float (&Scur)[9], float (&Snext)[9]) {
        const uint32_t slot = g & (NSLOT - 1);
        mbar_wait(smem_u32(&hdr->acc_full[slot]), (g / NSLOT) & 1);
        tc_fence_after();
        uint32_t r[NCOL];
        tmem_ld16(lane_base + slot * (2 * NCOL) + grp * NCOL, r);
        tmem_ld16(lane_base + slot * (2 * NCOL) + grp * NCOL + 16, r + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[slot]));
        ++g;
        const bool last = z == D - 1;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          Sprev[k] += __uint_as_float(r[18 + k]);                                      // dz = 2 -> output plane z-1 (complete)
          Scur[k] = (z == 0 ? 0.f : Scur[k]) + __uint_as_float(r[9 + k]);              // dz = 1 -> output plane z
          Snext[k] = __uint_as_float(r[k]);                                            // dz = 0 -> output plane z+1 (first term)
        }
        // The last plane also publishes its own sums, into the buffer that the OTHER group may still be reading for plane z-2 (it was
        // filled one plane ago): wait for those readers first (racecheck: write-after-read hazard without this barrier).
        if (last) named_bar_sync(1, 256);
        if (f < NFP) {
          if (z > 0) {
            float* dst = Ps + (size_t)((z - 1) & 1) * 9 * NFP + f;
#pragma unroll
            for (int k = 0; k < 9; ++k) dst[k * NFP] = Sprev[k];
          }
          if (last) {
            float* dst = Ps + (size_t)(z & 1) * 9 * NFP + f;
#pragma unroll
            for (int k = 0; k < 9; ++k) dst[k * NFP] = Scur[k];
          }
        }
        named_bar_sync(1, 256);
        // output planes that became computable: z-1, and z itself when it is the last plane
        const int zo = (last && (z & 1) == grp) ? z : z - 1;
        if (zo >= 0 && (zo & 1) == grp) {
          const float* src = Ps + (size_t)(zo & 1) * 9 * NFP + yl * PX + xl;
          float acc = bias;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) acc += src[(dy * 3 + dx) * NFP + dy * PX + dx];
          if (p.relu) acc = fmaxf(acc, 0.f);
          const long long vox = ((long long)n * D + zo) * HW + (long long)y * p.W + x;
          if (p.xhat) p.xhat[vox] = acc;
          if (p.bits) {
            const unsigned m = __ballot_sync(0xffffffffu, fminf(acc, 1.0f) > thr);
            if (xl == 0) reinterpret_cast<uint8_t*>(p.bits)[vox >> 3] = (uint8_t)(m >> (lane & 24));
            npop += __popc(m);
          }
        }
      };
      for (int z = 0; z < D; z += 3) {
        plane(z, S2, S0, S1);
        if (z + 1 < D) plane(z + 1, S0, S1, S2);
        if (z + 2 < D) plane(z + 2, S1, S2, S0);
      }
      if (lane == 0 && p.counts && npop) atomicAdd(p.counts + n, npop);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 get_encode_fn2() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
  fn = (EncodeTiledFn2)p;
  return fn;
}

static uint16_t f2bf_o1(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f_o1(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace pccgeo

using namespace pccgeo;

// B image per precision term: [kcore (2)][ngroup (4)][8 n][8 k] bf16; row n = tap (dz*3+dy)*3+dx of the equivalent
// correlation  out[v] = sum X[v + d - 1] * Wc[d]  (Wc = w for a conv, flipped w for a stride-1 transposed conv), k = channel.
extern "C" long long pccgeo_out1_pack_weights_host(const float* w, void* out, int cin, int transposed, int terms) {
  if (cin <= 0 || cin > 16 || (terms != 1 && terms != 2)) { set_error("out1_pack_weights: needs 1..16 input channels, 1 or 2 terms"); return PCCGEO_EINVAL; }
  const long long per_term = 2 * (out1::NCOL / 8) * 64 * 2;
  if (!out) return per_term * terms;
  if (!w) { set_error("out1_pack_weights: null weights"); return PCCGEO_EINVAL; }
  uint16_t* o = (uint16_t*)out;
  memset(o, 0, (size_t)per_term * terms);
  for (int t = 0; t < 27; ++t)
    for (int ci = 0; ci < cin; ++ci) {
      const int src = transposed ? 26 - t : t;   // flipping all three axes = reversing the tap index
      const float val = w[(long long)src * cin + ci];   // tap-major (27, Cin, 1)
      const int kcore = ci >> 3, ki = ci & 7;
      const long long idx = (((long long)kcore * (out1::NCOL / 8) + (t >> 3)) * 8 + (t & 7)) * 8 + ki;
      const uint16_t hi = f2bf_o1(val);
      o[idx] = hi;
      if (terms == 2) o[per_term / 2 + idx] = f2bf_o1(val - bf2f_o1(hi));
    }
  return per_term * terms;
}

extern "C" int pccgeo_conv3d_out1(const void* xb, const void* wpacked, const float* bias, float* x_hat, uint32_t* bits,
                                  const float* thresholds, int32_t* counts, int n, int cin, int d, int h, int wd, int relu,
                                  int terms, void* stream) {
  using namespace out1;
  PCCGEO_REQUIRE(xb && wpacked && (x_hat || bits), "conv3d_out1: null pointer");
  PCCGEO_REQUIRE(!bits || thresholds, "conv3d_out1: packed output needs thresholds");
  PCCGEO_REQUIRE(terms == 1 || terms == 2, "conv3d_out1: terms must be 1 or 2");
  PCCGEO_REQUIRE(cin > 0 && cin <= 16, "conv3d_out1: %d input channels unsupported (1..16)", cin);
  PCCGEO_REQUIRE(n > 0 && d > 0 && h > 0 && wd > 0 && h % TY == 0 && wd % TX == 0, "conv3d_out1: H must be a multiple of 16 and W of 8 (got %dx%dx%d)", d, h, wd);
  EncodeTiledFn2 enc = get_encode_fn2();
  PCCGEO_REQUIRE(enc, "conv3d_out1: cuTensorMapEncodeTiled unavailable");
  Params p{};
  p.bias = bias; p.wimg = (const uint8_t*)wpacked; p.xhat = x_hat; p.bits = bits; p.thr = thresholds; p.counts = bits ? counts : nullptr;
  p.N = n; p.D = d; p.H = h; p.W = wd; p.terms = terms; p.relu = relu;
  p.ytiles = h / TY; p.xtiles = wd / TX; p.items = n * p.ytiles * p.xtiles;
  const int stage_bytes = terms * 2 * PLANE_CG_BYTES;
  const int fixed = HEADER_BYTES + terms * 1024 + A_SLACK + P_BYTES;
  p.nstage = (112 * 1024 - fixed) / stage_bytes;   // two CTAs per SM
  if (p.nstage > MAX_STAGES) p.nstage = MAX_STAGES;
  PCCGEO_REQUIRE(p.nstage >= 2, "conv3d_out1: shared memory budget");
  const size_t smem = fixed + (size_t)p.nstage * stage_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.counts) PCCGEO_CUDA(cudaMemsetAsync(p.counts, 0, sizeof(int32_t) * n, st));

  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)wd * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)terms * n * 2};
  const cuuint64_t gstr[3] = {(cuuint64_t)wd * 16, (cuuint64_t)wd * h * 16, (cuuint64_t)wd * h * d * 16};
  const cuuint32_t box[4] = {PX * 8, PY, 1, 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xb), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCCGEO_REQUIRE(cr == CUDA_SUCCESS, "conv3d_out1: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  const int grid = p.items < 296 ? p.items : 296;
  static bool attr_set = false;
  if (!attr_set) {
    PCCGEO_CUDA(cudaFuncSetAttribute(conv3d_out1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PCCGEO_CUDA(cudaFuncSetAttribute(conv3d_out1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (terms == 2) conv3d_out1_kernel<2><<<grid, NTHREADS, smem, st>>>(tmap, p);
  else conv3d_out1_kernel<1><<<grid, NTHREADS, smem, st>>>(tmap, p);
  return check_launch("conv3d_out1_kernel");
}
