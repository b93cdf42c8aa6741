// 3x3x3 stride-1 conv / transposed conv for 16 -> 16 channel layers on tcgen05 with BOTH the z taps and the y taps stacked
// in the MMA N dimension ("y-stacked" variant of conv3d_umma.cu; same reference call sites: the second and third layer
// of AnalysisBlock / SynthesisBlock at 16 filters, src/model_transforms.py:62-81).
//
// Why: tools/umma_bench.cu measures 46 + N/2 cycles per tcgen05.mma (M=128, K=16, operands in shared memory), so the
// small-N MMAs of a 16-channel layer are bound by the fixed per-instruction cost, not by math.  conv3d_umma.cu issues 9
// (ky,kx) MMAs of N = 3*16 per input plane and precision pair; here the three ky taps are stacked as well:
//     D[(yi,x), (j, ky, co)] += sum_{kx,ci} X[z][yi][x+kx-1][ci] * W[kz(j)][ky][kx][ci][co]        (3 MMAs of N = 144)
// i.e. input row yi produces, per ky, its contribution to OUTPUT row yi-ky+1.  The epilogue adds the three partials of an
// output row, which live 8 TMEM lanes apart (lane = 8*row + x): warp shuffles inside a lane quadrant, a 1 KB shared-
// memory exchange across quadrant boundaries.  The M tile is 16 INPUT rows (one halo row each side), so 14 of its rows
// are valid outputs: 9 MMAs of N=144 for 112 output voxels instead of 27 of N=48 for 128.
//
// Pipeline: as conv3d_umma.cu (warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue in two groups alternating over
// output planes); two CTAs per SM (interleaved MMA streams hide part of the fixed cost), TMEM ring of 5 planes x 48 columns.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {
namespace ys {
constexpr int COUT = 16;
constexpr int TYI = 16, TYO = 14, TX = 8, PX = TX + 2;
constexpr int ROW_PITCH = PX * 16;               // 160 B between input rows (SBO of A)
constexpr int PLANE_CG_BYTES = TYI * ROW_PITCH;  // 2560 B: one channel group of one halo'd input plane
constexpr int SC = 3 * COUT;                     // TMEM columns of one output-plane slot: [ky][co]
constexpr int NB = 3 * SC;                       // rows of one B tile: [j (z plane)][ky][co]
constexpr int NUM_THREADS = 64 + 8 * 32;
constexpr int MAX_STAGES = 8, MAX_SLOTS = 10;
constexpr int HEADER_BYTES = 1024;
constexpr int XCH_FLOATS = 2 * 2 * 4 * 2 * COUT * 8;  // [group][buffer][quadrant][direction][c][x]
constexpr int XCH_BYTES = XCH_FLOATS * 4;

struct Params {
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  const uint8_t* wimg;
  int N, D, H, W, relu, cout_real;
  int ytiles, xtiles, items, nstage;
  int wbytes_term;
  long long term_stride_out;
};

struct __align__(8) Header {
  uint64_t in_full[MAX_STAGES], in_empty[MAX_STAGES];
  uint64_t acc_full[MAX_SLOTS], acc_empty[MAX_SLOTS];
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(Header) <= HEADER_BYTES - 64, "header too large");

__device__ __forceinline__ void group_bar_sync(int grp) {   // named barrier of one epilogue group (4 warps)
  if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
}  // namespace ys

template <int TERMS, int NSLOTS>   // precision terms (1,2); TMEM ring depth (5: two CTAs per SM, 10: one)
__global__ void __launch_bounds__(ys::NUM_THREADS, NSLOTS == 5 ? 2 : 1)
conv3d_umma_ys_kernel(const __grid_constant__ CUtensorMap tmap_x, const ys::Params p) {
  using namespace ys;
  extern __shared__ __align__(1024) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CGI = 2;
  constexpr int STAGE_BYTES = TERMS * CGI * PLANE_CG_BYTES;
  const int wbytes_all = p.wbytes_term * TERMS;
  uint8_t* wsm = smem + HEADER_BYTES;
  float* xch = reinterpret_cast<float*>(wsm + ((wbytes_all + 127) & ~127));
  uint8_t* stages = reinterpret_cast<uint8_t*>(xch) + XCH_BYTES;
  constexpr uint32_t TMEM_COLS = NSLOTS == 5 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int i = 0; i < p.nstage; ++i) { mbar_init(smem_u32(&hdr->in_full[i]), 1); mbar_init(smem_u32(&hdr->in_empty[i]), 1); }
    for (int i = 0; i < NSLOTS; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x * 16; i < wbytes_all; i += NUM_THREADS * 16)
    *reinterpret_cast<int4*>(wsm + i) = __ldg(reinterpret_cast<const int4*>(p.wimg + i));
  if (threadIdx.x < COUT)   // bias lives in the last 64 bytes of the header block (read as a broadcast by the epilogue)
    reinterpret_cast<float*>(smem + HEADER_BYTES - 64)[threadIdx.x] = (p.bias && (int)threadIdx.x < p.cout_real) ? __ldg(p.bias + threadIdx.x) : 0.f;
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
                        (xt * TX - 1) * 8, yt * TYO - 1, z, (t * p.N + n) * CGI);
          if (++s == (uint32_t)p.nstage) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t b_kcore = (NB / 8) * 128;           // bytes between the two K core matrices of a B tile
    constexpr uint32_t b_tile16 = 2 * b_kcore / 16;        // one kx tile, in 16-byte units
    constexpr uint32_t b_plane16 = (SC / 8) * 128 / 16;    // one stacked output plane (SC rows)
    constexpr int npairs = TERMS == 2 ? 3 : 1;
    const uint64_t bdesc0 = make_smem_desc(smem_u32(wsm), b_kcore, 128);
    const uint32_t b_lo0 = (uint32_t)bdesc0, b_hi = (uint32_t)(bdesc0 >> 32);
    const uint64_t adesc_proto = make_smem_desc(0, PLANE_CG_BYTES, ROW_PITCH);
    const uint32_t a_hi = (uint32_t)(adesc_proto >> 32), a_lo_proto = (uint32_t)adesc_proto;
    constexpr uint32_t a_term16 = CGI * PLANE_CG_BYTES / 16;
    const uint32_t w_term16 = p.wbytes_term / 16;
    const uint32_t stages16 = smem_u32(stages) / 16;
    uint32_t s = 0, in_phase = 0;
    uint32_t g0 = 0;  // running output-plane counter at the start of this item
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, g0 += D) {
      for (int z = 0; z < D; ++z) {
        mbar_wait(smem_u32(&hdr->in_full[s]), in_phase);
        const int pa = z > 0 ? z - 1 : 0, pb = z + 1 < D ? z + 1 : D - 1;
        auto wait_empty = [&](uint32_t g) { mbar_wait(smem_u32(&hdr->acc_empty[g % NSLOTS]), (g / NSLOTS) & 1); };
        if (z == 0) wait_empty(g0);
        if (z + 1 < D) wait_empty(g0 + z + 1);
        tc_fence_after();
        const uint32_t a_lo0 = a_lo_proto + stages16 + s * (STAGE_BYTES / 16);
        const uint32_t slot_a = (g0 + pa) % NSLOTS;
        const int nplanes = pb - pa + 1;
        const int n0 = min(nplanes, (int)(NSLOTS - slot_a));
        const uint32_t j0 = pa - (z - 1);
        const uint32_t seg_d0 = tmem_base + slot_a * SC, seg_b0 = b_lo0 + j0 * b_plane16, seg_i0 = make_idesc(n0 * SC);
        const uint32_t seg_d1 = tmem_base, seg_b1 = b_lo0 + (j0 + n0) * b_plane16, seg_i1 = make_idesc((nplanes - n0) * SC);
        const bool two = n0 < nplanes;
        if (elect_one()) {
          if (!two) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int pr = 0; pr < npairs; ++pr) {
                const uint32_t ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
                umma_bf16_lh(seg_d0, a_lo0 + kx + ta * a_term16, a_hi, seg_b0 + tb * w_term16 + kx * b_tile16, b_hi, seg_i0, 1u);
              }
          } else {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int pr = 0; pr < npairs; ++pr) {
                const uint32_t ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
                const uint32_t a_lo = a_lo0 + kx + ta * a_term16, b_off = tb * w_term16 + kx * b_tile16;
                umma_bf16_lh(seg_d0, a_lo, a_hi, seg_b0 + b_off, b_hi, seg_i0, 1u);
                umma_bf16_lh(seg_d1, a_lo, a_hi, seg_b1 + b_off, b_hi, seg_i1, 1u);
              }
          }
          umma_commit(smem_u32(&hdr->in_empty[s]));
          if (z >= 1) umma_commit(smem_u32(&hdr->acc_full[(g0 + z - 1) % NSLOTS]));
          if (z == D - 1) umma_commit(smem_u32(&hdr->acc_full[(g0 + z) % NSLOTS]));
        }
        __syncwarp();
        if (++s == (uint32_t)p.nstage) { s = 0; in_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int ew = warp - 2;
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int grp = ew >> 2;          // output planes with (g & 1) == grp
    const int row = quad * 32 + lane; // M row = TMEM lane = 8 * input row + x
    const int il = row >> 3, xl = row & 7;
    const float* bias_s = reinterpret_cast<const float*>(smem + HEADER_BYTES - 64);   // staged before the role split
    const long long HW = (long long)p.H * p.W, DHW = HW * D;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    // zero the accumulator slots this group drains first (slot parity == group for the first lap) ... every slot is
    // drained by whichever group owns the plane that lands in it, so zero them all once, split between the groups
    for (int slot = grp; slot < NSLOTS; slot += 2) {
#pragma unroll
      for (int c = 0; c < SC; c += 16) tmem_st16_zero(lane_base + (uint32_t)slot * SC + c);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[slot]));
    }
    uint32_t g0 = 0, nproc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, g0 += D) {
      const int xt = item % p.xtiles, yt = (item / p.xtiles) % p.ytiles, n = item / (p.xtiles * p.ytiles);
      const int yo = yt * TYO + il - 1;                         // output row of this thread (valid for 1 <= il <= 14)
      const bool valid = il >= 1 && il <= TYO && yo < p.H;
      const long long vox0 = (long long)yo * p.W + (xt * TX + xl);
      for (int pl = 0; pl < D; ++pl) {
        const uint32_t g = g0 + pl;
        if ((int)(g & 1) != grp) continue;
        const uint32_t slot = g % NSLOTS;
        mbar_wait(smem_u32(&hdr->acc_full[slot]), (g / NSLOTS) & 1);
        tc_fence_after();
        uint32_t r[SC];
#pragma unroll
        for (int c = 0; c < SC; c += 16) tmem_ld16(lane_base + slot * SC + c, r + c);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < SC; c += 16) tmem_st16_zero(lane_base + slot * SC + c);
        tmem_st_wait();  // the slot is handed back zeroed: its next tenant only ever accumulates
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[slot]));
        // cross-quadrant exchange: my first row's ky=2 partial belongs to the last row of the quadrant below, my last row's
        // ky=0 partial to the first row of the quadrant above
        float* xb = xch + ((grp * 2 + (nproc & 1)) * 4) * (2 * COUT * 8);
        ++nproc;
        if (lane < 8) {
          float* d = xb + (quad * 2 + 0) * (COUT * 8) + lane;
#pragma unroll
          for (int c = 0; c < COUT; ++c) d[c * 8] = __uint_as_float(r[2 * COUT + c]);
        } else if (lane >= 24) {
          float* d = xb + (quad * 2 + 1) * (COUT * 8) + (lane - 24);
#pragma unroll
          for (int c = 0; c < COUT; ++c) d[c * 8] = __uint_as_float(r[c]);
        }
        group_bar_sync(grp);
        // out[row] = P_ky1[row] + P_ky0[row-1] + P_ky2[row+1]; rows are 8 lanes apart.  Branch-free: every lane shuffles, the
        // boundary lanes of the quadrant (first / last row) take the neighbouring quadrant's value from the exchange buffer
        const bool first = lane < 8, last = lane >= 24;
        const float* xsrc = first ? xb + ((quad > 0 ? quad - 1 : 0) * 2 + 1) * (COUT * 8) + lane
                                  : xb + ((quad < 3 ? quad + 1 : 3) * 2 + 0) * (COUT * 8) + (lane & 7);
        float v[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float xv = xsrc[c * 8];
          const float up_s = __shfl_up_sync(0xffffffffu, __uint_as_float(r[c]), 8);
          const float dn_s = __shfl_down_sync(0xffffffffu, __uint_as_float(r[2 * COUT + c]), 8);
          const float up = first ? xv : up_s, dn = last ? xv : dn_s;
          v[c] = (__uint_as_float(r[COUT + c]) + up) + dn + bias_s[c];
          if (p.relu) v[c] = fmaxf(v[c], 0.f);
        }
        if (!valid) continue;
        const long long vox = (long long)pl * HW + vox0;
        if (p.res) {
#pragma unroll
          for (int t = 0; t < TERMS; ++t)
#pragma unroll
            for (int cg = 0; cg < COUT / 8; ++cg) {
              const long long e = t * p.term_stride_out + (((long long)n * (COUT / 8) + cg) * DHW + vox) * 8;
              const int4 qv = __ldg(reinterpret_cast<const int4*>(p.res + e));
              unpack_bf16x8_add(qv, v + cg * 8);
            }
        }
#pragma unroll
        for (int cg = 0; cg < COUT / 8; ++cg) {
          const long long e = (((long long)n * (COUT / 8) + cg) * DHW + vox) * 8;
          float* vv = v + cg * 8;
          __nv_bfloat16 hi[8];
          int4 qh;
          uint32_t* qh32 = reinterpret_cast<uint32_t*>(&qh);
#pragma unroll
          for (int i = 0; i < 8; ++i) hi[i] = __float2bfloat16_rn(vv[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 t2 = __halves2bfloat162(hi[2 * i], hi[2 * i + 1]);
            qh32[i] = *reinterpret_cast<uint32_t*>(&t2);
          }
          *reinterpret_cast<int4*>(p.y + e) = qh;
          if (TERMS == 2) {
            int4 ql;
            uint32_t* ql32 = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              ql32[i] = pack_bf16x2(vv[2 * i] - __bfloat162float(hi[2 * i]), vv[2 * i + 1] - __bfloat162float(hi[2 * i + 1]));
            *reinterpret_cast<int4*>(p.y + p.term_stride_out + e) = ql;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 get_encode_fn3() {
  static EncodeTiledFn3 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
  fn = (EncodeTiledFn3)p;
  return fn;
}
static uint16_t f2bf_ys(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f_ys(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace pccgeo

using namespace pccgeo;

// B image per precision term: [kx (3)][kcore (2)][ngroup (18)][8 n][8 k] bf16, n = j*48 + ky*16 + co (j = stacked output plane:
// j=0 takes tap kz=2, j=1 kz=1, j=2 kz=0 of the equivalent correlation; all three axes flipped for a transposed conv).
extern "C" long long pccgeo_umma_ys_pack_weights_host(const float* w, void* out, int cin, int cout, int transposed, int terms) {
  if (cin <= 0 || cin > 16 || cout <= 0 || cout > 16 || (terms != 1 && terms != 2)) {
    set_error("umma_ys_pack_weights: needs <= 16 input and output channels, 1 or 2 terms");
    return PCCGEO_EINVAL;
  }
  const long long per_term = 3LL * 2 * (ys::NB / 8) * 64 * 2;
  if (!out) return per_term * terms;
  if (!w) { set_error("umma_ys_pack_weights: null weights"); return PCCGEO_EINVAL; }
  uint16_t* o = (uint16_t*)out;
  memset(o, 0, (size_t)per_term * terms);
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 3; ++j)
      for (int ky = 0; ky < 3; ++ky)
        for (int co = 0; co < cout; ++co)
          for (int ci = 0; ci < cin; ++ci) {
            int kz = 2 - j, kyy = ky, kxx = kx;
            if (transposed) { kz = 2 - kz; kyy = 2 - kyy; kxx = 2 - kxx; }
            const float val = w[((long long)((kz * 3 + kyy) * 3 + kxx) * cin + ci) * cout + co];
            const int nrow = j * ys::SC + ky * ys::COUT + co, kcore = ci >> 3, ki = ci & 7;
            const long long idx = ((((long long)kx * 2 + kcore) * (ys::NB / 8) + (nrow >> 3)) * 8 + (nrow & 7)) * 8 + ki;
            const uint16_t hi = f2bf_ys(val);
            o[idx] = hi;
            if (terms == 2) o[per_term / 2 + idx] = f2bf_ys(val - bf2f_ys(hi));
          }
  return per_term * terms;
}

extern "C" int pccgeo_conv3d_umma_ys(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                                     int n, int cin, int d, int h, int wd, int cout, int relu, int terms, void* stream) {
  using namespace ys;
  PCCGEO_REQUIRE(xb && wpacked && yb, "conv3d_umma_ys: null pointer");
  PCCGEO_REQUIRE(terms == 1 || terms == 2, "conv3d_umma_ys: terms must be 1 or 2");
  PCCGEO_REQUIRE(cin > 0 && cin <= 16 && cout > 0 && cout <= 16, "conv3d_umma_ys: %d -> %d channels unsupported (<= 16 each)", cin, cout);
  PCCGEO_REQUIRE(n > 0 && d > 0 && h > 0 && wd > 0 && wd % TX == 0, "conv3d_umma_ys: W must be a multiple of 8 (got %dx%dx%d)", d, h, wd);
  EncodeTiledFn3 enc = get_encode_fn3();
  PCCGEO_REQUIRE(enc, "conv3d_umma_ys: cuTensorMapEncodeTiled unavailable");
  Params p{};
  p.bias = bias; p.res = (const __nv_bfloat16*)residual_b; p.y = (__nv_bfloat16*)yb; p.wimg = (const uint8_t*)wpacked;
  p.N = n; p.D = d; p.H = h; p.W = wd; p.relu = relu; p.cout_real = cout;
  p.ytiles = (h + TYO - 1) / TYO; p.xtiles = wd / TX; p.items = n * p.ytiles * p.xtiles;
  p.wbytes_term = 3 * 2 * (NB / 8) * 128;
  p.term_stride_out = (long long)n * 16 * d * h * wd;
  const int stage_bytes = terms * 2 * PLANE_CG_BYTES;
  const int fixed = HEADER_BYTES + ((p.wbytes_term * terms + 127) & ~127) + XCH_BYTES;
  p.nstage = (113 * 1024 - fixed) / stage_bytes;   // two CTAs per SM
  if (p.nstage > MAX_STAGES) p.nstage = MAX_STAGES;
  PCCGEO_REQUIRE(p.nstage >= 3, "conv3d_umma_ys: shared memory budget");
  const size_t smem = fixed + (size_t)p.nstage * stage_bytes;

  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)wd * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)terms * n * 2};
  const cuuint64_t gstr[3] = {(cuuint64_t)wd * 16, (cuuint64_t)wd * h * 16, (cuuint64_t)wd * h * d * 16};
  const cuuint32_t box[4] = {PX * 8, TYI, 1, 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xb), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCCGEO_REQUIRE(cr == CUDA_SUCCESS, "conv3d_umma_ys: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  const int grid = p.items < 296 ? p.items : 296;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    PCCGEO_CUDA(cudaFuncSetAttribute(conv3d_umma_ys_kernel<1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    PCCGEO_CUDA(cudaFuncSetAttribute(conv3d_umma_ys_kernel<2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr_set = true;
  }
  if (terms == 2) conv3d_umma_ys_kernel<2, 5><<<grid, NUM_THREADS, smem, st>>>(tmap, p);
  else conv3d_umma_ys_kernel<1, 5><<<grid, NUM_THREADS, smem, st>>>(tmap, p);
  return check_launch("conv3d_umma_ys_kernel");
}
