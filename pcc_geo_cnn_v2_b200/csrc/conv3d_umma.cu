// tcgen05 / TMEM / TMA implicit-GEMM 3x3x3 convolution for sm_100a.
//
// Replaces Keras Conv3D / Conv3DTranspose (3,3,3) 'same' + BiasAdd + Relu + ResidualLayer add for the layers of
// AnalysisBlock / SynthesisBlock / the V2, progressive and hyper transforms
// (reference src/model_transforms.py:62-81,84-158).
//
// Data layout ("blocked"): activations are bf16 (term, N, C/8, D, H, W, 8): eight channels of one voxel are 16
// contiguous bytes, voxels along W follow at a 16-byte pitch.  That is exactly the UMMA *no-swizzle K-major*
// core-matrix layout (8 rows x 16 bytes, rows 16 B apart), so an A operand of 128 output voxels (16 y x 8 x)
// for ANY filter tap is just a different 16-byte-aligned start address into one halo'd input plane in shared
// memory: start = plane + (dy+1)*row_pitch + (dx+1)*16,  SBO = row_pitch (next y),  LBO = channel-group pitch.
// No im2col copy, no swizzle phase to keep consistent.
//
// Pipeline (one CTA, persistent over work items = (n, y-tile, x-tile) columns, streamed along z):
//   warp 0      TMA producer: one 4-D box load {10 x * 8 ch = 160 B rows, 18 y, 1 z, C/8 groups} per input plane and precision
//               term into a ring of stages; out-of-bounds coordinates are zero-filled = TF 'SAME' padding.
//   warp 1      MMA issuer (one elected thread): for input plane z and each (ky,kx,k-chunk) ONE tcgen05.mma with
//               N = 3*Cout whose B operand stacks the three z-taps, accumulating into three neighbouring output
//               planes that live side by side in a TMEM ring (columns = plane slot x Cout).  This cuts the
//               A-operand shared-memory reads -- the bound for small Cout -- by 3x.
//   warps 2..9  epilogue: two groups of four warps (one per TMEM lane quadrant) alternate over finished planes:
//               tcgen05.ld -> +bias -> ReLU -> +residual -> bf16 (hi[/lo]) -> 16-byte coalesced global stores.
// UP = 2 (stride-2 transposed conv, SynthesisBlock's first layer): output o = 2*i + j per dimension (TF 'SAME', even sizes),
//   so even outputs take taps j=0 (input i) and j=2 (input i-1), odd outputs tap j=1 (input i).  The M tile is 128 INPUT
//   voxels; the four (y,x) parity classes of an output plane sit side by side in one TMEM slot (4*Cout columns), input
//   plane z feeds output planes 2z, 2z+1, 2z+2 (taps kz = 0,1,2) and ONE MMA per input shift (dy,dx) in {0,-1}^2 carries
//   all 3 planes x 4 classes (N = 12*Cout; classes that have no tap at that shift get zero weights) -- 4 MMAs per
//   k-chunk and precision pair instead of the 27 small ones of the gather kernel.
// Precision terms: terms=1 plain bf16 operands; terms=2 splits activations and weights into hi+lo bf16 and
// issues a_hi*w_hi + a_hi*w_lo + a_lo*w_hi into the same fp32 accumulator (fp32-class accuracy, 3x MMA work).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {

// --------------------------------------------------------------------------------------------------------
// debug / tuning options (pccgeo_set_option)
// --------------------------------------------------------------------------------------------------------
static int g_opt_swap_lbo_sbo = 0;
static int g_opt_max_ctas = 0;
static int g_opt_one_cta = 0;
namespace zy { extern int g_zy_groups; }   // conv3d_umma_zy.cu

// --------------------------------------------------------------------------------------------------------
// kernel
// --------------------------------------------------------------------------------------------------------
constexpr int TY = 16, TX = 8;                 // output rows of one MMA: 16 y x 8 x = 128 voxels
constexpr int PY = TY + 2, PX = TX + 2;        // halo'd plane
constexpr int PLANE_CG_BYTES = PY * PX * 16;   // one channel group of one input plane: 2880 B
constexpr int ROW_PITCH = PX * 16;             // 160 B between y rows  (SBO of A)
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
// warps 0..7 epilogue, 8 TMA producer, 9 MMA issuer: the warp schedulers favour high warp ids, and the issuing thread's
// instruction stream is the critical path (conv3d_umma_zy.cu measured it: +8 % as the last warp instead of warp 1)
constexpr int W_TMA = NUM_EPI_WARPS, W_MMA = NUM_EPI_WARPS + 1;
constexpr int MAX_STAGES = 8, MAX_SLOTS = 32;

struct UmmaConvParams {
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  const uint8_t* wimg;   // packed B-operand image (global)
  int N, D, H, W;
  int CGi, CGo;          // channel groups (of 8) in / out
  int terms, relu, cout_real;
  int ytiles, xtiles, items;
  int nstage, nslots, slot_shift;
  int wbytes_term;       // bytes of one precision term of the weight image
  int wterms;            // precision terms of the weight image (HL mode: 1 -- hi and lo weights sit side by side in N)
  int swap;              // debug: swap LBO/SBO roles
  long long term_stride_out;  // elements between precision terms of y / res
};

struct __align__(8) SmemHeader {
  uint64_t in_full[MAX_STAGES], in_empty[MAX_STAGES];
  uint64_t acc_full[MAX_SLOTS], acc_empty[MAX_SLOTS];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int HEADER_BYTES = 1024;
static_assert(sizeof(SmemHeader) <= HEADER_BYTES, "header too large");

// padded output channels (16,32,64); precision terms (1,2); Cin/16 (1,2,4); UP = 1 (stride 1) or 2 (stride-2 transposed)
// small stride-1 configurations (16 -> 16 channels) fit twice on an SM: two CTAs interleave their MMA streams, which hides
// part of the fixed per-instruction cost of small-N tcgen05.mma (tools/umma_bench.cu: 57 instead of 68.5 cycles at N = 48)
template <int COUT, int KC, int UP>
constexpr int umma_ctas_per_sm() { return (UP == 1 && COUT == 16 && KC == 1) ? 2 : 1; }

// HL = 1 (COUT = 16, TERMS = 2, stride 1): the hi and lo halves of the weights are stacked in N next to the z-taps, so one
// MMA of N = 96 per activation term does the work of two of N = 48: per tap a_hi x [w_hi|w_lo] and a_lo x [w_hi|w_lo] (the
// fourth product a_lo*w_lo comes for free and only adds accuracy).  An output plane's slot holds two accumulators
// (x*w_hi | x*w_lo) that the epilogue adds.  18 MMAs per plane instead of 27: fewer A-operand fetches, the bound of small-N MMAs.
template <int COUT, int TERMS, int KC, int UP, int HL = 0>
__global__ void __launch_bounds__(NUM_THREADS, umma_ctas_per_sm<COUT, KC, UP>())
conv3d_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const UmmaConvParams p) {
  static_assert(!HL || (COUT == 16 && TERMS == 2 && UP == 1), "HL mode: 16 output channels, two terms, stride 1");
  constexpr int SC = UP == 2 ? 4 * COUT : (HL ? 2 * COUT : COUT);   // TMEM columns of one output-plane slot
  constexpr int NSHIFT = UP == 2 ? 4 : 9;         // distinct (y,x) input shifts = MMAs per k-chunk and precision pair
  extern __shared__ __align__(1024) uint8_t smem[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wbytes_all = p.wbytes_term * p.wterms;
  uint8_t* wsm = smem + HEADER_BYTES;
  const int stage_bytes = p.terms * p.CGi * PLANE_CG_BYTES;
  uint8_t* stages = wsm + ((wbytes_all + 127) & ~127);
  const int tmem_cols = p.nslots * SC;                        // power of two >= 32 (host guarantees)

  // ---- one-time setup ----
  if (warp == W_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int i = 0; i < p.nstage; ++i) { mbar_init(smem_u32(&hdr->in_full[i]), 1); mbar_init(smem_u32(&hdr->in_empty[i]), 1); }
    for (int i = 0; i < p.nslots; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 4); }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights: global -> shared with plain 16-byte copies, then make them visible to the async (UMMA) proxy
  for (int i = threadIdx.x * 16; i < wbytes_all; i += NUM_THREADS * 16)
    *reinterpret_cast<int4*>(wsm + i) = __ldg(reinterpret_cast<const int4*>(p.wimg + i));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, hdr->tmem_base, 0);  // shuffle => provably warp-uniform

  const int D = p.D;
  if (warp == W_TMA) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t s = 0, phase = 0;  // stage ring position / phase
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int xt = item % p.xtiles, yt = (item / p.xtiles) % p.ytiles, n = item / (p.xtiles * p.ytiles);
        for (int z = 0; z < D; ++z) {
          mbar_wait(smem_u32(&hdr->in_empty[s]), phase ^ 1);
          const uint32_t full = smem_u32(&hdr->in_full[s]);
          mbar_expect_tx(full, (uint32_t)stage_bytes);
          for (int t = 0; t < p.terms; ++t)
            tma_load_4d(smem_u32(stages + (size_t)s * stage_bytes + (size_t)t * p.CGi * PLANE_CG_BYTES), &tmap_x, full,
                        (xt * TX - 1) * 8, yt * TY - 1, z, (t * p.N + n) * p.CGi);
          if (++s == (uint32_t)p.nstage) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer =================
    // The whole warp walks the (uniform) loop so that address arithmetic stays on the uniform datapath; only the
    // tcgen05 instructions themselves are issued by one elected lane.  Nothing in the per-MMA path divides or
    // branches on data: per input plane we precompute <= 2 accumulator segments (split only where the TMEM ring
    // wraps) and every descriptor is a 64-bit add on a per-plane base.
    const uint32_t lbo_a = p.swap ? ROW_PITCH : PLANE_CG_BYTES, sbo_a = p.swap ? PLANE_CG_BYTES : ROW_PITCH;
    constexpr uint32_t b_kcore = 3 * (SC / 8) * 128;  // bytes between the two K core matrices of a B tile
    const uint32_t lbo_b = p.swap ? 128 : b_kcore, sbo_b = p.swap ? b_kcore : 128;
    constexpr uint32_t b_tile16 = 2 * b_kcore / 16;     // one (shift,kc) tile, in 16-byte units
    constexpr uint32_t b_plane16 = (SC / 8) * 128 / 16; // one stacked output plane (SC rows), in 16-byte units
    constexpr int npairs = HL ? 2 : (TERMS == 2 ? 3 : 1);   // HL: (a_hi, a_lo) x one stacked weight tile
    const uint32_t slot_mask = p.nslots - 1;             // nslots is a power of two
    const uint64_t bdesc0 = make_smem_desc(smem_u32(wsm), lbo_b, sbo_b);
    const uint32_t b_lo0 = (uint32_t)bdesc0, b_hi = (uint32_t)(bdesc0 >> 32);
    const uint64_t adesc_proto = make_smem_desc(0, lbo_a, sbo_a);
    const uint32_t a_hi = (uint32_t)(adesc_proto >> 32), a_lo_proto = (uint32_t)adesc_proto;
    const uint32_t a_term16 = p.CGi * PLANE_CG_BYTES / 16, w_term16 = p.wbytes_term / 16;
    const uint32_t stage16 = stage_bytes / 16, stages16 = smem_u32(stages) / 16;
    const int Dout = D * UP;
    uint32_t s = 0, in_phase = 0;  // input stage ring position / phase
    uint32_t g0 = 0;               // running output-plane counter at the start of this item
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, g0 += Dout) {
      for (int z = 0; z < D; ++z) {
        mbar_wait(smem_u32(&hdr->in_full[s]), in_phase);
        // output planes touched by input plane z: pfirst .. pfirst+2, clipped to the volume
        const int pfirst = UP == 2 ? 2 * z : z - 1;
        const int pa = pfirst > 0 ? pfirst : 0, pb = pfirst + 2 < Dout ? pfirst + 2 : Dout - 1;
        // planes touched for the first time by this input plane (UP=1: z+1; UP=2: 2z+1, 2z+2; plus plane 0 when z == 0).
        // Their TMEM slots were zeroed by the epilogue when it drained the previous tenant (or at kernel start), so every
        // MMA below accumulates and all MMAs of a plane have one shape: no accumulate=0 special cases that drain the pipe.
        auto wait_empty = [&](uint32_t g) { mbar_wait(smem_u32(&hdr->acc_empty[g & slot_mask]), (g >> p.slot_shift) & 1); };
        if (z == 0) wait_empty(g0);
        if (UP == 2) {
          wait_empty(g0 + 2 * z + 1);
          if (z + 1 < D) wait_empty(g0 + 2 * z + 2);
        } else {
          if (z + 1 < D) wait_empty(g0 + z + 1);
        }
        tc_fence_after();
        const uint32_t a_lo0 = a_lo_proto + stages16 + s * stage16;
        // steady-state segments (accumulate): planes pa..pb, split where the ring wraps
        const uint32_t slot_a = (g0 + pa) & slot_mask;
        const int nplanes = pb - pa + 1;
        const int n0 = min(nplanes, (int)(p.nslots - slot_a));   // planes before the wrap
        const uint32_t j0 = pa - pfirst;                          // stacked index of plane pa
        const uint32_t seg_d0 = tmem_base + slot_a * SC, seg_b0 = b_lo0 + j0 * b_plane16, seg_i0 = make_idesc(n0 * SC);
        const uint32_t seg_d1 = tmem_base, seg_b1 = b_lo0 + (j0 + n0) * b_plane16, seg_i1 = make_idesc((nplanes - n0) * SC);
        const bool two = n0 < nplanes;
        if (elect_one()) {
          if (!two) {
#pragma unroll
            for (int sh = 0; sh < NSHIFT; ++sh) {
              // UP=1: taps (ky,kx) read the halo'd plane at (ky,kx); UP=2: shifts (dy,dx) in {0,-1} read it at (1+dy,1+dx)
              const uint32_t a_sh16 = UP == 2 ? ((1 - sh / 2) * ROW_PITCH + (1 - sh % 2) * 16) / 16
                                              : ((sh / 3) * ROW_PITCH + (sh % 3) * 16) / 16;
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
                for (int pr = 0; pr < npairs; ++pr) {
                  const uint32_t ta = HL ? pr : (pr == 2 ? 1 : 0), tb = HL ? 0 : (pr == 1 ? 1 : 0);
                  umma_bf16_lh(seg_d0, a_lo0 + a_sh16 + ta * a_term16 + (uint32_t)(2 * kc) * (PLANE_CG_BYTES / 16), a_hi,
                               seg_b0 + tb * w_term16 + (uint32_t)(sh * KC + kc) * b_tile16, b_hi, seg_i0, 1u);
                }
              }
            }
          } else {
#pragma unroll
            for (int sh = 0; sh < NSHIFT; ++sh) {
              const uint32_t a_sh16 = UP == 2 ? ((1 - sh / 2) * ROW_PITCH + (1 - sh % 2) * 16) / 16
                                              : ((sh / 3) * ROW_PITCH + (sh % 3) * 16) / 16;
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
                for (int pr = 0; pr < npairs; ++pr) {
                  const uint32_t ta = HL ? pr : (pr == 2 ? 1 : 0), tb = HL ? 0 : (pr == 1 ? 1 : 0);
                  const uint32_t a_lo = a_lo0 + a_sh16 + ta * a_term16 + (uint32_t)(2 * kc) * (PLANE_CG_BYTES / 16);
                  const uint32_t b_off = tb * w_term16 + (uint32_t)(sh * KC + kc) * b_tile16;
                  umma_bf16_lh(seg_d0, a_lo, a_hi, seg_b0 + b_off, b_hi, seg_i0, 1u);
                  umma_bf16_lh(seg_d1, a_lo, a_hi, seg_b1 + b_off, b_hi, seg_i1, 1u);
                }
              }
            }
          }
          umma_commit(smem_u32(&hdr->in_empty[s]));  // input stage may be refilled once these MMAs retire
          if (UP == 2) {
            // planes 2z and 2z+1 are complete (2z+2 still needs tap 0 of input plane z+1)
            umma_commit(smem_u32(&hdr->acc_full[(g0 + 2 * z) & slot_mask]));
            umma_commit(smem_u32(&hdr->acc_full[(g0 + 2 * z + 1) & slot_mask]));
          } else {
            if (z >= 1) umma_commit(smem_u32(&hdr->acc_full[(g0 + z - 1) & slot_mask]));
            if (z == D - 1) umma_commit(smem_u32(&hdr->acc_full[(g0 + z) & slot_mask]));
          }
        }
        __syncwarp();
        if (++s == (uint32_t)p.nstage) { s = 0; in_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int ew = warp;              // 0..7
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access (warp id % 4)
    const int grp = ew >> 2;          // planes with (g & 1) == grp
    const int row = quad * 32 + lane; // M row = TMEM lane
    const int yl = row >> 3, xl = row & 7;
    float bias_r[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias_r[c] = (p.bias && c < p.cout_real) ? __ldg(p.bias + c) : 0.f;
    const int Dout = D * UP, Wo = p.W * UP;
    const long long HWo = (long long)p.H * UP * Wo, DHWo = HWo * Dout;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    // zero every accumulator slot this group owns (slot parity == group) and publish it as empty: completes phase 0 of
    // acc_empty, which is what the MMA issuer waits for before the first use of a slot
    for (int slot = grp; slot < p.nslots; slot += 2) {
#pragma unroll
      for (int c = 0; c < SC; c += 16) tmem_st16_zero(lane_base + (uint32_t)slot * SC + c);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[slot]));
    }
    uint32_t g0 = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, g0 += Dout) {
      const int xt = item % p.xtiles, yt = (item / p.xtiles) % p.ytiles, n = item / (p.xtiles * p.ytiles);
      const long long vox0 = (long long)((yt * TY + yl) * UP) * Wo + (xt * TX + xl) * UP;
      for (int pl = 0; pl < Dout; ++pl) {
        const uint32_t g = g0 + pl;
        if ((int)(g & 1) != grp) continue;
        const int slot = g & (p.nslots - 1);
        mbar_wait(smem_u32(&hdr->acc_full[slot]), (g >> p.slot_shift) & 1);
        tc_fence_after();
        uint32_t r[SC];
#pragma unroll
        for (int c = 0; c < SC; c += 16) tmem_ld16(lane_base + (uint32_t)slot * SC + c, r + c);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < SC; c += 16) tmem_st16_zero(lane_base + (uint32_t)slot * SC + c);
        tmem_st_wait();  // the slot is handed back zeroed: its next tenant only ever accumulates
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[slot]));  // accumulator slot is free again
        if constexpr (UP == 2) {
          // this input voxel's 2x2 outputs of plane pl: for each output row (yc) the two x-neighbours (xc = 0, 1) of one
          // channel group are 32 contiguous bytes -> one 256-bit store per lane, 8 lanes = one full 256-byte run
#pragma unroll
          for (int yc = 0; yc < 2; ++yc) {
            const long long vox = (long long)pl * HWo + vox0 + (long long)yc * Wo;
#pragma unroll
            for (int cg = 0; cg < COUT / 8; ++cg) {
              uint32_t qh[8], ql[8];
#pragma unroll
              for (int xc = 0; xc < 2; ++xc)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int col = (yc * 2 + xc) * COUT + cg * 8 + 2 * i;
                  float v0 = __uint_as_float(r[col]) + bias_r[cg * 8 + 2 * i], v1 = __uint_as_float(r[col + 1]) + bias_r[cg * 8 + 2 * i + 1];
                  if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                  __nv_bfloat162 t2 = __halves2bfloat162(h0, h1);
                  qh[xc * 4 + i] = *reinterpret_cast<uint32_t*>(&t2);
                  if (TERMS == 2) ql[xc * 4 + i] = pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
                }
              const long long e = (((long long)n * p.CGo + cg) * DHWo + vox) * 8;
              st_global_v8(p.y + e, qh);
              if (TERMS == 2) st_global_v8(p.y + p.term_stride_out + e, ql);
            }
          }
        } else {
          float v[COUT];
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            v[c] = __uint_as_float(r[c]) + bias_r[c];
            if (HL) v[c] = (__uint_as_float(r[c]) + __uint_as_float(r[(HL ? COUT : 0) + c])) + bias_r[c];
            if (p.relu) v[c] = fmaxf(v[c], 0.f);
          }
          const long long vox = (long long)pl * HWo + vox0;
          if (p.res) {
            for (int t = 0; t < p.terms; ++t)
#pragma unroll
              for (int cg = 0; cg < COUT / 8; ++cg) {
                const long long e = t * p.term_stride_out + (((long long)n * p.CGo + cg) * DHWo + vox) * 8;
                const int4 qv = __ldg(reinterpret_cast<const int4*>(p.res + e));
                unpack_bf16x8_add(qv, v + cg * 8);
              }
          }
#pragma unroll
          for (int cg = 0; cg < COUT / 8; ++cg) {
            const long long e = (((long long)n * p.CGo + cg) * DHWo + vox) * 8;
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
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------
// layout conversion kernels
// --------------------------------------------------------------------------------------------------------
__global__ void f32_to_blocked_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, int N, int C, int CG,
                                      long long DHW, int terms) {
  const long long total = (long long)N * CG * DHW;
  const long long term_stride = total * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i % DHW;
    const int cg = (int)((i / DHW) % CG);
    const int n = (int)(i / (DHW * CG));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      v[j] = c < C ? x[((long long)n * C + c) * DHW + vox] : 0.f;
    }
    int4 qh, ql;
    uint32_t* qh32 = reinterpret_cast<uint32_t*>(&qh);
    uint32_t* ql32 = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
      __nv_bfloat162 t2 = __halves2bfloat162(h0, h1);
      qh32[j] = *reinterpret_cast<uint32_t*>(&t2);
      ql32[j] = pack_bf16x2(v[2 * j] - __bfloat162float(h0), v[2 * j + 1] - __bfloat162float(h1));
    }
    *reinterpret_cast<int4*>(xb + i * 8) = qh;
    if (terms == 2) *reinterpret_cast<int4*>(xb + term_stride + i * 8) = ql;
  }
}

// Phase split + layout conversion in one pass (the stride-2 layers' weight gradient, training.py::_wgrad_stride2): the eight phase
// volumes L_p[b] = L[2b + p] of x (N, Cb, 2S, 2S, 2S) stacked as channels and cut into chunks of C channels (C a multiple of Cb,
// 8*Cb a multiple of C): out[chunk][term][n][C/8][S][S][S][8], channel c of chunk j = phase j*(C/Cb) + c / Cb, source channel c % Cb.
__global__ void f32_phases_to_blocked_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, int N, int Cb, int C, int S0, int S1,
                                             int S2, int nch, int terms) {
  const int CG = C / 8, ppc = C / Cb;
  const long long V = (long long)S0 * S1 * S2, total = (long long)nch * N * CG * V;
  const long long term_stride = (long long)N * CG * V * 8, chunk_stride = term_stride * terms;
  const long long LV = 8 * V, L2 = 2LL * S2, L12 = 4LL * S1 * S2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i % V;
    const int cg = (int)((i / V) % CG);
    const int n = (int)((i / (V * CG)) % N);
    const int j = (int)(i / (V * CG * N));
    const int xx = (int)(vox % S2), yy = (int)((vox / S2) % S1), zz = (int)(vox / ((long long)S1 * S2));
    const int c0 = cg * 8, phase = j * ppc + c0 / Cb, cb0 = c0 % Cb;   // the 8 channels of a group share their phase (Cb >= 8)
    const int pz = phase >> 2, py = (phase >> 1) & 1, px = phase & 1;
    const float* src = x + ((long long)n * Cb + cb0) * LV + (2 * zz + pz) * L12 + (2 * yy + py) * L2 + (2 * xx + px);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(src + k * LV);
    int4 qh, ql;
    uint32_t* qh32 = reinterpret_cast<uint32_t*>(&qh);
    uint32_t* ql32 = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
      __nv_bfloat162 t2 = __halves2bfloat162(h0, h1);
      qh32[k] = *reinterpret_cast<uint32_t*>(&t2);
      ql32[k] = pack_bf16x2(v[2 * k] - __bfloat162float(h0), v[2 * k + 1] - __bfloat162float(h1));
    }
    __nv_bfloat16* dst = xb + j * chunk_stride + (((long long)n * CG + cg) * V + vox) * 8;
    *reinterpret_cast<int4*>(dst) = qh;
    if (terms == 2) *reinterpret_cast<int4*>(dst + term_stride) = ql;
  }
}

__global__ void blocked_to_f32_kernel(const __nv_bfloat16* __restrict__ xb, float* __restrict__ x, int N, int C, int CG,
                                      long long DHW, int terms) {
  const long long total = (long long)N * CG * DHW;
  const long long term_stride = total * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long vox = i % DHW;
    const int cg = (int)((i / DHW) % CG);
    const int n = (int)(i / (DHW * CG));
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < terms; ++t) {
      const int4 q = *reinterpret_cast<const int4*>(xb + t * term_stride + i * 8);
      unpack_bf16x8_add(q, v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      if (c < C) x[((long long)n * C + c) * DHW + vox] = v[j];
    }
  }
}

// --------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

static inline int round_up_i(int a, int m) { return (a + m - 1) / m * m; }

static uint16_t f32_to_bf16_rn_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f32_host(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

template <int COUT, int TERMS, int KC, int UP, int HL = 0>
static int launch_umma(const CUtensorMap& tmap, const UmmaConvParams& p, size_t smem, int grid, cudaStream_t st) {
  static bool attr_set = false;
  static size_t attr_smem = 0;
  if (!attr_set || smem > attr_smem) {
    PCCGEO_CUDA(cudaFuncSetAttribute(conv3d_umma_kernel<COUT, TERMS, KC, UP, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
    attr_smem = 227 * 1024;
  }
  conv3d_umma_kernel<COUT, TERMS, KC, UP, HL><<<grid, NUM_THREADS, smem, st>>>(tmap, p);
  return check_launch("conv3d_umma_kernel");
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" int pccgeo_set_option(const char* name, long long value) {
  if (!name) return PCCGEO_EINVAL;
  if (!strcmp(name, "umma_swap_lbo_sbo")) { g_opt_swap_lbo_sbo = (int)value; return PCCGEO_OK; }
  if (!strcmp(name, "umma_max_ctas")) { g_opt_max_ctas = (int)value; return PCCGEO_OK; }
  if (!strcmp(name, "umma_one_cta_per_sm")) { g_opt_one_cta = (int)value; return PCCGEO_OK; }
  if (!strcmp(name, "zy_groups") && (value == 2 || value == 3)) { zy::g_zy_groups = (int)value; return PCCGEO_OK; }
  set_error("set_option: unknown option %s", name);
  return PCCGEO_EINVAL;
}

extern "C" int pccgeo_f32_to_blocked(const float* x, void* xb, int n, int c, int d, int h, int wd, int terms, void* stream) {
  PCCGEO_REQUIRE(x && xb && n > 0 && c > 0 && d > 0 && h > 0 && wd > 0 && (terms == 1 || terms == 2), "f32_to_blocked: bad argument");
  const int CG = round_up_i(c, 16) / 8;
  const long long DHW = (long long)d * h * wd, total = (long long)n * CG * DHW;
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  f32_to_blocked_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)xb, n, c, CG, DHW, terms);
  return check_launch("f32_to_blocked_kernel");
}

extern "C" int pccgeo_f32_phases_to_blocked(const float* x, void* xb, int n, int cb, int c, int d, int h, int wd, int terms, void* stream) {
  PCCGEO_REQUIRE(x && xb && n > 0 && cb >= 8 && cb % 8 == 0 && c > 0 && c % 16 == 0 && c % cb == 0 && (8 * cb) % c == 0 && d > 0 && h > 0 && wd > 0 &&
                     (terms == 1 || terms == 2),
                 "f32_phases_to_blocked: bad argument (Cb a multiple of 8, C a multiple of 16 and of Cb that divides 8 * Cb)");
  const int nch = 8 * cb / c;
  const long long total = (long long)nch * n * (c / 8) * d * h * wd;
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  f32_phases_to_blocked_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)xb, n, cb, c, d, h, wd, nch, terms);
  return check_launch("f32_phases_to_blocked_kernel");
}

extern "C" int pccgeo_blocked_to_f32(const void* xb, float* x, int n, int c, int d, int h, int wd, int terms, void* stream) {
  PCCGEO_REQUIRE(x && xb && n > 0 && c > 0 && d > 0 && h > 0 && wd > 0 && (terms == 1 || terms == 2), "blocked_to_f32: bad argument");
  const int CG = round_up_i(c, 16) / 8;
  const long long DHW = (long long)d * h * wd, total = (long long)n * CG * DHW;
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  blocked_to_f32_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)xb, x, n, c, CG, DHW, terms);
  return check_launch("blocked_to_f32_kernel");
}

// B image, per precision term: [shift][kc (Cin_p/16)][kcore (2)][ngroup (3*SC/8)][8 n][8 k] bf16, k = channel kc*16 + kcore*8 + ki.
//   stride 1 (UP=1): shift = ky*3+kx (9), SC = Cout_p, n = j*SC + co stacks the three z-taps (j=0: kz=2, j=1: kz=1, j=2: kz=0).
//   stride-2 transposed (UP=2): shift = sy*2+sx (4) with input offset -sy / -sx, SC = 4*Cout_p,
//     n = j*SC + (yc*2+xc)*Cout_p + co: output plane 2z+j takes tap kz = j; output parity class (yc,xc) takes tap
//     k = 0 at offset 0 / k = 2 at offset -1 when even, k = 1 at offset 0 when odd (zero rows where a class has no tap).
static long long umma_pack_weights_impl(const float* w, void* out, int cin, int cout, int stride, int transposed, int terms, int hl) {
  if (cin <= 0 || cout <= 0 || (terms != 1 && terms != 2)) { set_error("umma_pack_weights: bad argument"); return PCCGEO_EINVAL; }
  const bool up2 = stride == 2 && transposed;
  if (hl) {
    // HL image (one "term"): [shift 9][kcore 2][ngroup 12][8 n][8 k]; n = j*32 + half*16 + co with half 0 = bf16(w), 1 = w - bf16(w)
    if (stride != 1 || cin > 16 || cout > 16 || terms != 2) { set_error("umma_pack_weights_hl: stride-1 layers with <= 16 channels and two terms only"); return PCCGEO_EINVAL; }
    const long long bytes = 9LL * 2 * 12 * 64 * 2;
    if (!out) return bytes;
    if (!w) { set_error("umma_pack_weights_hl: null weights"); return PCCGEO_EINVAL; }
    uint16_t* o = (uint16_t*)out;
    memset(o, 0, (size_t)bytes);
    for (int kyx = 0; kyx < 9; ++kyx)
      for (int j = 0; j < 3; ++j)
        for (int co = 0; co < cout; ++co)
          for (int ci = 0; ci < cin; ++ci) {
            int kz = 2 - j, ky = kyx / 3, kx = kyx % 3;
            if (transposed) { kz = 2 - kz; ky = 2 - ky; kx = 2 - kx; }
            const float val = w[((long long)((kz * 3 + ky) * 3 + kx) * cin + ci) * cout + co];
            const uint16_t hi = f32_to_bf16_rn_host(val), lo = f32_to_bf16_rn_host(val - bf16_to_f32_host(hi));
            const int kcore = ci >> 3, ki = ci & 7;
            for (int half = 0; half < 2; ++half) {
              const int nrow = j * 32 + half * 16 + co;
              o[((((long long)kyx * 2 + kcore) * 12 + (nrow >> 3)) * 8 + (nrow & 7)) * 8 + ki] = half ? lo : hi;
            }
          }
    return bytes;
  }
  if (stride != 1 && !up2) { set_error("umma_pack_weights: stride-%d forward convs are not supported by the TMA kernel", stride); return PCCGEO_EINVAL; }
  const int cip = round_up_i(cin, 16), cop = round_up_i(cout, 16), KC = cip / 16;
  const int SC = up2 ? 4 * cop : cop, nshift = up2 ? 4 : 9;
  const long long per_term = (long long)nshift * KC * 2 * (3 * SC / 8) * 64 * 2;
  if (!out) return per_term * terms;
  if (!w) { set_error("umma_pack_weights: null weights"); return PCCGEO_EINVAL; }
  uint16_t* o = (uint16_t*)out;
  memset(o, 0, (size_t)per_term * terms);
  auto put = [&](int sh, int nrow, int ci, float val) {
    const int kc = ci / 16, kk = ci % 16, kcore = kk >> 3, ki = kk & 7;
    const long long idx = ((((long long)(sh * KC + kc) * 2 + kcore) * (3 * SC / 8) + (nrow >> 3)) * 8 + (nrow & 7)) * 8 + ki;
    const uint16_t hi = f32_to_bf16_rn_host(val);
    o[idx] = hi;
    if (terms == 2) o[per_term / 2 + idx] = f32_to_bf16_rn_host(val - bf16_to_f32_host(hi));
  };
  if (!up2) {
    for (int kyx = 0; kyx < 9; ++kyx)
      for (int j = 0; j < 3; ++j)
        for (int co = 0; co < cout; ++co)
          for (int ci = 0; ci < cin; ++ci) {
            int kz = 2 - j, ky = kyx / 3, kx = kyx % 3;
            if (transposed) { kz = 2 - kz; ky = 2 - ky; kx = 2 - kx; }  // stride-1 transposed conv == conv with flipped taps
            put(kyx, j * cop + co, ci, w[((long long)((kz * 3 + ky) * 3 + kx) * cin + ci) * cout + co]);
          }
  } else {
    // tap of output parity c at input offset -s: (c=0,s=0) -> 0, (c=0,s=1) -> 2, (c=1,s=0) -> 1, (c=1,s=1) -> none
    auto tap = [](int c, int s) { return c == 0 ? (s == 0 ? 0 : 2) : (s == 0 ? 1 : -1); };
    for (int sy = 0; sy < 2; ++sy)
      for (int sx = 0; sx < 2; ++sx)
        for (int j = 0; j < 3; ++j)
          for (int yc = 0; yc < 2; ++yc)
            for (int xc = 0; xc < 2; ++xc) {
              const int ky = tap(yc, sy), kx = tap(xc, sx);
              if (ky < 0 || kx < 0) continue;
              for (int co = 0; co < cout; ++co)
                for (int ci = 0; ci < cin; ++ci)
                  put(sy * 2 + sx, j * SC + (yc * 2 + xc) * cop + co, ci, w[((long long)((j * 3 + ky) * 3 + kx) * cin + ci) * cout + co]);
            }
  }
  return per_term * terms;
}

extern "C" long long pccgeo_umma_pack_weights_host(const float* w, void* out, int cin, int cout, int stride, int transposed,
                                                   int terms) {
  return umma_pack_weights_impl(w, out, cin, cout, stride, transposed, terms, 0);
}

extern "C" long long pccgeo_umma_hl_pack_weights_host(const float* w, void* out, int cin, int cout, int transposed) {
  return umma_pack_weights_impl(w, out, cin, cout, 1, transposed, 2, 1);
}

static int conv3d_umma_impl(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                            int n, int cin, int d, int h, int wd, int cout, int stride, int transposed, int relu,
                            int terms, int hl, void* stream) {
  PCCGEO_REQUIRE(xb && wpacked && yb, "conv3d_umma: null pointer");
  PCCGEO_REQUIRE(terms == 1 || terms == 2, "conv3d_umma: terms must be 1 or 2");
  const bool up2 = stride == 2 && transposed;  // stride 1: taps are already flipped in the packed image
  PCCGEO_REQUIRE(stride == 1 || up2, "conv3d_umma: stride-%d forward convs are not supported by the TMA kernel", stride);
  PCCGEO_REQUIRE(!up2 || !residual_b, "conv3d_umma: no fused residual for stride-2 transposed layers");
  PCCGEO_REQUIRE(n > 0 && d > 0 && h % TY == 0 && wd % TX == 0 && h > 0 && wd > 0, "conv3d_umma: H must be a multiple of 16 and W of 8 (got %dx%dx%d)", d, h, wd);
  const int cip = round_up_i(cin, 16), cop = round_up_i(cout, 16);
  PCCGEO_REQUIRE(cop == 16 || cop == 32 || cop == 64, "conv3d_umma: Cout %d unsupported", cout);
  PCCGEO_REQUIRE(cip == 16 || cip == 32 || cip == 64, "conv3d_umma: Cin %d unsupported", cin);
  EncodeTiledFn enc = get_encode_fn();
  PCCGEO_REQUIRE(enc, "conv3d_umma: cuTensorMapEncodeTiled unavailable");

  UmmaConvParams p{};
  p.bias = bias; p.res = (const __nv_bfloat16*)residual_b; p.y = (__nv_bfloat16*)yb; p.wimg = (const uint8_t*)wpacked;
  p.N = n; p.D = d; p.H = h; p.W = wd; p.CGi = cip / 8; p.CGo = cop / 8; p.terms = terms; p.relu = relu; p.cout_real = cout;
  p.ytiles = h / TY; p.xtiles = wd / TX; p.items = n * p.ytiles * p.xtiles;
  PCCGEO_REQUIRE(!hl || (!up2 && cip == 16 && cop == 16 && terms == 2), "conv3d_umma_hl: stride-1 layers with <= 16 channels and two terms only");
  const int SC = up2 ? 4 * cop : (hl ? 2 * cop : cop);
  PCCGEO_REQUIRE(3 * SC <= 256, "conv3d_umma: %d output channels are too many for one stride-2 transposed MMA", cout);
  p.wbytes_term = (up2 ? 4 : 9) * (cip / 16) * 2 * (3 * SC / 8) * 128;
  p.wterms = hl ? 1 : terms;
  p.swap = g_opt_swap_lbo_sbo;
  p.term_stride_out = (long long)n * cop * d * h * wd * (up2 ? 8 : 1);
  // TMEM ring: power of two, >= 4 planes; stride 1 uses up to 256 columns, the stride-2 transposed form all 512
  p.nslots = (up2 ? 512 : 256) / SC;
  if (p.nslots > MAX_SLOTS) p.nslots = MAX_SLOTS;
  p.slot_shift = 0;
  while ((1 << p.slot_shift) < p.nslots) ++p.slot_shift;
  const int stage_bytes = terms * p.CGi * PLANE_CG_BYTES;
  const int wall = (p.wbytes_term * p.wterms + 127) & ~127;
  const int ctas_per_sm = (!up2 && cop == 16 && cip == 16 && !g_opt_one_cta) ? 2 : 1;
  const int avail = (ctas_per_sm == 2 ? 113 : 227) * 1024 - HEADER_BYTES - wall;
  PCCGEO_REQUIRE(avail >= 3 * stage_bytes, "conv3d_umma: weights (%d B) leave no room for the input pipeline", wall);
  p.nstage = avail / stage_bytes;
  if (p.nstage > MAX_STAGES) p.nstage = MAX_STAGES;
  const size_t smem = HEADER_BYTES + wall + (size_t)p.nstage * stage_bytes;

  CUtensorMap tmap;
  // the (x, 8-channel) pair is one contiguous run in the blocked layout: make it the (wide) innermost TMA dimension so a
  // halo'd row is ONE 160-byte request instead of ten 16-byte ones; a negative / past-the-end start is zero-filled.
  const cuuint64_t gdim[4] = {(cuuint64_t)wd * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)terms * n * p.CGi};
  const cuuint64_t gstr[3] = {(cuuint64_t)wd * 16, (cuuint64_t)wd * h * 16, (cuuint64_t)wd * h * d * 16};
  const cuuint32_t box[4] = {PX * 8, PY, 1, (cuuint32_t)p.CGi};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xb), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCCGEO_REQUIRE(cr == CUDA_SUCCESS, "conv3d_umma: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  int grid = p.items < 148 * ctas_per_sm ? p.items : 148 * ctas_per_sm;
  if (g_opt_max_ctas > 0 && grid > g_opt_max_ctas) grid = g_opt_max_ctas;
  cudaStream_t st = (cudaStream_t)stream;
  const int kc = cip / 16;
  if (hl) return launch_umma<16, 2, 1, 1, 1>(tmap, p, smem, grid, st);
  if (up2) {
#define PCCGEO_DISPATCH_UP2(CO, T, K) if (cop == CO && terms == T && kc == K) return launch_umma<CO, T, K, 2>(tmap, p, smem, grid, st);
    PCCGEO_DISPATCH_UP2(16, 1, 1) PCCGEO_DISPATCH_UP2(16, 1, 2) PCCGEO_DISPATCH_UP2(16, 1, 4)
    PCCGEO_DISPATCH_UP2(16, 2, 1) PCCGEO_DISPATCH_UP2(16, 2, 2)
#undef PCCGEO_DISPATCH_UP2
    set_error("conv3d_umma: unsupported stride-2 transposed configuration %d -> %d", cin, cout);
    return PCCGEO_EINVAL;
  }
#define PCCGEO_DISPATCH(CO, T, K) if (cop == CO && terms == T && kc == K) return launch_umma<CO, T, K, 1>(tmap, p, smem, grid, st);
  PCCGEO_DISPATCH(16, 1, 1) PCCGEO_DISPATCH(16, 1, 2) PCCGEO_DISPATCH(16, 1, 4)
  PCCGEO_DISPATCH(32, 1, 1) PCCGEO_DISPATCH(32, 1, 2) PCCGEO_DISPATCH(32, 1, 4)
  PCCGEO_DISPATCH(64, 1, 1) PCCGEO_DISPATCH(64, 1, 2) PCCGEO_DISPATCH(64, 1, 4)
  PCCGEO_DISPATCH(16, 2, 1) PCCGEO_DISPATCH(16, 2, 2) PCCGEO_DISPATCH(16, 2, 4)
  PCCGEO_DISPATCH(32, 2, 1) PCCGEO_DISPATCH(32, 2, 2) PCCGEO_DISPATCH(32, 2, 4)
  PCCGEO_DISPATCH(64, 2, 1) PCCGEO_DISPATCH(64, 2, 2) PCCGEO_DISPATCH(64, 2, 4)
#undef PCCGEO_DISPATCH
  set_error("conv3d_umma: unsupported channel configuration %d -> %d", cin, cout);
  return PCCGEO_EINVAL;
}

extern "C" int pccgeo_conv3d_umma(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                                  int n, int cin, int d, int h, int wd, int cout, int stride, int transposed, int relu,
                                  int terms, void* stream) {
  return conv3d_umma_impl(xb, wpacked, bias, residual_b, yb, n, cin, d, h, wd, cout, stride, transposed, relu, terms, 0, stream);
}

extern "C" int pccgeo_conv3d_umma_hl(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                                     int n, int cin, int d, int h, int wd, int cout, int transposed, int relu, void* stream) {
  (void)transposed;   // the taps are already flipped in the packed image
  return conv3d_umma_impl(xb, wpacked, bias, residual_b, yb, n, cin, d, h, wd, cout, 1, 0, relu, 2, 1, stream);
}
