// tcgen05 / TMEM / TMA implicit-GEMM 3x3x3 stride-1 convolution for 16-channel layers with BOTH the z and the y taps
// accumulated by the tensor core in a two-dimensional TMEM ring ("zy-ring" form).
//
// Replaces Keras Conv3D / Conv3DTranspose (3,3,3) 'same' + BiasAdd + Relu + ResidualLayer add for the second and third layer
// of AnalysisBlock / SynthesisBlock at 16 filters (reference src/model_transforms.py:62-81) -- s.b2.t1 / s.b2.t2 at 64^3 are
// the dominant layers of the c3p synthesis transform (38 % of the encode+decode step in round 1).
//
// Why: with 16 output channels an MMA of the z-stacked kernel (conv3d_umma.cu) has N = 48: 24 cycles of math behind a
// 4 KB A-operand fetch from shared memory -- the kernel sits on the shared-memory operand bandwidth (ncu: TC smem-read pipe
// 83 %, tensor pipe 42 %).  Here the M tile has no y extent at all: its 128 rows are 16 blocks x 8 x-voxels of ONE input row
// (z, y).  That row contributes to the 3 x 3 output rows (z-1..z+1, y-1..y+1), which live side by side in TMEM:
//     column(y', z') = ((y' - y0) * 3 + (z' + 1) mod 3) * 16 + co
// so the nine (dy, dz) taps of one x-offset are ONE MMA with N = 144 whose D operand is the contiguous 144-column window
// of rows y-1..y+1 (all three z-slots of a row are adjacent; which slot holds which plane rotates with z mod 3, so the weight
// image is stored in the three rotations).  9 MMAs (3 x-offsets x 3 precision products) of N = 144 replace 81 of N = 48
// per 3 input rows: a third of the A fetches per MAC, measured 2.0 PFLOP/s (executed) from one issuing thread per SM
// (tools/umma_power.cu) against 1.2 for the N = 48 stream of two CTAs.  The y-halo costs nothing in MMA width: the first and
// last input rows of a tile use narrower windows (N = 48 / 96) instead of computing rows that would be thrown away.
//
// Work item (one CTA each, no persistence: the TMEM ring bookkeeping then needs no cross-item state): 16 consecutive blocks
// x one 8-voxel x segment x a tile of <= 10 output rows, streamed over all z.  Virtual output planes -1 and D take the taps
// that fall outside the volume and are drained and discarded, so every input plane issues the same MMAs.
//   warp 0      TMA producer: per input row and precision term one 5-D box {10 x * 8 ch, 1 y, 1 z, 2 channel groups, 16 blocks}
//               = 5 KB into a ring of stages; out-of-range x / block coordinates are zero-filled.
//   warp 1      MMA issuer (one elected lane): waits for `go`, issues 9 MMAs, commits.
//   warp 2      scout: takes the waits that are not the issuer's own (data landed, accumulators drained) and signals `go`.
//   warps 3..   epilogue groups of four warps (one per TMEM lane quadrant) taking finished output rows in turn:
//               tcgen05.ld -> zero the slot (tcgen05.st) -> release -> +bias -> ReLU -> +residual -> bf16 hi[/lo] -> stores.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {
namespace zy {

constexpr int TX = 8, PX = TX + 2;          // x voxels of the M tile / with halo
constexpr int UNITS = 16;                   // blocks of the M tile (UNITS * TX = 128 rows)
constexpr int COUT = 16, CGI = 2;           // padded channels out / input channel groups of 8
constexpr int UNIT_BYTES = CGI * PX * 16;   // one block's row: 2 channel groups x 10 voxels x 16 B = 320 B  (SBO of A)
constexpr int CG_BYTES = PX * 16;           // 160 B between the two channel groups                            (LBO of A)
constexpr int TERM_BYTES = UNITS * UNIT_BYTES;   // 5120 B per precision term and input row
constexpr int MAX_YO = 10;                  // output rows per tile: 10 * 3 * 16 = 480 TMEM columns
int g_zy_groups = 3;                        // epilogue groups of 4 warps (pccgeo_set_option("zy_groups"): 2 or 3)
constexpr int num_threads(int ng) { return ng * 128 + 96; }   // epilogue groups, TMA producer, scout, MMA issuer
constexpr int MAX_STAGES = 12, MAX_SLOTS = 3 * MAX_YO;
constexpr int BT_BYTES = 2 * 18 * 128;      // one B tile: N = 144 rows x K = 16 bf16 = 4608 B; [kcore 2][18 groups][8 n][8 k]
constexpr int HEADER_BYTES = 1024;

struct Params {
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  const uint8_t* wimg;
  int N, D, H, W;
  int terms, relu, cout_real;
  int ngroups, xsegs, ytiles;
  int xch;                 // x chunks of 8 voxels per block in an M tile: the tile is (16 / xch) blocks x (8 * xch) x-voxels
  int nstage;
  long long term_stride;   // elements between precision terms of y / res
};

struct __align__(8) Header {
  uint64_t in_full[MAX_STAGES], in_empty[MAX_STAGES], go[MAX_STAGES];
  uint64_t acc_full[MAX_SLOTS], acc_empty[MAX_SLOTS];
  uint4 rows[MAX_YO + 2];   // per input row of the tile: {first TMEM column of its window, idesc, first B row group (16 B units),
                            //  first slot of the first-touched output rows | count << 8 | first slot of the completed rows << 16 | count << 24}
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(Header) <= HEADER_BYTES, "header too large");

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// rows [lo, hi] (inclusive, may be empty) of the tile [y0, y0+Yo) that are complete once input row yi (last input row yi1) is in
__device__ __forceinline__ void rows_done(int yi, int yi1, int y0, int Yo, int& lo, int& hi) {
  lo = yi - 1 > y0 ? yi - 1 : y0;
  hi = (yi == yi1 && yi <= y0 + Yo - 1) ? yi : yi - 1;
}
// one mbarrier probe (no spin): true when the phase with this parity has completed
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// tcgen05.mma / tcgen05.commit predicated on `lead` (1 in exactly one lane): no branch around the instruction, so the issue loop
// has no divergence / reconvergence points
__device__ __forceinline__ void umma_bf16_lh_if(uint32_t lead, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 q, %6, 0;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(lead)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t lead, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar),
      "r"(lead)
      : "memory");
}
// the rare commits of the issue loop (last plane: three planes complete per row; bottom tile: two rows complete at once)
__device__ __noinline__ void commit_many(uint32_t lead, uint32_t full_base, uint32_t dfirst, uint32_t dcnt, uint32_t g0, uint32_t g1) {
  for (uint32_t g = g0; g <= g1; ++g)
    for (uint32_t k = 0; k < dcnt; ++k) umma_commit_if(lead, full_base + ((dfirst + k) * 3 + g % 3u) * 8);
}

template <int TERMS, int NGROUPS>
__global__ void __launch_bounds__(NGROUPS * 128 + 96, 1) conv3d_umma_zy_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NUM_THREADS = NGROUPS * 128 + 96;
  constexpr int W_TMA = 4 * NGROUPS, W_SCOUT = W_TMA + 1, W_MMA = W_TMA + 2;   // the issuer gets the highest warp id: the warp
                                                                               // schedulers favour it over the epilogue warps
  constexpr int WBYTES = 3 * 3 * TERMS * BT_BYTES;
  constexpr int STAGE_BYTES = TERMS * TERM_BYTES;
  uint8_t* wsm = smem + HEADER_BYTES;
  uint8_t* stages = wsm + WBYTES;
  constexpr uint32_t TMEM_COLS = 512;

  // ---- the work item of this CTA ----
  const int item = blockIdx.x;
  const int yt = item % p.ytiles, xs = (item / p.ytiles) % p.xsegs, ng = item / (p.ytiles * p.xsegs);
  const int y0 = (int)((long long)yt * p.H / p.ytiles), y1 = (int)((long long)(yt + 1) * p.H / p.ytiles);   // output rows [y0, y1)
  const int Yo = y1 - y0;
  const int yi0 = y0 > 0 ? y0 - 1 : 0, yi1 = y1 < p.H ? y1 : p.H - 1;                                        // input rows [yi0, yi1]
  const int D = p.D;
  const int nslots = 3 * Yo;

  // ---- one-time setup ----
  if (warp == W_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int i = 0; i < p.nstage; ++i) { mbar_init(smem_u32(&hdr->in_full[i]), 1); mbar_init(smem_u32(&hdr->in_empty[i]), 1); mbar_init(smem_u32(&hdr->go[i]), 1); }
    for (int i = 0; i < nslots; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 4); }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < yi1 - yi0 + 1) {
    // the row pattern of the tile is the same for every plane: describe each input row once
    const int yi = yi0 + threadIdx.x;
    const int wlo = yi - 1 > y0 ? yi - 1 : y0, whi = yi + 1 < y1 - 1 ? yi + 1 : y1 - 1;
    // first-touched output rows: the whole window for the tile's first input row, else row yi+1 (if it is in the tile);
    // both these and the completed rows are contiguous ranges: {first slot index, count}
    const int flo = yi == yi0 ? wlo : yi + 1, fhi = whi;
    int rlo, rhi;
    rows_done(yi, yi1, y0, Yo, rlo, rhi);
    const uint32_t fcnt = fhi >= flo ? fhi - flo + 1 : 0, dcnt = rhi >= rlo ? rhi - rlo + 1 : 0;
    hdr->rows[threadIdx.x] = make_uint4((uint32_t)(wlo - y0) * 48u, make_idesc((whi - wlo + 1) * 48), (uint32_t)(wlo - (yi - 1)) * (6 * 128 / 16),
                                        (uint32_t)(flo - y0) * 3u | fcnt << 8 | (uint32_t)((rlo - y0) * 3) << 16 | dcnt << 24);
  }
  for (int i = threadIdx.x * 16; i < WBYTES; i += NUM_THREADS * 16)
    *reinterpret_cast<int4*>(wsm + i) = __ldg(reinterpret_cast<const int4*>(p.wimg + i));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, hdr->tmem_base, 0);

  if (warp == W_TMA) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t s = 0, phase = 0;
      for (int z = 0; z < D; ++z)
        for (int yi = yi0; yi <= yi1; ++yi) {
          mbar_wait(smem_u32(&hdr->in_empty[s]), phase ^ 1);
          const uint32_t full = smem_u32(&hdr->in_full[s]);
          mbar_expect_tx(full, (uint32_t)STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < TERMS; ++t)
            for (int c = 0; c < p.xch; ++c)   // one box per x chunk: (16 / xch) blocks x 10 voxels (its own halo)
              tma_load_5d(smem_u32(stages + (size_t)s * STAGE_BYTES + (size_t)t * TERM_BYTES + (size_t)c * (UNITS / p.xch) * UNIT_BYTES), &tmap_x, full,
                          ((xs * p.xch + c) * TX - 1) * 8, yi, z, 0, t * p.N + ng * (UNITS / p.xch));
          if (++s == (uint32_t)p.nstage) { s = 0; phase ^= 1; }
        }
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer =================
    // Output (pz, y') -- pz in [-1, D] counting the two virtual planes -- lives in slot (y'-y0)*3 + (pz+1)%3 and is that slot's
    // use number (pz+1)/3.  All waits are taken by the whole warp (uniform); one elected lane issues the tcgen05 instructions.
    constexpr int npairs = TERMS == 2 ? 3 : 1;
    const uint64_t adesc = make_smem_desc(0, CG_BYTES, UNIT_BYTES);
    const uint64_t bdesc = make_smem_desc(smem_u32(wsm), 18 * 128, 128);
    const uint32_t a_hi = (uint32_t)(adesc >> 32), a_lo_proto = (uint32_t)adesc;
    const uint32_t b_hi = (uint32_t)(bdesc >> 32), b_lo0 = (uint32_t)bdesc;
    const uint32_t stages16 = smem_u32(stages) / 16;
    // The issuing thread's instruction stream is the critical path (one warp runs ~5 cycles per dependent instruction and the
    // MMA queue holds only a few instructions).  So: the scout warp takes every wait that is not the issuer's own (data landed,
    // accumulators drained) and signals `go`; everything a row needs is computed from block-uniform values (no shared-memory
    // table, no vector-to-uniform register moves); the tcgen05 instructions are predicated on the elected lane instead of
    // branched around; the rare multi-commit cases live in a separate function.
    const int nrows = yi1 - yi0 + 1;
    const uint32_t lead = elect_one() ? 1u : 0u;
    uint32_t s = 0, in_phase = 0;
    uint32_t a_s = a_lo_proto + stages16;
    uint32_t go_bar = smem_u32(&hdr->go[0]), empty_bar = smem_u32(&hdr->in_empty[0]);
    const uint32_t full_base = smem_u32(&hdr->acc_full[0]);
    constexpr uint32_t idesc0 = make_idesc(0);
    uint32_t zs_done8 = 0;   // byte offset of z-slot (z % 3) inside a row's three acc_full barriers: plane z-1 has gp = z
    // the probe of the NEXT row's `go` barrier is issued between the MMAs of the current row: its ~90-cycle latency then
    // overlaps with queued MMAs instead of sitting between two rows (the scout runs ahead, so the probe normally succeeds)
    bool ready = mbar_try(go_bar, in_phase);
    for (int z = 0; z < D; ++z) {
      const uint32_t b_z = b_lo0 + (((uint32_t)(z + 1) % 3u) * 3u) * TERMS * (BT_BYTES / 16);
      const bool fast_z = z != D - 1;
      for (int yi = yi0; yi <= yi1; ++yi) {
        const int wlo = yi - 1 > y0 ? yi - 1 : y0, whi = yi + 1 < y1 - 1 ? yi + 1 : y1 - 1;
        const uint32_t d0 = tmem_base + (uint32_t)(wlo - y0) * 48u;
        const uint32_t idesc = idesc0 | ((uint32_t)(whi - wlo + 1) * 6u) << 17;
        const uint32_t b0 = b_z + (uint32_t)(wlo - yi + 1) * (6 * 128 / 16);
        const int rhi = (yi == yi1 && yi <= y1 - 1) ? yi : yi - 1;    // completed output rows: [wlo, rhi]
        if (!ready) mbar_wait(go_bar, in_phase);
        tc_fence_after();
        const uint32_t a_cur = a_s, empty_cur = empty_bar;
        a_s += STAGE_BYTES / 16; go_bar += 8; empty_bar += 8;
        if (++s == (uint32_t)p.nstage) {
          s = 0; in_phase ^= 1; a_s = a_lo_proto + stages16;
          go_bar = smem_u32(&hdr->go[0]); empty_bar = smem_u32(&hdr->in_empty[0]);
        }
#pragma unroll
        for (int i = 0; i < 3 * npairs; ++i) {
          const int kx = i / npairs, pr = i % npairs;
          const uint32_t ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
          umma_bf16_lh_if(lead, d0, a_cur + ta * (TERM_BYTES / 16) + (uint32_t)kx, a_hi, b0 + ((uint32_t)kx * TERMS + tb) * (BT_BYTES / 16), b_hi, idesc);
          if (i == (3 * npairs) / 2) ready = mbar_try(go_bar, in_phase);   // next row (a failed probe at the very end is harmless)
        }
        umma_commit_if(lead, empty_cur);
        if (fast_z && rhi == wlo) umma_commit_if(lead, full_base + (uint32_t)(wlo - y0) * 24u + zs_done8);
        else if (rhi >= wlo) commit_many(lead, full_base, (uint32_t)(wlo - y0), (uint32_t)(rhi - wlo + 1), (uint32_t)z, fast_z ? (uint32_t)z : (uint32_t)z + 2u);
      }
      zs_done8 = zs_done8 == 16 ? 0 : zs_done8 + 8;
    }
  } else if (warp == W_SCOUT) {
    // ================= scout =================
    // row (z, yi) may be issued once its data has landed and the accumulators it touches for the first time -- plane z+1
    // (planes -1, 0, 1 for z = 0), rows from the row record -- have been drained and zeroed by the epilogue
    if (lane == 0) {
      const int nrows = yi1 - yi0 + 1;
      uint32_t s = 0, phase = 0;
      for (int z = 0; z < D; ++z) {
        const uint32_t gnew0 = z == 0 ? 0u : (uint32_t)z + 2u, gnew1 = (uint32_t)z + 2u;
        for (int j = 0; j < nrows; ++j) {
          const uint32_t rw = hdr->rows[j].w;
          const uint32_t ffirst = rw & 0xffu, fcnt = (rw >> 8) & 0xffu;
          for (uint32_t g = gnew0; g <= gnew1; ++g)
            for (uint32_t k = 0; k < fcnt; ++k) mbar_wait(smem_u32(&hdr->acc_empty[ffirst + 3 * k + g % 3u]), (g / 3u) & 1u);
          mbar_wait(smem_u32(&hdr->in_full[s]), phase);
          mbar_arrive(smem_u32(&hdr->go[s]));
          if (++s == (uint32_t)p.nstage) { s = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue =================
    const int ew = warp;
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access (warp id % 4)
    const int grp = ew >> 2;
    const int row = quad * 32 + lane; // M row = TMEM lane
    // M row -> (block, x): unit u = row / 8 is x chunk u / (16 / xch) of block u % (16 / xch) of the tile
    const int upc = UNITS / p.xch, unit = row >> 3;
    const int n = ng * upc + unit % upc, x = (xs * p.xch + unit / upc) * TX + (row & 7);
    const bool live = n < p.N;
    float bias_r[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias_r[c] = (p.bias && c < p.cout_real) ? __ldg(p.bias + c) : 0.f;
    const long long HW = (long long)p.H * p.W, DHW = HW * D;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    if (grp == 0) {
      // zero every accumulator and publish the slots as empty (completes phase 0 of acc_empty)
      for (int c = 0; c < nslots * 16; c += 16) tmem_st16_zero(lane_base + c);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int i = 0; i < nslots; ++i) mbar_arrive(smem_u32(&hdr->acc_empty[i]));
    }
    // group g drains output rows y0+g, y0+g+NGROUPS, ...: per accumulator slot the uses come in plane order, which is all the
    // hand-over protocol needs; no scan over the other groups' units
    for (int z = 0; z < D; ++z) {
      const int plo = z - 1, phi = z == D - 1 ? D : z - 1;
      for (int yo = y0 + grp; yo < y1; yo += NGROUPS) {
        for (int pz = plo; pz <= phi; ++pz) {
          {
            const bool real = pz >= 0 && pz < D;
            const long long vox = (long long)pz * HW + (long long)yo * p.W + x;
            // the residual operand does not depend on the accumulator: fetch it before waiting
            int4 rq[TERMS * 2];
            if (p.res && real && live) {
#pragma unroll
              for (int t = 0; t < TERMS; ++t)
#pragma unroll
                for (int cg = 0; cg < 2; ++cg)
                  rq[t * 2 + cg] = __ldg(reinterpret_cast<const int4*>(p.res + t * p.term_stride + (((long long)n * 2 + cg) * DHW + vox) * 8));
            }
            const uint32_t gp = (uint32_t)(pz + 1);
            const int slot = (yo - y0) * 3 + (int)(gp % 3u);
            mbar_wait(smem_u32(&hdr->acc_full[slot]), (gp / 3u) & 1u);
            tc_fence_after();
            uint32_t r[COUT];
            tmem_ld16(lane_base + (uint32_t)slot * 16, r);
            tmem_ld_wait();
            tmem_st16_zero(lane_base + (uint32_t)slot * 16);
            tmem_st_wait();   // the slot is handed back zeroed: its next tenant only ever accumulates
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[slot]));
            if (!real || !live) continue;
            float v[COUT];
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
              v[c] = __uint_as_float(r[c]) + bias_r[c];
              if (p.relu) v[c] = fmaxf(v[c], 0.f);
            }
            if (p.res) {
#pragma unroll
              for (int t = 0; t < TERMS; ++t)
#pragma unroll
                for (int cg = 0; cg < 2; ++cg) unpack_bf16x8_add(rq[t * 2 + cg], v + cg * 8);
            }
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {
              const long long e = (((long long)n * 2 + cg) * DHW + vox) * 8;
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
                *reinterpret_cast<int4*>(p.y + p.term_stride + e) = ql;
              }
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

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

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// number of y tiles: all SMs busy in as few waves as possible, at most MAX_YO rows per tile, little halo
static int choose_ytiles(int base_items, int H) {
  int best = -1;
  double best_cost = 0;
  const int lo = (H + MAX_YO - 1) / MAX_YO;
  for (int yt = lo; yt <= H; ++yt) {
    const int waves = (base_items * yt + 147) / 148;
    const double cost = waves * ((double)H / yt + 2.0);
    if (best < 0 || cost < best_cost - 1e-9) { best = yt; best_cost = cost; }
  }
  return best;
}

}  // namespace zy
}  // namespace pccgeo

using namespace pccgeo;

// B image: [rot 3][kx 3][term T][kcore 2][ngroup 18][8 n][8 k] bf16.  n = (dyi*3 + zs)*16 + co: output row y' = yi - 1 + dyi
// (tap ky = 2 - dyi), TMEM z-slot zs holding the plane z + dzs with dzs = ((zs - rot + 1) mod 3) - 1 (tap kz = 1 - dzs), where
// rot = (z + 1) mod 3 is the slot of the input plane's own output plane.
extern "C" long long pccgeo_umma_zy_pack_weights_host(const float* w, void* out, int cin, int cout, int transposed, int terms) {
  if (cin <= 0 || cin > 16 || cout <= 0 || cout > 16 || (terms != 1 && terms != 2)) {
    set_error("umma_zy_pack_weights: <= 16 channels in and out, terms 1 or 2");
    return PCCGEO_EINVAL;
  }
  const long long total = 3LL * 3 * terms * zy::BT_BYTES;
  if (!out) return total;
  if (!w) { set_error("umma_zy_pack_weights: null weights"); return PCCGEO_EINVAL; }
  uint16_t* o = (uint16_t*)out;
  memset(o, 0, (size_t)total);
  for (int rot = 0; rot < 3; ++rot)
    for (int kx = 0; kx < 3; ++kx)
      for (int dyi = 0; dyi < 3; ++dyi)
        for (int zs = 0; zs < 3; ++zs) {
          const int dzs = ((zs - rot + 1) % 3 + 3) % 3 - 1;
          int kz = 1 - dzs, ky = 2 - dyi, kxx = kx;
          if (transposed) { kz = 2 - kz; ky = 2 - ky; kxx = 2 - kxx; }   // stride-1 transposed conv == conv with flipped taps
          for (int co = 0; co < cout; ++co)
            for (int ci = 0; ci < cin; ++ci) {
              const float val = w[((long long)((kz * 3 + ky) * 3 + kxx) * cin + ci) * cout + co];
              const int nrow = (dyi * 3 + zs) * 16 + co, kcore = ci >> 3, ki = ci & 7;
              const uint16_t hi = zy::f2bf(val);
              for (int t = 0; t < terms; ++t) {
                const long long idx = (((((long long)(rot * 3 + kx) * terms + t) * 2 + kcore) * 18 + (nrow >> 3)) * 8 + (nrow & 7)) * 8 + ki;
                o[idx] = t == 0 ? hi : zy::f2bf(val - zy::bf2f(hi));
              }
            }
        }
  return total;
}

extern "C" int pccgeo_conv3d_umma_zy(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb, int n,
                                     int cin, int d, int h, int wd, int cout, int relu, int terms, void* stream) {
  PCCGEO_REQUIRE(xb && wpacked && yb, "conv3d_umma_zy: null pointer");
  PCCGEO_REQUIRE(terms == 1 || terms == 2, "conv3d_umma_zy: terms must be 1 or 2");
  PCCGEO_REQUIRE(cin > 0 && cin <= 16 && cout > 0 && cout <= 16, "conv3d_umma_zy: <= 16 channels in and out (got %d -> %d)", cin, cout);
  PCCGEO_REQUIRE(n > 0 && d > 0 && h > 0 && wd > 0 && wd % zy::TX == 0, "conv3d_umma_zy: W must be a multiple of 8 (got %dx%dx%d)", d, h, wd);
  zy::EncodeTiledFn enc = zy::get_encode_fn();
  PCCGEO_REQUIRE(enc, "conv3d_umma_zy: cuTensorMapEncodeTiled unavailable");
  zy::Params p{};
  p.bias = bias; p.res = (const __nv_bfloat16*)residual_b; p.y = (__nv_bfloat16*)yb; p.wimg = (const uint8_t*)wpacked;
  p.N = n; p.D = d; p.H = h; p.W = wd; p.terms = terms; p.relu = relu; p.cout_real = cout;
  // small batches: fill the 16 units of the M tile with several x chunks of fewer blocks
  p.xch = 1;
  while (p.xch < 4 && n % (zy::UNITS / p.xch) != 0 && wd % (zy::TX * p.xch * 2) == 0) p.xch *= 2;
  p.ngroups = (n + zy::UNITS / p.xch - 1) / (zy::UNITS / p.xch);
  p.xsegs = wd / (zy::TX * p.xch);
  p.ytiles = zy::choose_ytiles(p.ngroups * p.xsegs, h);
  p.term_stride = (long long)n * 16 * d * h * wd;
  const int wbytes = 3 * 3 * terms * zy::BT_BYTES, stage_bytes = terms * zy::TERM_BYTES;
  p.nstage = (227 * 1024 - zy::HEADER_BYTES - wbytes) / stage_bytes;
  if (p.nstage > zy::MAX_STAGES) p.nstage = zy::MAX_STAGES;
  const size_t smem = zy::HEADER_BYTES + wbytes + (size_t)p.nstage * stage_bytes;

  CUtensorMap tmap;
  // blocked layout (term, N, C/8, D, H, W, 8): {x*8ch, y, z, channel group, term*N + block}
  const cuuint64_t gdim[5] = {(cuuint64_t)wd * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)zy::CGI, (cuuint64_t)terms * n};
  const cuuint64_t gstr[4] = {(cuuint64_t)wd * 16, (cuuint64_t)wd * h * 16, (cuuint64_t)wd * h * d * 16, (cuuint64_t)zy::CGI * wd * h * d * 16};
  const cuuint32_t box[5] = {zy::PX * 8, 1, 1, zy::CGI, (cuuint32_t)(zy::UNITS / p.xch)};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xb), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCCGEO_REQUIRE(cr == CUDA_SUCCESS, "conv3d_umma_zy: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const int grid = p.ngroups * p.xsegs * p.ytiles;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kern, int ng) -> int {
    PCCGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<grid, zy::num_threads(ng), smem, st>>>(tmap, p);
    return PCCGEO_OK;
  };
  int lrc;
  if (zy::g_zy_groups == 2) lrc = terms == 2 ? launch(zy::conv3d_umma_zy_kernel<2, 2>, 2) : launch(zy::conv3d_umma_zy_kernel<1, 2>, 2);
  else lrc = terms == 2 ? launch(zy::conv3d_umma_zy_kernel<2, 3>, 3) : launch(zy::conv3d_umma_zy_kernel<1, 3>, 3);
  if (lrc) return lrc;
  return check_launch("conv3d_umma_zy_kernel");
}
