// tcgen05 / TMEM / TMA weight gradient of the 3x3x3 stride-1 'same' convolutions (conv and transposed conv) with
// C = 16 / 32 / 64 channels in and out.
//
// Replaces what `tf.train.AdamOptimizer.minimize` differentiates for the Conv3D / Conv3DTranspose kernels of AnalysisBlock /
// SynthesisBlock and the hyper transforms (reference src/model_types.py:364-369 over src/model_transforms.py:62-81): the
// backward-filter contraction
//     dW[kz,ky,kx,ci,co] = sum_{n,z,y,x} X[n,ci,z+kz-1,y+ky-1,x+kx-1] * G[n,co,z,y,x]
// which is 85 % of the weight-gradient FLOPs of a c3p training step; the stride-2 layers reach this kernel through a phase
// decomposition of their large tensor (training.py::_wgrad_stride2), W = 8 volumes with K chunks that span two rows (rk = 2 below);
// the 4^3 / 2^3 volumes stay on the fp32 kernel of train.cu.
//
// Formulation.  The contraction index is the voxel, so both operands are "MN-major" for the tensor core: in the blocked bf16
// layout (term, N, C/8, D, H, W, 8) a row of voxels is a run of 16-byte items (8 channels each), which read as K = voxel,
// MN = channel is exactly the canonical no-swizzle MN-major core matrix (8 K-rows x 16 bytes).  One MMA (M = 128, K = 16 voxels
// of one x run):
//     A = X rows:  M = (input row j, ci)         -- JR = 128 / C consecutive y rows of ONE input plane z', smem [j][cg][x][8]
//     B = G rows:  N = (plane dz, row r, co)     -- R rows of the THREE gradient planes z'-1, z', z'+1,  smem [slot][r][cg][x][8]
//     D[(j,ci), (dz,r,co)] += sum_x X[ci, z', y0-1+j, x+kx-1] * G[co, z'-1+dz, y0+r, x]
// i.e. the entry belongs to tap kz = 2 - dz, ky = j - r (used when 0 <= ky <= 2; the other (j, r) pairs are the price of the
// band structure), kx = the x shift, which is a 16-byte shift of A's start address inside the halo row.  Both smem strides
// between 8-channel groups are uniform because a row holds its channel groups back to back.  Accumulators (one per kx, and per
// group of JR input rows when R + 2 > JR) stay in TMEM for the whole kernel: the sum over voxels never leaves the tensor core.
//   C = 16: JR = 8, R = 3, N = 144, 3 accumulators (432 columns), 9 of 24 (j, r) pairs used
//   C = 32: JR = 4, R = 1, N =  96, 3 accumulators (288 columns), 3 of 4
//   C = 64: JR = 2, R = 1, N = 192, 2 row groups x 1 kx per CTA (384 columns; the three kx are split over CTAs), 3 of 4
// bf16x3: X and G are hi/lo pairs; three products (hi*hi, hi*lo, lo*hi) per MMA position, fp32 accumulation.
//
// Persistent CTAs (one per SM), work item = (block n, y tile of R rows, z segment).  Per step = one input plane z':
//   warp 4  TMA producer: the X tile (R+2 rows with x halo; out-of-range rows / planes zero-filled = SAME padding) into a ring of
//           stages, and ONE new gradient plane (z'+1) into a ring of 8 plane slots -- the three planes an MMA reads must be at
//           uniform stride in ascending order, so slots 0 and 1 are mirrored behind slot 7 (those planes are loaded twice).
//   warp 5  MMA issuer (one lane): per step  groups x kx x (W/16) x 3 products  MMAs, commit -> stage free.
//   warps 0..3  epilogue, once at the very end: TMEM -> the band entries -> partial[cta][r][tap][ci][co] (every entry written by
//           exactly one thread; a second kernel adds CTAs and r in a fixed order in double: deterministic).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {
namespace wg {

constexpr int S_RING = 8, S_SLOTS = S_RING + 2;   // gradient plane ring + the two mirrored slots
constexpr int MAX_XS = 4;                          // X stages (<= S_RING - 4: a stage's reuse wait also covers the plane slots)
constexpr int HEADER_BYTES = 128;
constexpr int NUM_THREADS = 192;

struct Params {
  float* partial;
  int N, D, H, W;
  int ytiles, zsegs, nitems;
  int xs;          // X stages
  int rk;          // y rows per K chunk (1; 2 for W = 8)
  int workers;     // CTAs per kx class
};

struct __align__(8) Header {
  uint64_t full[MAX_XS], empty[MAX_XS], done;
  uint32_t tmem_base;
};
static_assert(sizeof(Header) <= HEADER_BYTES, "header too large");

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both MN-major (transposed), M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int C, int R, int KXC>
struct Geo {
  static constexpr int CG = C / 8;
  static constexpr int JR = 128 / C;                  // input rows per M = 128
  static constexpr int XROWS = R + 2;                 // input rows a tile needs
  static constexpr int MG = (XROWS + JR - 1) / JR;    // row groups (accumulators per kx)
  static constexpr int N = 3 * R * C;
  static constexpr int NACC = MG * KXC;
  static_assert(N <= 256 && N % 16 == 0, "bad N");
  static_assert(NACC * N <= 512, "accumulators exceed TMEM");
  // rk = y rows per K chunk: 1 (K = 16 voxels of one row), or 2 for W = 8 (K = 8 voxels of row y + 8 of row y+1; R = 1 only): a tile
  // then covers 2 gradient rows and needs 4 input rows
  __host__ __device__ static constexpr int x_rows(int rk) { return R * rk + 2; }
  __host__ __device__ static constexpr int x_term_bytes(int w, int rk) { return ((x_rows(rk) * CG * (w + 2) * 16) + 127) / 128 * 128; }
  __host__ __device__ static constexpr int x_load_bytes(int w, int rk) { return x_rows(rk) * CG * (w + 2) * 16; }
  __host__ __device__ static constexpr int slot_bytes(int w) { return R * CG * w * 16; }   // ONE row set of a plane (per K half when rk = 2)
};

template <int C, int R, int KXC, int TERMS>
__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                    const __grid_constant__ CUtensorMap tmap_g, const Params p) {
  using G = Geo<C, R, KXC>;
  extern __shared__ __align__(1024) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 4, W_MMA = 5;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int KXS = 3 / KXC;   // kx classes of CTAs
  const int W = p.W, PXW = W + 2;
  const int RK = p.rk;
  const int XT = G::x_term_bytes(W, RK), XSTAGE = TERMS * XT;
  const int SLOT = G::slot_bytes(W), GH = S_SLOTS * SLOT, GT = RK * GH;   // gradient ring: [term][K half][slot]
  uint8_t* xst = smem + HEADER_BYTES;
  uint8_t* gring = xst + (size_t)p.xs * XSTAGE;
  const int kxclass = (int)blockIdx.x % KXS, worker = (int)blockIdx.x / KXS;

  if (warp == W_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_g) : "memory");
    for (int i = 0; i < p.xs; ++i) { mbar_init(smem_u32(&hdr->full[i]), 1); mbar_init(smem_u32(&hdr->empty[i]), 1); }
    mbar_init(smem_u32(&hdr->done), 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the M rows beyond the loaded input rows read whatever follows a stage: make sure it is finite once (their lanes are never read,
  // but NaN payloads cost nothing to avoid)
  for (int i = threadIdx.x * 16; i < p.xs * XSTAGE + TERMS * GT; i += NUM_THREADS * 16) *reinterpret_cast<int4*>(xst + i) = make_int4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp == W_TMA) {
    if (lane == 0) {
      uint32_t s = 0, phase = 0, q = 0;
      for (int item = worker; item < p.nitems; item += p.workers) {
        const int zs = item % p.zsegs, col = item / p.zsegs, yt = col % p.ytiles, n = col / p.ytiles;
        const int z0 = (int)((long long)zs * p.D / p.zsegs), z1 = (int)((long long)(zs + 1) * p.D / p.zsegs);
        const int y0 = yt * R * RK;
        for (int z = z0; z < z1; ++z) {
          mbar_wait(smem_u32(&hdr->empty[s]), phase ^ 1);
          const uint32_t full = smem_u32(&hdr->full[s]);
          const int pfirst = z == z0 ? z0 - 1 : z + 1, pcount = z == z0 ? 3 : 1;
          uint32_t copies = 0;
          for (int k = 0; k < pcount; ++k) copies += ((q + k) % S_RING < 2u) ? 2u : 1u;
          mbar_expect_tx(full, (uint32_t)(TERMS * G::x_load_bytes(W, RK)) + copies * (uint32_t)(TERMS * SLOT * RK));
#pragma unroll
          for (int t = 0; t < TERMS; ++t)
            tma_load_5d(smem_u32(xst + (size_t)s * XSTAGE + (size_t)t * XT), &tmap_x, full, 0, -1, (t * p.N + n) * G::CG, y0 - 1, z);
          for (int k = 0; k < pcount; ++k, ++q) {
            const uint32_t slot = q % S_RING;
#pragma unroll
            for (int t = 0; t < TERMS; ++t)
              for (int kk = 0; kk < RK; ++kk) {   // rk = 2: row y0 feeds the first K half, row y0 + 1 the second (boxes of one row each)
                uint8_t* dst = gring + (size_t)t * GT + (size_t)kk * GH;
                tma_load_5d(smem_u32(dst + (size_t)slot * SLOT), &tmap_g, full, 0, 0, (t * p.N + n) * G::CG, y0 + kk, pfirst + k);
                if (slot < 2u)
                  tma_load_5d(smem_u32(dst + (size_t)(slot + S_RING) * SLOT), &tmap_g, full, 0, 0, (t * p.N + n) * G::CG, y0 + kk, pfirst + k);
              }
          }
          if (++s == (uint32_t)p.xs) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc_mn(G::N);
      constexpr int NPROD = TERMS == 2 ? 3 : 1;
      // descriptor words: lo = start address / 16 | LBO (K direction: 8 voxels = 128 B) << 16; hi = SBO (next 8-channel group) | version
      const uint32_t a_hi = (uint32_t)((PXW * 16) >> 4) | (1u << 14), b_hi = (uint32_t)((W * 16) >> 4) | (1u << 14);
      // K direction (LBO): the next 8 voxels of the row (128 B), or with rk = 2 the same 8 voxels of the next row
      const uint32_t lo_proto_a = (uint32_t)((RK == 2 ? G::CG * PXW * 16 : 128) >> 4) << 16;
      const uint32_t lo_proto_b = (uint32_t)((RK == 2 ? GH : 128) >> 4) << 16;
      const uint32_t xst16 = smem_u32(xst) >> 4, g16 = smem_u32(gring) >> 4;
      const int kchunks = RK == 2 ? 1 : W / 16;
      const uint32_t rowgroup16 = (uint32_t)(G::JR * G::CG * PXW);   // JR rows of the X tile, in 16-byte units
      uint32_t s = 0, phase = 0, qw = 0, started = 0;
      for (int item = worker; item < p.nitems; item += p.workers) {
        const int zs = item % p.zsegs;
        const int z0 = (int)((long long)zs * p.D / p.zsegs), z1 = (int)((long long)(zs + 1) * p.D / p.zsegs);
        for (int z = z0; z < z1; ++z, ++qw) {
          mbar_wait(smem_u32(&hdr->full[s]), phase);
          tc_fence_after();
          const uint32_t a0 = xst16 + (uint32_t)(s * XSTAGE >> 4), b0 = g16 + (uint32_t)((qw % S_RING) * SLOT >> 4);
#pragma unroll
          for (int g = 0; g < G::MG; ++g)
#pragma unroll
            for (int kx = 0; kx < KXC; ++kx) {
              const uint32_t d = tmem_base + (uint32_t)((g * KXC + kx) * G::N);
              const uint32_t shift = KXC == 3 ? (uint32_t)kx : (uint32_t)kxclass;
              for (int c = 0; c < kchunks; ++c)
#pragma unroll
                for (int pr = 0; pr < NPROD; ++pr) {
                  const int ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
                  const uint32_t a = a0 + (uint32_t)(ta * XT >> 4) + g * rowgroup16 + (uint32_t)(16 * c) + shift;
                  const uint32_t b = b0 + (uint32_t)(tb * GT >> 4) + (uint32_t)(16 * c);
                  umma_bf16_lh(d, (a & 0x3FFFu) | lo_proto_a, a_hi, (b & 0x3FFFu) | lo_proto_b, b_hi, IDESC, started | (uint32_t)(c | pr));
                }
            }
          umma_commit(smem_u32(&hdr->empty[s]));
          started = 1;
          if (++s == (uint32_t)p.xs) { s = 0; phase ^= 1; }
        }
        qw += 2;
      }
      umma_commit(smem_u32(&hdr->done));
    }
  } else {
    // ================= epilogue (once) =================
    mbar_wait(smem_u32(&hdr->done), 0);
    tc_fence_after();
    const int m = warp * 32 + lane, j = m / C, ci = m % C;
    float* out = p.partial + (size_t)blockIdx.x * (R * 9 * KXC * C * C);
#pragma unroll 1
    for (int g = 0; g < G::MG; ++g)
#pragma unroll 1
      for (int kx = 0; kx < KXC; ++kx)
#pragma unroll 1
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll 1
          for (int r = 0; r < R; ++r) {
            const int ky = g * G::JR + j - r, kz = 2 - dz;
            const bool use = ky >= 0 && ky <= 2;
            float* dst = out + ((((size_t)r * 3 + kz) * 3 + (use ? ky : 0)) * KXC + kx) * (C * C) + (size_t)ci * C;
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((g * KXC + kx) * G::N + (dz * R + r) * C + c0), v);
              tmem_ld_wait();
              if (use) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  *reinterpret_cast<float4*>(dst + c0 + i) =
                      make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
              }
            }
          }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// dw[tap][ci][co] = sum over the CTAs of the tap's kx class and over r; taps mirrored for the transposed conv.  Eight threads share an
// output (interleaved slices of the (cta, r) list, double sums) and are combined by shuffles in a fixed order: deterministic.
__global__ void __launch_bounds__(256) wgrad_umma_finish_kernel(const float* __restrict__ partial, float* __restrict__ dw, int C, int R,
                                                                int KXC, int ctas, int transposed) {
  const int total = 27 * C * C;
  const int KXS = 3 / KXC;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sl = threadIdx.x & 7;
  double s = 0.0;
  int tap = 0, pr = 0;
  if (i < total) {
    tap = i / (C * C); pr = i % (C * C);
    const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
    const int kxl = KXC == 3 ? kx : 0, cls = KXC == 3 ? 0 : kx;
    const int nsrc = ((ctas - cls + KXS - 1) / KXS) * R;
    for (int q = sl; q < nsrc; q += 8) {
      const int cta = cls + (q / R) * KXS, r = q % R;
      s += (double)partial[(size_t)cta * (R * 9 * KXC * C * C) + ((((size_t)r * 3 + kz) * 3 + ky) * KXC + kxl) * (C * C) + pr];
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (i < total && sl == 0) dw[(size_t)(transposed ? 26 - tap : tap) * (C * C) + pr] = (float)s;
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

struct Plan {
  int r, kxc, ctas, workers, ytiles, zsegs, nitems, xs, rk;
  size_t smem, partial_floats;
};

template <int C, int R, int KXC>
static bool make_plan(int n, int d, int h, int w, int terms, Plan& pl) {
  using G = Geo<C, R, KXC>;
  pl.r = R; pl.kxc = KXC;
  pl.rk = w == 8 ? 2 : 1;
  if (pl.rk == 2 && R != 1) return false;
  const int kxs = 3 / KXC;
  pl.ytiles = (h + R * pl.rk - 1) / (R * pl.rk);
  const int cols = n * pl.ytiles;
  const int max_workers = 148 / kxs;
  // z segments: balance the persistent CTAs against the two extra gradient planes every item loads
  int best = 1;
  double best_cost = 0;
  for (int zs = 1; zs <= 8 && zs <= d; zs *= 2) {
    const long long items = (long long)cols * zs;
    const int workers = items < max_workers ? (int)items : max_workers;
    const double cost = (double)((items + workers - 1) / workers) * ((double)d / zs + 2.0);
    if (zs == 1 || cost < best_cost - 1e-9) { best = zs; best_cost = cost; }
  }
  pl.zsegs = best;
  pl.nitems = cols * best;
  pl.workers = pl.nitems < max_workers ? pl.nitems : max_workers;
  pl.ctas = pl.workers * kxs;
  const size_t ring = (size_t)terms * pl.rk * S_SLOTS * G::slot_bytes(w), xstage = (size_t)terms * G::x_term_bytes(w, pl.rk);
  const size_t room = 227 * 1024 - HEADER_BYTES - ring - 1024;   // slack: the unused M rows of the last stage read past it
  if (ring + 2 * xstage + HEADER_BYTES + 1024 > 227 * 1024) return false;
  int xs = (int)(room / xstage);
  pl.xs = xs > MAX_XS ? MAX_XS : xs;
  pl.smem = HEADER_BYTES + ring + (size_t)pl.xs * xstage;
  pl.partial_floats = (size_t)pl.ctas * R * 9 * KXC * C * C;
  return true;
}

static bool plan_for(int c, int n, int d, int h, int w, int terms, Plan& pl) {
  if (n <= 0 || d <= 0 || h <= 0 || (w != 8 && w != 16 && w != 32 && w != 64)) return false;
  if (c == 16) return make_plan<16, 3, 3>(n, d, h, w, terms, pl);
  if (c == 32) return make_plan<32, 1, 3>(n, d, h, w, terms, pl);
  if (c == 64) return make_plan<64, 1, 1>(n, d, h, w, terms, pl);
  return false;
}

template <int C, int R, int KXC>
static int launch(const void* xb, const void* gb, float* dw, float* ws, int n, int d, int h, int w, int transposed, int terms, const Plan& pl,
                  cudaStream_t st) {
  using G = Geo<C, R, KXC>;
  EncodeTiledFn enc = get_encode_fn();
  PCCGEO_REQUIRE(enc, "conv3d_wgrad_umma: cuTensorMapEncodeTiled unavailable");
  // blocked layout (term, N, C/8, D, H, W, 8) seen as {8 ch, x, (term, block, channel group), y, z}: a box lands as [y][cg][x][8]
  const cuuint64_t gdim[5] = {8, (cuuint64_t)w, (cuuint64_t)terms * n * G::CG, (cuuint64_t)h, (cuuint64_t)d};
  const cuuint64_t gstr[4] = {16, (cuuint64_t)d * h * w * 16, (cuuint64_t)w * 16, (cuuint64_t)h * w * 16};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const cuuint32_t box_x[5] = {8, (cuuint32_t)(w + 2), (cuuint32_t)G::CG, (cuuint32_t)G::x_rows(pl.rk), 1};
  const cuuint32_t box_g[5] = {8, (cuuint32_t)w, (cuuint32_t)G::CG, (cuuint32_t)R, 1};   // rk = 2: one box per K half (row)
  CUtensorMap tx, tg;
  CUresult cr = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xb), gdim, gstr, box_x, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCCGEO_REQUIRE(cr == CUDA_SUCCESS, "conv3d_wgrad_umma: cuTensorMapEncodeTiled (x) failed (%d)", (int)cr);
  cr = enc(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(gb), gdim, gstr, box_g, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCCGEO_REQUIRE(cr == CUDA_SUCCESS, "conv3d_wgrad_umma: cuTensorMapEncodeTiled (g) failed (%d)", (int)cr);
  Params p{};
  p.partial = ws; p.N = n; p.D = d; p.H = h; p.W = w;
  p.ytiles = pl.ytiles; p.zsegs = pl.zsegs; p.nitems = pl.nitems; p.xs = pl.xs; p.workers = pl.workers; p.rk = pl.rk;
  auto run = [&](auto kern) -> int {
    PCCGEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<pl.ctas, NUM_THREADS, pl.smem + 1024, st>>>(tx, tg, p);
    return PCCGEO_OK;
  };
  const int lrc = terms == 2 ? run(wgrad_umma_kernel<C, R, KXC, 2>) : run(wgrad_umma_kernel<C, R, KXC, 1>);
  if (lrc) return lrc;
  const int rc = check_launch("wgrad_umma_kernel");
  if (rc) return rc;
  wgrad_umma_finish_kernel<<<(27 * C * C * 8 + 255) / 256, 256, 0, st>>>(ws, dw, C, R, KXC, pl.ctas, transposed);
  return check_launch("wgrad_umma_finish_kernel");
}

}  // namespace wg
}  // namespace pccgeo

using namespace pccgeo;

extern "C" long long pccgeo_wgrad_umma_ws_floats(int c, int n, int d, int h, int wd, int terms) {
  wg::Plan pl;
  if ((terms != 1 && terms != 2) || !wg::plan_for(c, n, d, h, wd, terms, pl)) return 0;   // 0: this geometry is not supported
  return (long long)pl.partial_floats;
}

extern "C" int pccgeo_conv3d_wgrad_umma(const void* xb, const void* gb, float* dw, float* ws, int n, int c, int d, int h, int wd, int transposed,
                                        int terms, void* stream) {
  PCCGEO_REQUIRE(xb && gb && dw && ws, "conv3d_wgrad_umma: null pointer");
  PCCGEO_REQUIRE(terms == 1 || terms == 2, "conv3d_wgrad_umma: terms must be 1 or 2");
  wg::Plan pl;
  PCCGEO_REQUIRE(wg::plan_for(c, n, d, h, wd, terms, pl),
                 "conv3d_wgrad_umma: needs C in {16, 32, 64} (in == out) and W in {16, 32, 64} (8 for C >= 32) (got C=%d, %dx%dx%d)", c, d, h, wd);
  cudaStream_t st = (cudaStream_t)stream;
  if (c == 16) return wg::launch<16, 3, 3>(xb, gb, dw, ws, n, d, h, wd, transposed, terms, pl, st);
  if (c == 32) return wg::launch<32, 1, 3>(xb, gb, dw, ws, n, d, h, wd, transposed, terms, pl, st);
  return wg::launch<64, 1, 1>(xb, gb, dw, ws, n, d, h, wd, transposed, terms, pl, st);
}
