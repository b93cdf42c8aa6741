// General tcgen05 implicit-GEMM 3D convolution: gather-im2col -> UMMA.  Covers what the TMA halo-plane kernel
// (conv3d_umma.cu) does not: stride-2 convs, stride-2 transposed convs, 64-channel layers (weights streamed per tap),
// 5^3 / 9^3 kernels, and volumes of any size -- the batch is folded into the GEMM M dimension, so a 4^3 latent volume
// still fills 128-row tiles.
//
// Replaces Keras Conv3D / Conv3DTranspose 'same' + BiasAdd + Relu + ResidualLayer add
// (reference src/model_transforms.py:45-47,56-58,67-69,78-80,93,107,121,135,144-146,155-157).
//
// Formulation (same as conv3d_direct.cu): outputs are split into sub-pixel classes (1, or 8 parity classes for a
// stride-2 transposed conv); within a class, output voxel o (written at o*s_out + P) gathers input voxel o*s_in + off_t
// for each tap t of the class.  Two tilings:
//   NACC = 1  "class mode": a GEMM tile is 128 consecutive (n, o) rows of ONE class; K runs over that class's taps.
//   NACC = 8  "shift mode" (stride-2 transposed convs): a tile is 128 (n, o) rows and carries all 8 class accumulators
//             in TMEM.  The A tile of one input offset vector is gathered ONCE and multiplied with the weight chunk of
//             every class that uses that offset (3^3 kernel: 8 gathers feed 27 MMAs instead of 27 gathers).
//
// Pipeline per CTA (persistent over tiles):
//   warps 0-3 (0-7 for wide layers)  gather producers: thread r owns tile row r (wide layers: two threads per row, each half
//              of the (term, channel group) vectors -- a lone producer warp per scheduler runs ~5 cycles per dependent
//              instruction and was the bound: ncu showed the MMA warp waiting for `full` 83 % of the time); per tap it loads the row's Cin channels from the blocked
//              bf16 layout (16-byte LDG per channel group and precision term, zeros when out of bounds) and stores
//              them into the UMMA no-swizzle K-major layout of a ring stage (conflict-free 16-byte STS);
//              fence.proxy.async + mbarrier arrive.  The loads of the next (tile, tap) pair are already in flight
//              (software pipeline across tile boundaries).  Thread 0 streams the tap's weight chunk(s) with cp.async.bulk.
//   warp 4     MMA issuer: per tap and class, KC x {1|3} tcgen05.mma (M=128, N=Cout) into TMEM accumulators.
//   warps 5-8  epilogue: tcgen05.ld -> +bias -> ReLU -> +residual -> bf16 hi[/lo] -> 16-byte global stores, overlapped
//              with the next tile's main loop when a second accumulator set fits in TMEM.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace pccgeo {

constexpr int GM = 128;
__host__ __device__ constexpr int g_threads(int pg) { return (4 * pg + 5) * 32; }   // pg producer groups of 4 warps, MMA warp, 4 epilogue warps
constexpr int G_MAX_STAGES = 8;
constexpr int G_HEADER_BYTES = 1024;

struct GemmConvParams {
  const __nv_bfloat16* x;
  const uint8_t* wchunks;  // weight chunk groups: NACC chunk slots per tap, each chunk holds all precision terms
  const int4* taps;        // (dz, dy, dx, group index | class mask << 16) per tap, classes concatenated
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  int N, Din, Hin, Win, Dout, Hout, Wout, Dc, Hc, Wc;
  int CGi, CGo, cout_real;
  int s_in, s_out, ncls, relu;
  int cls_tap_begin[9];
  long long rows_per_cls;
  int tiles_per_cls, total_tiles;
  long long term_stride_in, term_stride_out;
  int nstage, wchunk_bytes, a_stage_bytes, stage_bytes;
};

struct __align__(8) GemmSmemHeader {
  uint64_t full[G_MAX_STAGES], empty[G_MAX_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <int COUT, int NACC>
struct GemmTmem {
  static constexpr int kBufs = (2 * NACC * COUT <= 512) ? 2 : 1;
  static constexpr uint32_t kCols = (kBufs * NACC * COUT < 32) ? 32 : kBufs * NACC * COUT;
};

// small class-mode configurations (<= 8 gather vectors per tap, <= 32 output channels) fit twice on an SM
// PG = 2 producer groups (two threads per tile row) for the configurations that run one CTA per SM
template <int COUT, int TERMS, int KC, int NACC>
__host__ __device__ constexpr int gemm_ctas_per_sm() { return (TERMS * 2 * KC <= 8 && COUT <= 32 && GemmTmem<COUT, NACC>::kCols <= 256) ? 2 : 1; }
template <int COUT, int TERMS, int KC, int NACC>
__host__ __device__ constexpr int gemm_pg() { return (gemm_ctas_per_sm<COUT, TERMS, KC, NACC>() == 1 && TERMS * 2 * KC >= 4) ? 2 : 1; }

template <int COUT, int TERMS, int KC, int NACC>
__global__ void __launch_bounds__(g_threads(gemm_pg<COUT, TERMS, KC, NACC>()), gemm_ctas_per_sm<COUT, TERMS, KC, NACC>())
conv3d_gemm_kernel(const GemmConvParams p) {
  constexpr int PG = gemm_pg<COUT, TERMS, KC, NACC>();
  constexpr int W_MMA = 4 * PG + 4;   // producers 0 .. 4*PG-1, epilogue 4*PG .. 4*PG+3, the MMA issuer last: the warp schedulers
                                      // favour high warp ids and the issuer's instruction stream is on the critical path
  extern __shared__ __align__(1024) uint8_t smem[];
  GemmSmemHeader* hdr = reinterpret_cast<GemmSmemHeader*>(smem);
  uint8_t* stages = smem + G_HEADER_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CGI = 2 * KC;
  constexpr int NBUF = GemmTmem<COUT, NACC>::kBufs;
  constexpr uint32_t tmem_cols = GemmTmem<COUT, NACC>::kCols;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstage; ++i) { mbar_init(smem_u32(&hdr->full[i]), GM * PG); mbar_init(smem_u32(&hdr->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 4); }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, hdr->tmem_base, 0);

  if (warp < 4 * PG) {
    // ================= gather producers =================
    const int r = threadIdx.x & (GM - 1), half = threadIdx.x >> 7;   // half: which of the PG interleaved vector subsets
    constexpr int NV = TERMS * CGI / PG;                              // vectors (16 B) per thread and tap
    const long long HWin = (long long)p.Hin * p.Win, DHWin = HWin * p.Din;
    const uint32_t row_off = (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
    uint32_t s = 0, ph = 0;
    struct Cursor {
      int tile, ti, t_end, iz0, iy0, ix0;
      bool valid;
      const __nv_bfloat16* xn;
    };
    auto open_tile = [&](Cursor& c) {
      if (c.tile >= p.total_tiles) return;
      const int cls = c.tile / p.tiles_per_cls, t = c.tile - cls * p.tiles_per_cls;
      uint32_t L = (uint32_t)t * GM + r;  // rows_per_cls < 2^31 (host-checked): 32-bit decode
      c.valid = L < (uint32_t)p.rows_per_cls;
      if (!c.valid) L = 0;
      const int ox = (int)(L % (uint32_t)p.Wc); L /= (uint32_t)p.Wc;
      const int oy = (int)(L % (uint32_t)p.Hc); L /= (uint32_t)p.Hc;
      const int oz = (int)(L % (uint32_t)p.Dc);
      const int n = (int)(L / (uint32_t)p.Dc);
      c.iz0 = oz * p.s_in; c.iy0 = oy * p.s_in; c.ix0 = ox * p.s_in;
      c.xn = p.x + (long long)n * CGI * DHWin * 8;
      c.ti = p.cls_tap_begin[cls];
      c.t_end = p.cls_tap_begin[cls + 1];
    };
    auto advance = [&](Cursor& c) {
      if (++c.ti == c.t_end) { c.tile += gridDim.x; open_tile(c); }
    };
    auto issue_loads = [&](const Cursor& c, const int4& tap, int4* dst) {
      const int iz = c.iz0 + tap.x, iy = c.iy0 + tap.y, ix = c.ix0 + tap.z;
      const bool inb = c.valid && iz >= 0 && iz < p.Din && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
      const __nv_bfloat16* src = c.xn + (iz * HWin + (long long)iy * p.Win + ix) * 8;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int i = j * PG + half, tt = i / CGI, cg = i % CGI;   // vector i = (term, channel group)
        dst[j] = inb ? __ldg(reinterpret_cast<const int4*>(src + tt * p.term_stride_in + (long long)cg * DHWin * 8)) : make_int4(0, 0, 0, 0);
      }
    };
    Cursor cur;
    cur.tile = blockIdx.x;
    cur.ti = cur.t_end = cur.iz0 = cur.iy0 = cur.ix0 = 0;
    cur.valid = false;
    cur.xn = p.x;
    open_tile(cur);
    int4 v[NV], vn[NV];
    int4 tp = make_int4(0, 0, 0, 0);
    if (cur.tile < p.total_tiles) {
      tp = __ldg(p.taps + cur.ti);
      issue_loads(cur, tp, v);
    }
    while (cur.tile < p.total_tiles) {
      Cursor nxt = cur;
      advance(nxt);
      int4 tpn = tp;
      if (nxt.tile < p.total_tiles) {
        tpn = __ldg(p.taps + nxt.ti);
        issue_loads(nxt, tpn, vn);
      }
      mbar_wait(smem_u32(&hdr->empty[s]), ph ^ 1);
      uint8_t* st = stages + (size_t)s * p.stage_bytes;
      if (threadIdx.x == 0) {
        const uint32_t full = smem_u32(&hdr->full[s]);
        const uint32_t group = (uint32_t)tp.w & 0xffffu;
        const uint32_t bytes = (NACC == 1 ? 1u : (uint32_t)__popc((uint32_t)tp.w >> 16)) * (uint32_t)p.wchunk_bytes;
        mbar_expect_tx_only(full, bytes);
        bulk_g2s(smem_u32(st + p.a_stage_bytes), p.wchunks + (size_t)group * NACC * p.wchunk_bytes, bytes, full);
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) *reinterpret_cast<int4*>(st + (j * PG + half) * 2048 + row_off) = v[j];
      fence_proxy_async();
      mbar_arrive(smem_u32(&hdr->full[s]));
      if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; }
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = vn[j];
      tp = tpn;
      cur = nxt;
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer =================
    constexpr uint32_t b_kc16 = 2 * (COUT / 8) * 128 / 16;     // one k-chunk of B (two K core matrices), 16-byte units
    constexpr uint32_t b_term16 = KC * b_kc16;
    constexpr int npairs = TERMS == 2 ? 3 : 1;
    const uint64_t a_proto = make_smem_desc(0, 2048, 128);
    const uint64_t b_proto = make_smem_desc(0, (COUT / 8) * 128, 128);
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
    const uint32_t stages16 = smem_u32(stages) / 16, stage16 = p.stage_bytes / 16, aw16 = p.a_stage_bytes / 16;
    const uint32_t wchunk16 = p.wchunk_bytes / 16;
    constexpr uint32_t idesc = make_idesc(COUT);
    uint32_t s = 0, ph = 0, u = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++u) {
      const int cls = tile / p.tiles_per_cls;
      const int t0 = p.cls_tap_begin[cls], t1 = p.cls_tap_begin[cls + 1];
      const uint32_t buf = u % NBUF;
      mbar_wait(smem_u32(&hdr->acc_empty[buf]), (((u / NBUF) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * NACC * COUT;
      uint32_t touched = 0;  // accumulators that already hold a partial sum in this tile
      for (int ti = t0; ti < t1; ++ti) {
        const uint32_t mask = NACC == 1 ? 1u : ((uint32_t)__ldg(&p.taps[ti].w) >> 16);
        mbar_wait(smem_u32(&hdr->full[s]), ph);
        tc_fence_after();
        const uint32_t a_lo0 = (uint32_t)a_proto + stages16 + s * stage16;
        uint32_t b_lo0 = (uint32_t)b_proto + stages16 + s * stage16 + aw16;
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < NACC; ++c) {
            if (!((mask >> c) & 1)) continue;
            const uint32_t fresh = ((touched >> c) & 1) ^ 1;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc)
#pragma unroll
              for (int pr = 0; pr < npairs; ++pr) {
                const uint32_t ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
                umma_bf16_lh(d0 + c * COUT, a_lo0 + (ta * CGI + 2 * kc) * (2048 / 16), a_hi, b_lo0 + tb * b_term16 + kc * b_kc16, b_hi,
                             idesc, (fresh && kc == 0 && pr == 0) ? 0u : 1u);
              }
            b_lo0 += wchunk16;  // next chunk slot of this tap's group
          }
          umma_commit(smem_u32(&hdr->empty[s]));
          if (ti == t1 - 1) umma_commit(smem_u32(&hdr->acc_full[buf]));
        }
        touched |= mask;
        __syncwarp();
        if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float bias_r[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias_r[c] = (p.bias && c < p.cout_real) ? __ldg(p.bias + c) : 0.f;
    const long long HWo = (long long)p.Hout * p.Wout, DHWo = HWo * p.Dout;
    uint32_t u = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++u) {
      const int tcls = tile / p.tiles_per_cls, t = tile - tcls * p.tiles_per_cls;
      const uint32_t buf = u % NBUF;
      mbar_wait(smem_u32(&hdr->acc_full[buf]), (u / NBUF) & 1);
      tc_fence_after();
      uint32_t L = (uint32_t)t * GM + row;
      const bool valid = L < (uint32_t)p.rows_per_cls;
      if (!valid) L = 0;
      const int ox = (int)(L % (uint32_t)p.Wc); L /= (uint32_t)p.Wc;
      const int oy = (int)(L % (uint32_t)p.Hc); L /= (uint32_t)p.Hc;
      const int oz = (int)(L % (uint32_t)p.Dc);
      const int n = (int)(L / (uint32_t)p.Dc);
#pragma unroll 1
      for (int c = 0; c < NACC; ++c) {
        uint32_t rg[COUT];
#pragma unroll
        for (int k = 0; k < COUT; k += 16) tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (buf * NACC + c) * COUT + k, rg + k);
        tmem_ld_wait();
        if (c == NACC - 1) {  // every accumulator of this buffer is in registers: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[buf]));
        }
        if (!valid) continue;
        const int cls = NACC == 1 ? tcls : c;
        const int Pz = (cls >> 2) & 1, Py = (cls >> 1) & 1, Px = cls & 1;
        const long long vox = (long long)(oz * p.s_out + Pz) * HWo + (long long)(oy * p.s_out + Py) * p.Wout + (ox * p.s_out + Px);
        float v[COUT];
#pragma unroll
        for (int k = 0; k < COUT; ++k) {
          v[k] = __uint_as_float(rg[k]) + bias_r[k];
          if (p.relu) v[k] = fmaxf(v[k], 0.f);
        }
        if (p.res) {
#pragma unroll
          for (int tt = 0; tt < TERMS; ++tt)
#pragma unroll
            for (int cg = 0; cg < COUT / 8; ++cg) {
              const long long e = tt * p.term_stride_out + (((long long)n * p.CGo + cg) * DHWo + vox) * 8;
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

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------
// host side: tap tables + weight image
// --------------------------------------------------------------------------------------------------------
struct TapTable {
  int ncls = 1, s_in = 1, s_out = 1, nacc = 1;
  int begin[9] = {0};
  std::vector<int4> taps;                 // (dz,dy,dx, group | mask << 16)
  std::vector<std::vector<int>> kernels;  // per tap: kernel tap index (kz*k+ky)*k+kx of each chunk slot
};

static inline int round_up_i(int a, int m) { return (a + m - 1) / m * m; }

// Even spatial dims are assumed for stride 2 (TF SAME pad_before then does not depend on the size).
static bool build_taps(int k, int stride, int transposed, bool shift_mode, TapTable& T) {
  if (k < 1 || k > 9 || !(k & 1) || (stride != 1 && stride != 2)) return false;
  int nt[2] = {0, 0}, tk[2][9], toff[2][9];
  const int pb = stride == 1 ? (k - 1) / 2 : (k - 2) / 2;  // SAME, even sizes: total = k - stride, before = total / 2
  if (!transposed) {
    T.ncls = 1; T.s_in = stride; T.s_out = 1;
    nt[0] = k;
    for (int j = 0; j < k; ++j) { tk[0][j] = j; toff[0][j] = j - pb; }
  } else {
    T.ncls = stride == 1 ? 1 : 8; T.s_in = 1; T.s_out = stride;
    for (int P = 0; P < stride; ++P) {
      int cnt = 0;
      for (int j = 0; j < k; ++j) {
        const int num = P + pb - j;  // output p = i*s + j - pb  =>  i = (p + pb - j)/s
        if (((num % stride) + stride) % stride != 0) continue;
        tk[P][cnt] = j; toff[P][cnt] = num / stride; ++cnt;
      }
      nt[P] = cnt;
    }
  }
  T.taps.clear();
  T.kernels.clear();
  if (!(shift_mode && transposed && stride == 2)) {
    T.nacc = 1;
    for (int cls = 0; cls < T.ncls; ++cls) {
      const int Pz = (cls >> 2) & 1, Py = (cls >> 1) & 1, Px = cls & 1;
      T.begin[cls] = (int)T.taps.size();
      for (int jz = 0; jz < nt[Pz]; ++jz)
        for (int jy = 0; jy < nt[Py]; ++jy)
          for (int jx = 0; jx < nt[Px]; ++jx) {
            const int g = (int)T.taps.size();
            T.taps.push_back(make_int4(toff[Pz][jz], toff[Py][jy], toff[Px][jx], g | (1 << 16)));
            T.kernels.push_back({(tk[Pz][jz] * k + tk[Py][jy]) * k + tk[Px][jx]});
          }
    }
    T.begin[T.ncls] = (int)T.taps.size();
    return true;
  }
  // shift mode: one tap per distinct input offset vector; mask = classes that have a kernel tap at that offset
  T.nacc = 8;
  int omin = 0, omax = 0;
  for (int P = 0; P < 2; ++P)
    for (int j = 0; j < nt[P]; ++j) { omin = toff[P][j] < omin ? toff[P][j] : omin; omax = toff[P][j] > omax ? toff[P][j] : omax; }
  auto tap_of = [&](int P, int off) {  // kernel index of parity P at input offset off, or -1
    for (int j = 0; j < nt[P]; ++j)
      if (toff[P][j] == off) return tk[P][j];
    return -1;
  };
  for (int oz = omin; oz <= omax; ++oz)
    for (int oy = omin; oy <= omax; ++oy)
      for (int ox = omin; ox <= omax; ++ox) {
        int mask = 0;
        std::vector<int> ks;
        for (int cls = 0; cls < 8; ++cls) {
          const int kz = tap_of((cls >> 2) & 1, oz), ky = tap_of((cls >> 1) & 1, oy), kx = tap_of(cls & 1, ox);
          if (kz < 0 || ky < 0 || kx < 0) continue;
          mask |= 1 << cls;
          ks.push_back((kz * k + ky) * k + kx);
        }
        if (!mask) continue;
        const int g = (int)T.taps.size();
        T.taps.push_back(make_int4(oz, oy, ox, g | (mask << 16)));
        T.kernels.push_back(ks);
      }
  T.ncls = 8;
  T.begin[0] = 0;
  for (int i = 1; i < 9; ++i) T.begin[i] = (int)T.taps.size();
  return true;
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

// image = [int32 header (128 B)] [int4 taps] [pad to 128] [chunk groups: ntaps x nacc slots x
//          (terms x [kc][kcore][Cout_p/8][8 n][8 k] bf16)]
constexpr int kImgMagic = 0x47454d32;  // "GEM2"
constexpr int kImgHeaderBytes = 128;

template <int COUT, int TERMS, int KC, int NACC>
static int launch_gemm(const GemmConvParams& p, size_t smem, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    PCCGEO_CUDA(cudaFuncSetAttribute(conv3d_gemm_kernel<COUT, TERMS, KC, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  conv3d_gemm_kernel<COUT, TERMS, KC, NACC><<<grid, g_threads(gemm_pg<COUT, TERMS, KC, NACC>()), smem, st>>>(p);
  return check_launch("conv3d_gemm_kernel");
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" long long pccgeo_gemm_pack_weights_host(const float* w, void* out, int cin, int cout, int k, int stride,
                                                   int transposed, int terms) {
  if (cin <= 0 || cout <= 0 || (terms != 1 && terms != 2)) {
    set_error("gemm_pack_weights: bad argument");
    return PCCGEO_EINVAL;
  }
  const int cip = round_up_i(cin, 16), cop = round_up_i(cout, 16), KC = cip / 16;
  const long long per_term = (long long)KC * 2 * (cop / 8) * 128;
  const long long wchunk = per_term * terms;
  // shift mode needs (A tile + 8 chunk slots) twice in shared memory
  const long long a_stage = (long long)terms * (cip / 8) * 2048;
  const bool shift_mode = transposed && stride == 2 && 2 * (a_stage + 8 * wchunk) + G_HEADER_BYTES <= 227 * 1024;
  TapTable T;
  if (!build_taps(k, stride, transposed, shift_mode, T)) {
    set_error("gemm_pack_weights: unsupported kernel %d / stride %d", k, stride);
    return PCCGEO_EINVAL;
  }
  const int ntaps = (int)T.taps.size();
  const long long taps_off = kImgHeaderBytes;
  const long long chunks_off = (taps_off + (long long)ntaps * 16 + 127) / 128 * 128;
  const long long total = chunks_off + wchunk * T.nacc * ntaps;
  if (!out) return total;
  if (!w) { set_error("gemm_pack_weights: null weights"); return PCCGEO_EINVAL; }
  uint8_t* img = (uint8_t*)out;
  memset(img, 0, (size_t)total);
  int32_t* h = (int32_t*)img;
  h[0] = kImgMagic; h[1] = ntaps; h[2] = T.ncls;
  for (int i = 0; i < 9; ++i) h[3 + i] = T.begin[i];
  h[12] = (int32_t)wchunk; h[13] = (int32_t)taps_off; h[14] = (int32_t)chunks_off;
  h[15] = cin; h[16] = cout; h[17] = k; h[18] = stride; h[19] = transposed; h[20] = terms; h[21] = T.s_in; h[22] = T.s_out;
  h[23] = T.nacc;
  int4* taps = (int4*)(img + taps_off);
  for (int i = 0; i < ntaps; ++i) {
    taps[i] = T.taps[i];
    for (size_t slot = 0; slot < T.kernels[i].size(); ++slot) {
      const int kidx = T.kernels[i][slot];
      uint16_t* o = (uint16_t*)(img + chunks_off + wchunk * ((long long)i * T.nacc + (long long)slot));
      // runs once per layer and training step: walk the source contiguously
      for (int ci = 0; ci < cin; ++ci) {
        const float* wr = w + ((long long)kidx * cin + ci) * cout;
        const int kc = ci / 16, kk = ci % 16, kcore = kk >> 3, ki = kk & 7;
        uint16_t* orow = o + ((long long)kc * 2 + kcore) * (cop / 8) * 64 + ki;
        uint16_t* lrow = orow + per_term / 2;
        for (int co = 0; co < cout; ++co) {
          const float val = wr[co];
          const uint16_t hi = f2bf(val);
          orow[co * 8] = hi;
          if (terms == 2) lrow[co * 8] = f2bf(val - bf2f(hi));
        }
      }
    }
  }
  return total;
}

extern "C" int pccgeo_conv3d_gemm(const void* xb, const void* wimg_dev, const void* wimg_header_host, const float* bias,
                                  const void* residual_b, void* yb, int n, int cin, int d, int h, int wd, int cout, int relu,
                                  void* stream) {
  PCCGEO_REQUIRE(xb && wimg_dev && wimg_header_host && yb, "conv3d_gemm: null pointer");
  const int32_t* hh = (const int32_t*)wimg_header_host;
  PCCGEO_REQUIRE(hh[0] == kImgMagic, "conv3d_gemm: bad weight image");
  PCCGEO_REQUIRE(hh[15] == cin && hh[16] == cout, "conv3d_gemm: weight image is %dx%d, layer is %dx%d", hh[15], hh[16], cin, cout);
  const int stride = hh[18], transposed = hh[19], terms = hh[20], nacc = hh[23];
  PCCGEO_REQUIRE(n > 0 && d > 0 && h > 0 && wd > 0, "conv3d_gemm: bad shape");
  PCCGEO_REQUIRE(stride == 1 || transposed || (d % 2 == 0 && h % 2 == 0 && wd % 2 == 0), "conv3d_gemm: stride-2 conv needs even dims");
  const int cip = round_up_i(cin, 16), cop = round_up_i(cout, 16), KC = cip / 16;
  PCCGEO_REQUIRE((cop == 16 || cop == 32 || cop == 64) && (KC == 1 || KC == 2 || KC == 4), "conv3d_gemm: channels %d -> %d unsupported", cin, cout);
  PCCGEO_REQUIRE(nacc == 1 || nacc == 8, "conv3d_gemm: bad weight image (nacc)");
  GemmConvParams p{};
  p.x = (const __nv_bfloat16*)xb;
  p.taps = (const int4*)((const uint8_t*)wimg_dev + hh[13]);
  p.wchunks = (const uint8_t*)wimg_dev + hh[14];
  p.bias = bias; p.res = (const __nv_bfloat16*)residual_b; p.y = (__nv_bfloat16*)yb;
  p.N = n; p.Din = d; p.Hin = h; p.Win = wd;
  p.s_in = hh[21]; p.s_out = hh[22]; p.ncls = hh[2];
  if (transposed) { p.Dout = d * stride; p.Hout = h * stride; p.Wout = wd * stride; }
  else { p.Dout = (d + stride - 1) / stride; p.Hout = (h + stride - 1) / stride; p.Wout = (wd + stride - 1) / stride; }
  p.Dc = p.Dout / p.s_out; p.Hc = p.Hout / p.s_out; p.Wc = p.Wout / p.s_out;
  p.CGi = cip / 8; p.CGo = cop / 8; p.cout_real = cout; p.relu = relu;
  for (int i = 0; i < 9; ++i) p.cls_tap_begin[i] = hh[3 + i];
  p.rows_per_cls = (long long)n * p.Dc * p.Hc * p.Wc;
  PCCGEO_REQUIRE(p.rows_per_cls < (1LL << 31) - GM, "conv3d_gemm: too many output voxels per launch (%lld)", p.rows_per_cls);
  p.tiles_per_cls = (int)((p.rows_per_cls + GM - 1) / GM);
  p.total_tiles = p.tiles_per_cls * (nacc == 8 ? 1 : p.ncls);
  p.term_stride_in = (long long)n * cip * d * h * wd;
  p.term_stride_out = (long long)n * cop * p.Dout * p.Hout * p.Wout;
  p.wchunk_bytes = hh[12];
  p.a_stage_bytes = terms * p.CGi * 2048;
  p.stage_bytes = (p.a_stage_bytes + nacc * p.wchunk_bytes + 127) & ~127;
  const int tmem_cols = (2 * nacc * cop <= 512 ? 2 : 1) * nacc * cop;
  const int ctas_per_sm = (terms * 2 * KC <= 8 && cop <= 32 && tmem_cols <= 256) ? 2 : 1;
  const int avail = (ctas_per_sm == 2 ? 110 : 227) * 1024 - G_HEADER_BYTES;
  p.nstage = avail / p.stage_bytes;
  if (p.nstage > G_MAX_STAGES) p.nstage = G_MAX_STAGES;
  PCCGEO_REQUIRE(p.nstage >= 2, "conv3d_gemm: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
  const size_t smem = G_HEADER_BYTES + (size_t)p.nstage * p.stage_bytes;
  int grid = p.total_tiles < 148 * ctas_per_sm ? p.total_tiles : 148 * ctas_per_sm;
  cudaStream_t st = (cudaStream_t)stream;
#define PCCGEO_DISPATCH(CO, T, K) \
  if (cop == CO && terms == T && KC == K) return nacc == 8 ? launch_gemm<CO, T, K, 8>(p, smem, grid, st) : launch_gemm<CO, T, K, 1>(p, smem, grid, st);
  PCCGEO_DISPATCH(16, 1, 1) PCCGEO_DISPATCH(16, 1, 2) PCCGEO_DISPATCH(16, 1, 4)
  PCCGEO_DISPATCH(32, 1, 1) PCCGEO_DISPATCH(32, 1, 2) PCCGEO_DISPATCH(32, 1, 4)
  PCCGEO_DISPATCH(64, 1, 1) PCCGEO_DISPATCH(64, 1, 2) PCCGEO_DISPATCH(64, 1, 4)
  PCCGEO_DISPATCH(16, 2, 1) PCCGEO_DISPATCH(16, 2, 2) PCCGEO_DISPATCH(16, 2, 4)
  PCCGEO_DISPATCH(32, 2, 1) PCCGEO_DISPATCH(32, 2, 2) PCCGEO_DISPATCH(32, 2, 4)
  PCCGEO_DISPATCH(64, 2, 1) PCCGEO_DISPATCH(64, 2, 2) PCCGEO_DISPATCH(64, 2, 4)
#undef PCCGEO_DISPATCH
  set_error("conv3d_gemm: unsupported configuration");
  return PCCGEO_EINVAL;
}
