// Range coder on the GPU, byte-identical to the host coder (range_coder.cpp): one independent stream
// per block and latent, same symbol semantics (tfc's unbounded-index contract: value = symbol - offset[row], escape symbol +
// Elias-gamma-like overflow code in 4-bit chunks) and the same 32-bit range coder (carry propagation, terminator, stripped
// zero bytes).  Replaces the host half of EntropyBottleneck.compress/.decompress and GaussianConditional.compress/.decompress
// (reference src/model_types.py:383-387,404-408; src/utils/patch_gaussian_conditional.py:27-31) inside the block loops: with
// several GPUs per host the host coder is what bounds the end-to-end throughput.
//
// A stream is inherently serial (every symbol's interval depends on the previous one), so the unit of parallelism is the
// stream: ONE WARP per stream, and the lanes are used for everything around the serial chain:
//   encode: rc_map_kernel (fully parallel) first turns symbols into (cdf_lo, freq-1, overflow) so that the chain never
//           touches a table; rc_encode_kernel's lanes fetch 32 mapped symbols at a time (coalesced) and broadcast them one
//           by one; lane 0 emits the bytes.
//   decode: the lanes prefetch the stream bytes (128 at a time, double buffered) and the next 32 table indexes, and the
//           CDF search is lane-parallel and division-free (see rc_decode_kernel), on compact tables in shared memory.
// The time of a launch is the length of ONE stream's chain (16 384 symbols: 1.4 ms encode, 3.9 ms decode) for any number of
// streams up to the machine's warp slots, so the callers batch as many streams as they can into one launch.
#include "common.cuh"

namespace pccgeo {

constexpr int kRcPrecision = 16;
constexpr int kRcOverflowWidth = 4;
constexpr uint32_t kRcMaxOverflow = (1u << kRcOverflowWidth) - 1;
constexpr uint32_t kRcTop = 1u << 24;

// ---------------------------------------------------------------- encoder core (host + device: the CPU tests run it too)
// Same arithmetic as the host coder's low/cache/cache_size machine, arranged for a GPU thread: `low` is 32 bits, a carry out of
// it is applied at once to the bytes already emitted -- the last four of them are still in a register (`tail`), so a carry
// is one increment; only when all four are 0xFF does it walk back through memory -- and every renormalisation step is one
// (delayed) byte store and two shifts, without the data-dependent flush loop of the cache/cache_size form.
struct RcEnc {
  uint32_t low, range;
  uint32_t tail;     // emitted bytes cnt-4 .. cnt-1, oldest in the top byte; not yet in memory
  uint32_t cnt;      // bytes emitted so far, counting the always-zero first byte (emitted "before" the stream starts)
  uint8_t* out;      // stream buffer, null on the lanes that only follow the arithmetic
  uint32_t cap;
};

__host__ __device__ __forceinline__ void rc_enc_init(RcEnc& s, uint8_t* out, uint32_t cap) {
  s.low = 0; s.range = 0xFFFFFFFFu; s.tail = 0; s.cnt = 1;
  s.out = out; s.cap = cap;
}
__host__ __device__ __forceinline__ void rc_enc_emit(RcEnc& s, uint32_t byte) {
  if (s.cnt >= 4 && s.out && s.cnt - 4 < s.cap) s.out[s.cnt - 4] = (uint8_t)(s.tail >> 24);
  s.tail = (s.tail << 8) | (byte & 0xffu);
  ++s.cnt;
}
__host__ __device__ __forceinline__ void rc_enc_carry(RcEnc& s) {
  if (++s.tail == 0 && s.out) {   // 0xFFFFFFFF + 1: the carry leaves the register (the zero first byte stops it at the latest)
    long long p = (long long)s.cnt - 5;
    while (p > 0 && (p >= s.cap || s.out[p] == 0xFF)) {
      if (p < s.cap) s.out[p] = 0;
      --p;
    }
    if (p >= 0 && p < s.cap) ++s.out[p];
  }
}
__host__ __device__ __forceinline__ void rc_enc_interval(RcEnc& s, uint32_t lower, uint32_t freq, int precision) {
  const uint32_t r = s.range >> precision;
  const uint32_t add = r * lower;   // r < 2^(32 - precision), lower < 2^precision
  s.low += add;
  if (s.low < add) rc_enc_carry(s);
  s.range = r * freq;
  while (s.range < kRcTop) {
    rc_enc_emit(s, s.low >> 24);
    s.low <<= 8;
    s.range <<= 8;
  }
}
// one mapped symbol: lf = cdf_lo | (freq-1) << 16; ov = 0, or 1 + the overflow value of an escape
__host__ __device__ __forceinline__ void rc_enc_symbol(RcEnc& s, uint32_t lf, uint32_t ov) {
  rc_enc_interval(s, lf & 0xffffu, (lf >> 16) + 1u, kRcPrecision);
  if (ov) {
    const uint32_t overflow = ov - 1u;
    int widths = 0;
    while (widths < 8 && (overflow >> (widths * kRcOverflowWidth)) != 0) ++widths;
    uint32_t val = (uint32_t)widths;
    while (val >= kRcMaxOverflow) {
      rc_enc_interval(s, kRcMaxOverflow, 1u, kRcOverflowWidth);
      val -= kRcMaxOverflow;
    }
    rc_enc_interval(s, val, 1u, kRcOverflowWidth);
    for (int k = 0; k < widths; ++k) rc_enc_interval(s, (overflow >> (k * kRcOverflowWidth)) & kRcMaxOverflow, 1u, kRcOverflowWidth);
  }
}
// terminator: the value in [low, low + range) with the most trailing zero bits; -> stream length (first byte and trailing
// zeros dropped; valid on the writing lane), or -1 when the buffer was too small
__host__ __device__ __forceinline__ int32_t rc_enc_finish(RcEnc& s) {
  const unsigned long long lo = s.low, hi = lo + s.range - 1;
  unsigned long long v = lo;
  for (int nbits = 32; nbits >= 0; --nbits) {
    const unsigned long long mask = (1ull << nbits) - 1;
    v = (lo + mask) & ~mask;
    if (v <= hi) break;
  }
  if (v >> 32) rc_enc_carry(s);
  s.low = (uint32_t)v;
  for (int k = 0; k < 4; ++k) {
    rc_enc_emit(s, s.low >> 24);
    s.low <<= 8;
  }
  if (s.cnt > s.cap) return -1;
  if (!s.out) return 0;
  for (int k = 0; k < 4; ++k) {   // the bytes still in the register
    const long long p = (long long)s.cnt - 4 + k;
    if (p >= 0) s.out[p] = (uint8_t)(s.tail >> (24 - 8 * k));
  }
  long long p = (long long)s.cnt - 1;
  while (p >= 1 && s.out[p] == 0) --p;
  return (int32_t)p;
}
// symbol -> (lf, ov); 0 = ok, 1 = table index out of range, 2 = the escape value does not fit the 32-bit `ov` word (|value| of the
// order of 2^31: the host coder carries 64 bits there; callers send such a latent through the host coder -- same bytes)
__host__ __device__ __forceinline__ int rc_map_symbol(int32_t sym, int row, const int32_t* cdf, int cdf_stride, const int32_t* cdf_length,
                                                       const int32_t* offset, int rows, uint32_t& lf, uint32_t& ov) {
  lf = 0; ov = 0;
  if (row < 0 || row >= rows) return 1;
  const int32_t* r = cdf + (long long)row * cdf_stride;
  const int32_t max_value = cdf_length[row] - 2;
  long long value = (long long)sym - offset[row];
  unsigned long long wide = 0;
  if (value < 0) { wide = (unsigned long long)(-2 * value - 1) + 1ull; value = max_value; }
  else if (value >= max_value) { wide = (unsigned long long)(2 * (value - max_value)) + 1ull; value = max_value; }
  if (wide > 0xffffffffull) return 2;
  ov = (uint32_t)wide;
  const uint32_t lo = (uint32_t)r[value], hi = (uint32_t)r[value + 1];
  lf = lo | ((hi - lo - 1) << 16);
  return 0;
}

// row of symbol i: indexes[i] (mode 0) or (position in stream / channel_stride) % rows (mode 1: one table per channel)
__global__ void rc_map_kernel(const int32_t* __restrict__ sym, const int32_t* __restrict__ indexes, const int32_t* __restrict__ cdf,
                              int cdf_stride, const int32_t* __restrict__ cdf_length, const int32_t* __restrict__ offset, int rows,
                              int index_mode, long long channel_stride, long long per_stream, long long count,
                              uint32_t* __restrict__ lf, uint32_t* __restrict__ ovf, int* __restrict__ err) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const int row = index_mode == 0 ? indexes[i] : (int)(((i % per_stream) / channel_stride) % rows);
    uint32_t a, b;
    const int rc = rc_map_symbol(sym[i], row, cdf, cdf_stride, cdf_length, offset, rows, a, b);
    if (rc) atomicOr(err, rc);
    lf[i] = a;
    ovf[i] = b;
  }
}

// grid = streams, block = 32
__global__ void __launch_bounds__(32) rc_encode_kernel(const uint32_t* __restrict__ lf, const uint32_t* __restrict__ ovf,
                                                       long long per_stream, uint8_t* __restrict__ out, uint32_t cap,
                                                       int32_t* __restrict__ lengths) {
  const int sidx = blockIdx.x, lane = threadIdx.x;
  const uint32_t* l = lf + (long long)sidx * per_stream;
  const uint32_t* o = ovf + (long long)sidx * per_stream;
  RcEnc s;
  rc_enc_init(s, lane == 0 ? out + (long long)sidx * cap : nullptr, cap);
  uint32_t nl = lane < per_stream ? l[lane] : 0u, no = lane < per_stream ? o[lane] : 0u;
  for (long long base = 0; base < per_stream; base += 32) {
    const uint32_t my_l = nl, my_o = no;
    const long long i = base + 32 + lane;   // next group's loads fly during this group's arithmetic
    nl = i < per_stream ? l[i] : 0u;
    no = i < per_stream ? o[i] : 0u;
    const int n = (int)min(32LL, per_stream - base);
    for (int j = 0; j < n; ++j)
      rc_enc_symbol(s, __shfl_sync(0xffffffffu, my_l, j), __shfl_sync(0xffffffffu, my_o, j));   // warp-uniform control flow
  }
  const int32_t len = rc_enc_finish(s);
  if (lane == 0) lengths[sidx] = len;
}

// streams packed back to back: offsets = exclusive prefix sum of the lengths (nstreams is small: one warp)
__global__ void rc_offsets_kernel(const int32_t* __restrict__ lengths, int nstreams, long long* __restrict__ offsets) {
  const int lane = threadIdx.x;
  long long carry = 0;
  for (int base = 0; base < nstreams; base += 32) {
    const int s = base + lane;
    long long v = s < nstreams && lengths[s] > 0 ? lengths[s] : 0;
    long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (s < nstreams) offsets[s] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) offsets[nstreams] = carry;
}
__global__ void rc_pack_kernel(const uint8_t* __restrict__ streams, uint32_t cap, const int32_t* __restrict__ lengths,
                               const long long* __restrict__ offsets, uint8_t* __restrict__ packed, long long packed_cap) {
  const int s = blockIdx.x;
  const int n = lengths[s];
  const uint8_t* src = streams + (long long)s * cap + 1;   // the first byte of a stream is always 0 and is not stored
  const long long o = offsets[s];
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (o + i < packed_cap) packed[o + i] = src[i];
}

// ---------------------------------------------------------------- decoder
// stream bytes through the lanes' registers: 128 bytes per refill (the next 128 already in flight); the current 4 bytes sit
// in a warp-uniform register, so a byte costs a shift and every fourth byte one shuffle
struct RcBytes {
  const uint8_t* p;
  long long len;
  long long chunk;     // index of the 128-byte chunk in `cur`
  uint32_t cur, nxt;   // this lane's 4 bytes of chunk / chunk + 1
  uint32_t word;       // bytes [pos & ~3, +4) of the chunk (warp-uniform)
  uint32_t pos;        // next byte within the chunk (warp-uniform)
};
__device__ __forceinline__ uint32_t rc_load4(const uint8_t* p, long long len, long long at) {
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (at + k < len) w |= (uint32_t)p[at + k] << (8 * k);
  return w;
}
__device__ __forceinline__ void rc_bytes_init(RcBytes& b, const uint8_t* p, long long len, int lane) {
  b.p = p; b.len = len; b.chunk = 0; b.pos = 0;
  b.cur = rc_load4(p, len, 4 * lane);
  b.nxt = rc_load4(p, len, 128 + 4 * lane);
  b.word = __shfl_sync(0xffffffffu, b.cur, 0);
}
__device__ __forceinline__ uint32_t rc_bytes_next(RcBytes& b, int lane) {
  const uint32_t v = (b.word >> ((b.pos & 3u) * 8u)) & 0xffu;
  ++b.pos;
  if ((b.pos & 3u) == 0) {
    if (b.pos == 128u) {
      b.pos = 0;
      ++b.chunk;
      b.cur = b.nxt;
      b.nxt = rc_load4(b.p, b.len, (b.chunk + 1) * 128 + 4 * lane);
    }
    b.word = __shfl_sync(0xffffffffu, b.cur, b.pos >> 2);
  }
  return v;
}

struct RcDec {
  uint32_t range, code;
};
__device__ __forceinline__ void rc_dec_normalize(RcDec& d, RcBytes& b, int lane) {
  while (d.range < kRcTop) {
    d.code = (d.code << 8) | rc_bytes_next(b, lane);
    d.range <<= 8;
  }
}
// min(code / r, 15) without the division: the number of k in 1..15 with k * r <= code, one k per lane
__device__ __forceinline__ uint32_t rc_dec_uniform(RcDec& d, RcBytes& b, int lane) {
  const uint32_t r = d.range >> kRcOverflowWidth;   // < 2^28: lane * r fits
  const uint32_t s = __popc(__ballot_sync(0xffffffffu, lane >= 1 && lane <= (int)kRcMaxOverflow && (uint32_t)lane * r <= d.code));
  d.code -= r * s;
  d.range = r;
  rc_dec_normalize(d, b, lane);
  return s;
}

// The symbol search is lane-parallel and division-free: with r = range >> 16, symbol s is the largest one with
// cdf[s] * r <= code (== cdf[s] <= min(code / r, 65535), the host decoder's test, since cdf[s] < 65536 for s < n).
// Level 1: lane L holds the pivot cdf[L * step], step = n / 32 + 1 -- fetched while the PREVIOUS symbol was decoded, the
// table row of a symbol does not depend on the coder state -- and a ballot picks the segment; rows of up to 31 symbols
// (scales < 5.4: most of a trained model's latents) are resolved right there.  Level 2: 32 lanes compare 32 consecutive
// entries of the segment: one dependent table load per symbol -- from SHARED memory: every CTA first copies the compact
// tables (16-bit entries, rows back to back: 27 KB for the 64 Gaussian scale rows), because an L2 round trip per symbol
// (the tables do not fit L1) was most of the chain (4.1 ms per 16 384-symbol stream with the tables in global memory).
// grid = ceil(streams / kRcDecWarps), block = kRcDecWarps warps, one stream per warp.
constexpr int kRcDecWarps = 4;

__global__ void __launch_bounds__(kRcDecWarps * 32) rc_decode_kernel(
    const uint8_t* __restrict__ bytes, const long long* __restrict__ byte_offsets, const int32_t* __restrict__ indexes, int nstreams,
    long long per_stream, const uint16_t* __restrict__ cdf16, const int32_t* __restrict__ row_start, const int32_t* __restrict__ cdf_length,
    const int32_t* __restrict__ offset, int rows, int total_entries, int index_mode, long long channel_stride, int32_t* __restrict__ out,
    int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char rc_smem[];
  int32_t* s_start = (int32_t*)rc_smem;
  int32_t* s_len = s_start + rows;
  int32_t* s_off = s_len + rows;
  uint16_t* s_cdf = (uint16_t*)(s_off + rows);
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    s_start[i] = row_start[i];
    s_len[i] = cdf_length[i];
    s_off[i] = offset[i];
  }
  for (int i = threadIdx.x; i < total_entries; i += blockDim.x) s_cdf[i] = cdf16[i];
  __syncthreads();   // the only block-wide step: the warps are independent from here on

  const int sidx = blockIdx.x * kRcDecWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (sidx >= nstreams) return;
  const long long b0 = byte_offsets[sidx], b1 = byte_offsets[sidx + 1];
  RcBytes bs;
  rc_bytes_init(bs, bytes + b0, b1 - b0, lane);
  RcDec d;
  d.range = 0xFFFFFFFFu;
  d.code = 0;
  for (int k = 0; k < 4; ++k) d.code = (d.code << 8) | rc_bytes_next(bs, lane);
  const int32_t* idx = index_mode == 0 ? indexes + (long long)sidx * per_stream : nullptr;
  int32_t* o = out + (long long)sidx * per_stream;
  bool bad = false;

  auto row_of = [&](long long i) -> int {
    int r = 0;
    if (i < per_stream) r = index_mode == 0 ? idx[i] : (int)((i / channel_stride) % rows);
    if (r < 0 || r >= rows) { bad = true; r = 0; }
    return r;
  };
  // entry k of a row with n symbols: cdf[n] = 2^16 does not fit the 16-bit table and is implied; beyond the row: "infinite"
  auto entry = [&](int start, int k, int n) -> uint32_t { return k < n ? (uint32_t)s_cdf[start + k] : (k == n ? 0x10000u : 0x7fffffffu); };
  int nrow = row_of(lane);
  for (long long base = 0; base < per_stream; base += 32) {
    const int my_row = nrow;
    nrow = row_of(base + 32 + lane);   // next group's table indexes fly during this group's decoding
    const int cnt_syms = (int)min(32LL, per_stream - base);
    int32_t my_out = 0;
    // Fast path: every lane prepares, for ITS symbol of the group, the interval of the value 0 (table entry -offset: the mode of
    // the zero-mean Gaussian rows and of a trained factorized prior) as cdf_lo | (freq - 1) << 16.  Decoding symbol j then starts
    // with ONE shuffle and a range test; only when the value is not 0 does the lane-parallel search below run.  At codec
    // operating points (~98 % zeros) the chain per symbol shrinks from ~120 to ~25 instructions; the result is the same
    // symbol the search would find (it is the s with cdf[s] * r <= code < cdf[s+1] * r).
    uint32_t my_pack = 0;
    bool my_fast = false;
    {
      const int st_ = s_start[my_row], n_ = s_len[my_row] - 1, m_ = -s_off[my_row];
      if (m_ >= 0 && m_ < n_ - 1) {   // a regular symbol (not the escape slot)
        const uint32_t lo_ = entry(st_, m_, n_), hi_ = entry(st_, m_ + 1, n_);
        if (hi_ > lo_) { my_pack = lo_ | (hi_ - lo_ - 1u) << 16; my_fast = true; }
      }
    }
    const uint32_t fast_mask = __ballot_sync(0xffffffffu, my_fast);
    for (int j = 0; j < cnt_syms; ++j) {
      const uint32_t r = d.range >> kRcPrecision;   // < 2^16, cdf <= 2^16: the products fit
      if ((fast_mask >> j) & 1u) {
        const uint32_t pk = __shfl_sync(0xffffffffu, my_pack, j);
        const uint32_t lo_r = (pk & 0xffffu) * r, fr = ((pk >> 16) + 1u) * r;
        if (d.code - lo_r < fr) {    // unsigned: also false when code < lo_r.  Warp-uniform: every lane carries the coder state
          d.code -= lo_r;
          d.range = fr;
          rc_dec_normalize(d, bs, lane);
          continue;                  // my_out of lane j stays 0
        }
      }
      const int row = __shfl_sync(0xffffffffu, my_row, j);
      const int start = s_start[row], n = s_len[row] - 1, off = s_off[row];   // n symbols incl. the escape slot; cdf[0..n]
      const int step = (n >> 5) + 1;
      const uint32_t piv = entry(start, lane * step, n);
      const int seg = __popc(__ballot_sync(0xffffffffu, lane >= 1 && lane * step < n && piv * r <= d.code));
      int lo = seg * step;
      uint32_t c_lo, c_hi;
      if (step == 1) {   // n <= 31: the pivots are the row
        c_lo = __shfl_sync(0xffffffffu, piv, seg);
        c_hi = __shfl_sync(0xffffffffu, piv, seg + 1);
      } else {
        uint32_t c;
        int cnt;
        do {   // lane L looks at cdf[lo + L]; lanes >= 1 vote whether the symbol is at or beyond lo + L
          const int k = lo + lane;
          c = entry(start, k, n);
          cnt = __popc(__ballot_sync(0xffffffffu, lane >= 1 && k < n && c * r <= d.code));
          lo += cnt;
        } while (cnt == 31);
        c_lo = __shfl_sync(0xffffffffu, c, cnt);
        c_hi = __shfl_sync(0xffffffffu, c, cnt + 1);
      }
      d.code -= r * c_lo;
      d.range = r * (c_hi - c_lo);
      rc_dec_normalize(d, bs, lane);
      long long v = lo;
      if (lo == n - 1) {   // escape (warp-uniform)
        int widths = 0;
        for (;;) {
          const uint32_t w = rc_dec_uniform(d, bs, lane);
          widths += (int)w;
          if (w != kRcMaxOverflow) break;
          if (widths > 64) break;
        }
        if (widths > 16) { bad = true; widths = 0; }
        unsigned long long overflow = 0;
        for (int q = 0; q < widths; ++q) overflow |= (unsigned long long)rc_dec_uniform(d, bs, lane) << (q * kRcOverflowWidth);
        v = (long long)(overflow >> 1);
        if (overflow & 1) v = -v - 1; else v += n - 1;
      }
      if (lane == j) my_out = (int32_t)(v + off);
    }
    if (base + lane < per_stream) o[base + lane] = my_out;
  }
  if (bad) *err = 1;
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" size_t pccgeo_rc_encode_ws_bytes(int nstreams, long long per_stream) {
  const size_t count = (size_t)nstreams * (size_t)per_stream;
  return count * 8 + (size_t)nstreams * ((size_t)per_stream * 4 + 64) + 256;   // lf + ovf, per-stream byte buffers
}

extern "C" int pccgeo_range_encode_device(const int32_t* symbols, const int32_t* indexes, int nstreams, long long per_stream,
                                          const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, const int32_t* offset, int rows,
                                          int index_mode, long long channel_stride, void* ws, uint8_t* packed, long long packed_capacity,
                                          int32_t* lengths, long long* offsets, int* err, void* stream) {
  PCCGEO_REQUIRE(symbols && cdf && cdf_length && offset && ws && packed && lengths && offsets && err, "range_encode_device: null pointer");
  PCCGEO_REQUIRE(nstreams > 0 && per_stream > 0 && rows > 0 && packed_capacity > 0 && (index_mode == 0 || index_mode == 1),
                 "range_encode_device: bad argument");
  PCCGEO_REQUIRE(index_mode == 1 || indexes, "range_encode_device: index mode 0 needs indexes");
  PCCGEO_REQUIRE(index_mode == 0 || channel_stride > 0, "range_encode_device: index mode 1 needs channel_stride");
  cudaStream_t st = (cudaStream_t)stream;
  const long long count = (long long)nstreams * per_stream;
  const uint32_t cap = (uint32_t)(per_stream * 4 + 64);
  uint32_t* lf = (uint32_t*)ws;
  uint32_t* ovf = lf + count;
  uint8_t* streams = (uint8_t*)(ovf + count);
  long long b = (count + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  rc_map_kernel<<<(int)b, 256, 0, st>>>(symbols, indexes, cdf, cdf_stride, cdf_length, offset, rows, index_mode, channel_stride, per_stream,
                                        count, lf, ovf, err);
  int rc = check_launch("rc_map_kernel");
  if (rc) return rc;
  rc_encode_kernel<<<nstreams, 32, 0, st>>>(lf, ovf, per_stream, streams, cap, lengths);
  rc = check_launch("rc_encode_kernel");
  if (rc) return rc;
  rc_offsets_kernel<<<1, 32, 0, st>>>(lengths, nstreams, offsets);
  rc = check_launch("rc_offsets_kernel");
  if (rc) return rc;
  rc_pack_kernel<<<nstreams, 256, 0, st>>>(streams, cap, lengths, offsets, packed, packed_capacity);
  return check_launch("rc_pack_kernel");
}

extern "C" int pccgeo_range_decode_device(const uint8_t* bytes, const long long* byte_offsets, const int32_t* indexes, int nstreams,
                                          long long per_stream, const uint16_t* cdf16, const int32_t* row_start, const int32_t* cdf_length,
                                          const int32_t* offset, int rows, int total_entries, int index_mode, long long channel_stride,
                                          int32_t* symbols_out, int* err, void* stream) {
  PCCGEO_REQUIRE(bytes && byte_offsets && cdf16 && row_start && cdf_length && offset && symbols_out && err, "range_decode_device: null pointer");
  PCCGEO_REQUIRE(nstreams > 0 && per_stream > 0 && rows > 0 && total_entries > 0 && (index_mode == 0 || index_mode == 1),
                 "range_decode_device: bad argument");
  PCCGEO_REQUIRE(index_mode == 1 || indexes, "range_decode_device: index mode 0 needs indexes");
  PCCGEO_REQUIRE(index_mode == 0 || channel_stride > 0, "range_decode_device: index mode 1 needs channel_stride");
  const size_t smem = (size_t)rows * 12 + (size_t)total_entries * 2;
  PCCGEO_REQUIRE(smem <= 200 * 1024, "range_decode_device: tables of %zu bytes do not fit shared memory", smem);
  PCCGEO_CUDA(cudaFuncSetAttribute(rc_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rc_decode_kernel<<<(nstreams + kRcDecWarps - 1) / kRcDecWarps, kRcDecWarps * 32, smem, (cudaStream_t)stream>>>(
      bytes, byte_offsets, indexes, nstreams, per_stream, cdf16, row_start, cdf_length, offset, rows, total_entries, index_mode, channel_stride,
      symbols_out, err);
  return check_launch("rc_decode_kernel");
}

// The decoder's tables, on the host (uploaded once per table set): the rows' first cdf_length[r] - 1 entries as 16 bits, back to
// back (the last entry of a row, 2^16, is implied); row_start (rows) receives the rows' positions.  -> number of entries, or
// -1 if a row is not a valid 16-bit CDF.  Pass cdf16 == NULL to query the size.
extern "C" long long pccgeo_range_compact_tables_host(const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, int rows, uint16_t* cdf16,
                                                      int32_t* row_start) {
  if (!cdf || !cdf_length || rows <= 0) {
    set_error("range_compact_tables: bad argument");
    return -1;
  }
  long long pos = 0;
  for (int r = 0; r < rows; ++r) {
    const int32_t* row = cdf + (long long)r * cdf_stride;
    const int n = cdf_length[r] - 1;
    if (n < 1 || n + 1 > cdf_stride || row[0] != 0 || row[n] != (1 << kRcPrecision)) {
      set_error("range_compact_tables: row %d is not a %d-bit CDF", r, kRcPrecision);
      return -1;
    }
    for (int k = 0; k < n; ++k) {
      if (row[k] < 0 || row[k] >= (1 << kRcPrecision) || row[k + 1] < row[k]) {
        set_error("range_compact_tables: row %d is not a %d-bit CDF", r, kRcPrecision);
        return -1;
      }
      if (cdf16) cdf16[pos + k] = (uint16_t)row[k];
    }
    if (cdf16 && row_start) row_start[r] = (int32_t)pos;
    pos += n;
  }
  return pos;
}

// The device encoder's arithmetic (rc_map_symbol + rc_enc_symbol + rc_enc_finish, the same inline functions the kernels
// call) run sequentially on the host over HOST arrays: lets the CPU test-suite pin it against the production host coder
// without a GPU.  Test hook, not a product path.
extern "C" int pccgeo_range_encode_emulate_host(const int32_t* symbols, const int32_t* indexes, int nstreams, long long per_stream,
                                                const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, const int32_t* offset,
                                                int rows, int index_mode, long long channel_stride, uint8_t* packed,
                                                long long packed_capacity, int32_t* lengths, long long* offsets) {
  PCCGEO_REQUIRE(symbols && cdf && cdf_length && offset && packed && lengths && offsets, "range_encode_emulate: null pointer");
  const uint32_t cap = (uint32_t)(per_stream * 4 + 64);
  uint8_t* buf = (uint8_t*)malloc(cap);
  long long pos = 0;
  for (int s = 0; s < nstreams; ++s) {
    RcEnc e;
    rc_enc_init(e, buf, cap);
    for (long long i = 0; i < per_stream; ++i) {
      const long long g = (long long)s * per_stream + i;
      const int row = index_mode == 0 ? indexes[g] : (int)((i / channel_stride) % rows);
      uint32_t lf, ov;
      if (const int rc = rc_map_symbol(symbols[g], row, cdf, cdf_stride, cdf_length, offset, rows, lf, ov)) {
        free(buf);
        set_error(rc == 1 ? "range_encode_emulate: table index out of range"
                          : "range_encode_emulate: escape value beyond 32 bits (host coder only)");
        return PCCGEO_EINVAL;
      }
      rc_enc_symbol(e, lf, ov);
    }
    lengths[s] = rc_enc_finish(e);
    offsets[s] = pos;
    for (int i = 0; i < lengths[s]; ++i)
      if (pos + i < packed_capacity) packed[pos + i] = buf[1 + i];
    pos += lengths[s] > 0 ? lengths[s] : 0;
  }
  offsets[nstreams] = pos;
  free(buf);
  return PCCGEO_OK;
}

// Test hook for the one branch of the device encoder that random data practically never takes: a carry that leaves the
// 4-byte register and walks back through memory.  buf holds n already-stored bytes; -> the register after the carry.
extern "C" unsigned pccgeo_rc_carry_probe_host(uint8_t* buf, int n) {
  RcEnc e;
  rc_enc_init(e, buf, (uint32_t)n + 16);
  e.tail = 0xFFFFFFFFu;
  e.cnt = (uint32_t)n + 4;
  rc_enc_carry(e);
  return e.tail;
}
