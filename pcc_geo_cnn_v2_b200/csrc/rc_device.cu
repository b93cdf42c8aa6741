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
//           CDF search is lane-parallel: a 256-entry LUT gives the start, 32 lanes compare 32 consecutive CDF entries,
//           ballot/popc gives the symbol -- one dependent table load instead of a binary search.
// A warp needs < 64 registers and no shared memory, so these kernels co-reside with the persistent conv CTAs of the
// transforms that run meanwhile on the other streams of the block loops.
#include "common.cuh"

namespace pccgeo {

constexpr int kRcPrecision = 16;
constexpr int kRcOverflowWidth = 4;
constexpr uint32_t kRcMaxOverflow = (1u << kRcOverflowWidth) - 1;
constexpr uint32_t kRcTop = 1u << 24;

// ---------------------------------------------------------------- encoder core (host + device: the CPU tests run it too)
struct RcEnc {
  unsigned long long low;
  uint32_t range, cache, cache_size;
  uint8_t* out;      // stream buffer, null on the lanes that only follow the arithmetic
  uint32_t pos;      // bytes emitted so far (including the always-zero first byte)
  uint32_t last_nz;  // index after the last non-zero byte
  uint32_t cap;
};

__host__ __device__ __forceinline__ void rc_enc_init(RcEnc& s, uint8_t* out, uint32_t cap) {
  s.low = 0; s.range = 0xFFFFFFFFu; s.cache = 0; s.cache_size = 1;
  s.out = out; s.pos = 0; s.last_nz = 0; s.cap = cap;
}
__host__ __device__ __forceinline__ void rc_enc_put(RcEnc& s, uint32_t byte) {
  byte &= 0xffu;
  if (s.out && s.pos < s.cap) s.out[s.pos] = (uint8_t)byte;
  ++s.pos;
  if (byte) s.last_nz = s.pos;
}
__host__ __device__ __forceinline__ void rc_enc_shift_low(RcEnc& s) {
  if ((uint32_t)s.low < 0xFF000000u || (s.low >> 32) != 0) {
    const uint32_t carry = (uint32_t)(s.low >> 32);
    uint32_t temp = s.cache;
    do {
      rc_enc_put(s, temp + carry);
      temp = 0xFF;
    } while (--s.cache_size != 0);
    s.cache = (uint32_t)(s.low >> 24) & 0xffu;
  }
  ++s.cache_size;
  s.low = (s.low & 0x00FFFFFFull) << 8;
}
__host__ __device__ __forceinline__ void rc_enc_interval(RcEnc& s, uint32_t lower, uint32_t freq, int precision) {
  const uint32_t r = s.range >> precision;
  s.low += (unsigned long long)r * lower;
  s.range = r * freq;
  while (s.range < kRcTop) {
    rc_enc_shift_low(s);
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
// zeros dropped), or -1 when the buffer was too small
__host__ __device__ __forceinline__ int32_t rc_enc_finish(RcEnc& s) {
  const unsigned long long hi = s.low + s.range - 1;
  for (int nbits = 32; nbits >= 0; --nbits) {
    const unsigned long long mask = (1ull << nbits) - 1;
    const unsigned long long v = (s.low + mask) & ~mask;
    if (v <= hi) { s.low = v; break; }
  }
  for (int k = 0; k < 5; ++k) rc_enc_shift_low(s);
  return s.pos > s.cap ? -1 : (s.last_nz > 1 ? (int32_t)(s.last_nz - 1) : 0);
}
// symbol -> (lf, ov); false when the table index is out of range
__host__ __device__ __forceinline__ bool rc_map_symbol(int32_t sym, int row, const int32_t* cdf, int cdf_stride, const int32_t* cdf_length,
                                                       const int32_t* offset, int rows, uint32_t& lf, uint32_t& ov) {
  lf = 0; ov = 0;
  if (row < 0 || row >= rows) return false;
  const int32_t* r = cdf + (long long)row * cdf_stride;
  const int32_t max_value = cdf_length[row] - 2;
  long long value = (long long)sym - offset[row];
  if (value < 0) { ov = (uint32_t)(-2 * value - 1) + 1u; value = max_value; }
  else if (value >= max_value) { ov = (uint32_t)(2 * (value - max_value)) + 1u; value = max_value; }
  const uint32_t lo = (uint32_t)r[value], hi = (uint32_t)r[value + 1];
  lf = lo | ((hi - lo - 1) << 16);
  return true;
}

// row of symbol i: indexes[i] (mode 0) or (position in stream / channel_stride) % rows (mode 1: one table per channel)
__global__ void rc_map_kernel(const int32_t* __restrict__ sym, const int32_t* __restrict__ indexes, const int32_t* __restrict__ cdf,
                              int cdf_stride, const int32_t* __restrict__ cdf_length, const int32_t* __restrict__ offset, int rows,
                              int index_mode, long long channel_stride, long long per_stream, long long count,
                              uint32_t* __restrict__ lf, uint32_t* __restrict__ ovf, int* __restrict__ err) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const int row = index_mode == 0 ? indexes[i] : (int)(((i % per_stream) / channel_stride) % rows);
    uint32_t a, b;
    if (!rc_map_symbol(sym[i], row, cdf, cdf_stride, cdf_length, offset, rows, a, b)) *err = 1;
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
// stream bytes through the lanes' registers: 128 bytes per refill, the next 128 already in flight
struct RcBytes {
  const uint8_t* p;
  long long len;
  long long chunk;     // index of the 128-byte chunk in `cur`
  uint32_t cur, nxt;   // this lane's 4 bytes of chunk / chunk + 1
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
}
__device__ __forceinline__ uint32_t rc_bytes_next(RcBytes& b, int lane) {
  const uint32_t w = __shfl_sync(0xffffffffu, b.cur, b.pos >> 2);
  const uint32_t v = (w >> ((b.pos & 3u) * 8u)) & 0xffu;
  if (++b.pos == 128u) {
    b.pos = 0;
    ++b.chunk;
    b.cur = b.nxt;
    b.nxt = rc_load4(b.p, b.len, (b.chunk + 1) * 128 + 4 * lane);
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
__device__ __forceinline__ uint32_t rc_dec_uniform(RcDec& d, RcBytes& b, int lane) {
  const uint32_t r = d.range >> kRcOverflowWidth;
  uint32_t s = d.code / r;
  if (s > kRcMaxOverflow) s = kRcMaxOverflow;
  d.code -= r * s;
  d.range = r;
  rc_dec_normalize(d, b, lane);
  return s;
}

// grid = streams, block = 32.  lut (rows, 256): lut[r][v] = largest s in [0, n) with cdf[r][s] <= v << 8.
__global__ void __launch_bounds__(32) rc_decode_kernel(const uint8_t* __restrict__ bytes, const long long* __restrict__ byte_offsets,
                                                       const int32_t* __restrict__ indexes, long long per_stream,
                                                       const int32_t* __restrict__ cdf, int cdf_stride,
                                                       const int32_t* __restrict__ cdf_length, const int32_t* __restrict__ offset,
                                                       const uint16_t* __restrict__ lut, int rows, int index_mode,
                                                       long long channel_stride, int32_t* __restrict__ out, int* __restrict__ err) {
  const int sidx = blockIdx.x, lane = threadIdx.x;
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
  int nrow = row_of(lane);
  int nlen = cdf_length[nrow], noff = offset[nrow];
  for (long long base = 0; base < per_stream; base += 32) {
    const int my_row = nrow, my_len = nlen, my_off = noff;
    nrow = row_of(base + 32 + lane);   // next group's table parameters fly during this group's decoding
    nlen = cdf_length[nrow];
    noff = offset[nrow];
    const int cnt_syms = (int)min(32LL, per_stream - base);
    int32_t my_out = 0;
    for (int j = 0; j < cnt_syms; ++j) {
      const int row = __shfl_sync(0xffffffffu, my_row, j);
      const int n = __shfl_sync(0xffffffffu, my_len, j) - 1;   // symbols incl. the escape slot; cdf[0..n]
      const int off = __shfl_sync(0xffffffffu, my_off, j);
      const int32_t* rowp = cdf + (long long)row * cdf_stride;
      const uint32_t r = d.range >> kRcPrecision;
      uint32_t value = d.code / r;
      if (value > 0xffffu) value = 0xffffu;
      int lo = lut[row * 256 + (int)(value >> 8)];
      int32_t c;
      int cnt;
      do {   // lane L looks at cdf[lo + L]; lanes >= 1 vote whether the symbol is at or beyond lo + L
        const int k = lo + lane;
        c = k <= n ? rowp[k] : 0x7fffffff;
        const unsigned m = __ballot_sync(0xffffffffu, lane >= 1 && k < n && (uint32_t)c <= value);
        cnt = __popc(m);
        lo += cnt;
      } while (cnt == 31);
      const uint32_t c_lo = (uint32_t)__shfl_sync(0xffffffffu, c, cnt), c_hi = (uint32_t)__shfl_sync(0xffffffffu, c, cnt + 1);
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
                                          long long per_stream, const int32_t* cdf, int cdf_stride, const int32_t* cdf_length,
                                          const int32_t* offset, const uint16_t* lut, int rows, int index_mode, long long channel_stride,
                                          int32_t* symbols_out, int* err, void* stream) {
  PCCGEO_REQUIRE(bytes && byte_offsets && cdf && cdf_length && offset && lut && symbols_out && err, "range_decode_device: null pointer");
  PCCGEO_REQUIRE(nstreams > 0 && per_stream > 0 && rows > 0 && (index_mode == 0 || index_mode == 1), "range_decode_device: bad argument");
  PCCGEO_REQUIRE(index_mode == 1 || indexes, "range_decode_device: index mode 0 needs indexes");
  PCCGEO_REQUIRE(index_mode == 0 || channel_stride > 0, "range_decode_device: index mode 1 needs channel_stride");
  rc_decode_kernel<<<nstreams, 32, 0, (cudaStream_t)stream>>>(bytes, byte_offsets, indexes, per_stream, cdf, cdf_stride, cdf_length, offset,
                                                              lut, rows, index_mode, channel_stride, symbols_out, err);
  return check_launch("rc_decode_kernel");
}

// The decoder's search accelerators, on the host (uploaded once per table set): lut[r][b] = largest s in [0, n) with
// cdf[r][s] <= b << 8, n = cdf_length[r] - 1.
extern "C" int pccgeo_range_lut_host(const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, int rows, uint16_t* lut) {
  PCCGEO_REQUIRE(cdf && cdf_length && lut && rows > 0, "range_lut: bad argument");
  for (int r = 0; r < rows; ++r) {
    const int32_t* row = cdf + (long long)r * cdf_stride;
    const int n = cdf_length[r] - 1;
    int s = 0;
    for (int b = 0; b < 256; ++b) {
      const int32_t v = b << (kRcPrecision - 8);
      while (s + 1 < n && row[s + 1] <= v) ++s;
      lut[(size_t)r * 256 + b] = (uint16_t)s;
    }
  }
  return PCCGEO_OK;
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
      if (!rc_map_symbol(symbols[g], row, cdf, cdf_stride, cdf_length, offset, rows, lf, ov)) {
        free(buf);
        set_error("range_encode_emulate: table index out of range");
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
