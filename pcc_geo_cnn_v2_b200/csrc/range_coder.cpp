// Host range coder: the symbol-level contract of tensorflow-compression 1.3's
// unbounded_index_range_encode/_decode (precision 16, overflow_width 4; reference call sites
// src/utils/patch_gaussian_conditional.py:27-31 and the .compress/.decompress methods used at
// src/model_types.py:291-292,382-387,404-407) over this project's own 32-bit carry-propagating range coder.
// Byte format: DESIGN.md "Bitstream".  Independent streams (one per block and per latent) are spread over
// worker threads; there is no shared state between streams.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../include/pccgeo.h"

namespace pccgeo {
void set_error(const char* fmt, ...);
}

namespace {

constexpr uint32_t kTop = 1u << 24;

// The low/cache/cache_size machine of an LZMA-style range encoder, restated without its data-dependent loops (same bytes):
// `low` is 32 bits; a carry out of it is added at once to the bytes already written (a walk back over 0xFF bytes: the
// always-zero first byte stops it at the latest); renormalisation emits the `nb` leading bytes of `low`, nb = number of
// leading zero bytes of `range`, with one unaligned 4-byte store.  Whether a symbol emits 0, 1 or 2 bytes is close to a coin
// flip, so the byte loop mispredicted on most symbols (21 -> 13 ns per symbol on the bench's latents).
struct Encoder {
  uint32_t low = 0;
  uint32_t range = 0xFFFFFFFFu;
  std::vector<uint8_t> out;  // out[0]: the always-zero first byte; every encode() may write 4 bytes at pos
  size_t pos = 1;

  explicit Encoder(size_t reserve) : out(reserve + 64, 0) {}
  inline void ensure() {  // once per symbol: a symbol emits at most 2 + 11 bytes (escape: 3 + 8 four-bit intervals)
    if (pos + 32 > out.size()) out.resize(out.size() * 2);
  }
  inline void carry() {
    uint8_t* p = out.data() + pos - 1;
    while (*p == 0xFF) *p-- = 0;
    ++*p;
  }
  inline void put4() {
    const uint32_t be = __builtin_bswap32(low);
    std::memcpy(out.data() + pos, &be, 4);
  }
  inline void encode(uint32_t lower, uint32_t upper, int precision) {
    const uint32_t r = range >> precision;
    const uint32_t add = r * lower;  // r < 2^(32 - precision), lower < 2^precision
    low += add;
    if (low < add) carry();
    range = r * (upper - lower);
    const int nb = __builtin_clz(range) >> 3;  // range >= 2^8: nb <= 2
    put4();
    pos += nb;
    low <<= 8 * nb;
    range <<= 8 * nb;
  }
  void finish() {
    // the value in [low, low + range) with the most trailing zero bits
    const uint64_t lo = low, hi = lo + range - 1;
    uint64_t v = lo;
    for (int nbits = 32; nbits >= 0; --nbits) {
      const uint64_t mask = (1ull << nbits) - 1;
      v = (lo + mask) & ~mask;
      if (v <= hi) break;
    }
    if (v >> 32) carry();
    low = (uint32_t)v;
    ensure();
    put4();
    pos += 4;
    while (pos > 1 && out[pos - 1] == 0) --pos;  // trailing zero bytes are implied
    out.resize(pos);
    out.erase(out.begin());  // the first byte is always 0
  }
};

struct Decoder {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t range = 0xFFFFFFFFu;
  uint32_t code = 0;
  Decoder(const uint8_t* b, const uint8_t* e) : p(b), end(e) {
    for (int i = 0; i < 4; ++i) code = (code << 8) | next();
  }
  inline uint32_t next() { return p < end ? *p++ : (++p, 0u); }
  // Renormalisation without a data-dependent loop: the number of bytes to shift in is the number of leading zero bytes of
  // `range` (the loop `while (range < 2^24) { code = code << 8 | next(); range <<= 8; }` in closed form), and they are taken
  // from one unaligned big-endian 4-byte load.  Whether a symbol needs 0, 1 or 2 bytes is close to a coin flip, so the loop's
  // branch mispredicted on most symbols.  Past the end of the stream the bytes read as zero (next() semantics).
  inline void normalize() {
    const int nb = __builtin_clz(range) >> 3;  // range != 0
    uint32_t w;
    if (end - p >= 4) {
      std::memcpy(&w, p, 4);
      w = __builtin_bswap32(w);
    } else {
      w = 0;
      for (int k = 0; k < 4; ++k) w = (w << 8) | (p + k < end ? p[k] : 0u);
    }
    code = (uint32_t)(((((uint64_t)code) << 32) | w) >> (32 - 8 * nb));
    range <<= 8 * nb;
    p += nb;
  }
  // lut (optional, 256 entries): lut[b] = largest symbol s with cdf[s] <= (b << (precision-8)); the search then scans
  // forward from there -- for the peaked tables of this codec that is 0-2 steps instead of an 11-step binary search.
  // Try ONE candidate symbol m (the callers pass the table entry of the value 0, the mode of this codec's tables: ~98 % of the
  // latents at its operating points): cdf[m] * r <= code < cdf[m+1] * r is exactly "decode() would return m", without the
  // division and the search.  Returns false (state untouched) when the symbol is another one.
  inline bool try_symbol(const int32_t* cdf, int m, int precision) {
    const uint32_t r = range >> precision;
    const uint32_t lo = r * (uint32_t)cdf[m], fr = r * (uint32_t)(cdf[m + 1] - cdf[m]);
    if (code - lo >= fr) return false;   // unsigned: also when code < lo
    code -= lo;
    range = fr;
    normalize();
    return true;
  }
  inline int decode(const int32_t* cdf, int n, int precision, const uint16_t* lut = nullptr) {
    const uint32_t r = range >> precision;
    uint32_t value = code / r;
    const uint32_t maxv = (1u << precision) - 1;
    if (value > maxv) value = maxv;
    int lo = 0, hi = n;
    if (lut) {
      lo = lut[value >> (precision - 8)];
      while (lo + 1 < n && (uint32_t)cdf[lo + 1] <= value) ++lo;
      hi = lo + 1;
    }
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((uint32_t)cdf[mid] <= value) lo = mid; else hi = mid;
    }
    code -= r * (uint32_t)cdf[lo];
    range = r * (uint32_t)(cdf[lo + 1] - cdf[lo]);
    normalize();
    return lo;
  }
  inline uint32_t decode_uniform(int bits) {
    const uint32_t r = range >> bits;
    uint32_t s = code / r;
    const uint32_t maxv = (1u << bits) - 1;
    if (s > maxv) s = maxv;
    code -= r * s;
    range = r;
    normalize();
    return s;
  }
};

constexpr int kPrecision = 16;
constexpr int kOverflowWidth = 4;

struct Tables {
  const int32_t* cdf;
  int cdf_stride;
  const int32_t* cdf_length;
  const int32_t* offset;
  int rows;
  int index_mode;
  long long channel_stride;
  const uint16_t* lut = nullptr;  // (rows, 256) decoder search accelerators, or null
  inline int index(const int32_t* idx, long long pos) const {
    return index_mode == 0 ? idx[pos] : (int)((pos / channel_stride) % rows);
  }
};

bool encode_stream(const Tables& t, const int32_t* sym, const int32_t* idx, long long n, Encoder& enc) {
  const uint32_t max_overflow = (1u << kOverflowWidth) - 1;
  for (long long i = 0; i < n; ++i) {
    const int row_i = t.index(idx, i);
    if (row_i < 0 || row_i >= t.rows) return false;
    const int32_t* row = t.cdf + (long long)row_i * t.cdf_stride;
    const int32_t max_value = t.cdf_length[row_i] - 2;
    long long value = (long long)sym[i] - t.offset[row_i];
    uint64_t overflow = 0;
    if (value < 0) {
      overflow = (uint64_t)(-2 * value - 1);
      value = max_value;
    } else if (value >= max_value) {
      overflow = (uint64_t)(2 * (value - max_value));
      value = max_value;
    }
    enc.ensure();
    enc.encode((uint32_t)row[value], (uint32_t)row[value + 1], kPrecision);
    if (value == max_value) {
      int widths = 0;
      while ((overflow >> (widths * kOverflowWidth)) != 0) ++widths;
      uint32_t val = (uint32_t)widths;
      while (val >= max_overflow) {
        enc.encode(max_overflow, max_overflow + 1, kOverflowWidth);
        val -= max_overflow;
      }
      enc.encode(val, val + 1, kOverflowWidth);
      for (int j = 0; j < widths; ++j) {
        val = (uint32_t)((overflow >> (j * kOverflowWidth)) & max_overflow);
        enc.encode(val, val + 1, kOverflowWidth);
      }
    }
  }
  enc.finish();
  return true;
}

bool decode_stream(const Tables& t, const uint8_t* b, const uint8_t* e, const int32_t* idx, long long n, int32_t* out) {
  Decoder dec(b, e);
  const uint32_t max_overflow = (1u << kOverflowWidth) - 1;
  for (long long i = 0; i < n; ++i) {
    const int row_i = t.index(idx, i);
    if (row_i < 0 || row_i >= t.rows) return false;
    const int32_t* row = t.cdf + (long long)row_i * t.cdf_stride;
    const int32_t max_value = t.cdf_length[row_i] - 2;
    const int32_t zero = -t.offset[row_i];    // table entry of the value 0
    if (zero >= 0 && zero < max_value && dec.try_symbol(row, zero, kPrecision)) {
      out[i] = 0;
      continue;
    }
    long long value = dec.decode(row, max_value + 1, kPrecision, t.lut ? t.lut + (long long)row_i * 256 : nullptr);
    if (value == max_value) {
      int widths = 0;
      for (;;) {
        const uint32_t v = dec.decode_uniform(kOverflowWidth);
        widths += (int)v;
        if (v != max_overflow) break;
        if (widths > 64) return false;
      }
      if (widths > 16) return false;
      uint64_t overflow = 0;
      for (int j = 0; j < widths; ++j) overflow |= (uint64_t)dec.decode_uniform(kOverflowWidth) << (j * kOverflowWidth);
      value = (long long)(overflow >> 1);
      if (overflow & 1) value = -value - 1; else value += max_value;
    }
    out[i] = (int32_t)(value + t.offset[row_i]);
  }
  // The encoder strips every trailing zero byte, so reading (implicit zeros) past the end is legitimate however far it
  // goes; a truncated stream cannot be told apart from that and simply decodes to different symbols.
  return true;
}

// Two independent streams of equal length decoded in lockstep by one thread: the per-symbol dependency chain (divide ->
// table search -> range update -> renormalise) is latency-bound, so interleaving two chains nearly doubles the symbols per
// second of an out-of-order core.  Same results as two decode_stream calls.
bool decode_stream_x2(const Tables& t, const uint8_t* b0, const uint8_t* e0, const uint8_t* b1, const uint8_t* e1,
                      const int32_t* idx0, const int32_t* idx1, long long n, int32_t* out0, int32_t* out1) {
  Decoder d0(b0, e0), d1(b1, e1);
  const uint32_t max_overflow = (1u << kOverflowWidth) - 1;
  auto escape = [&](Decoder& dec, long long& value, int32_t max_value) -> bool {
    int widths = 0;
    for (;;) {
      const uint32_t v = dec.decode_uniform(kOverflowWidth);
      widths += (int)v;
      if (v != max_overflow) break;
      if (widths > 64) return false;
    }
    if (widths > 16) return false;
    uint64_t overflow = 0;
    for (int j = 0; j < widths; ++j) overflow |= (uint64_t)dec.decode_uniform(kOverflowWidth) << (j * kOverflowWidth);
    value = (long long)(overflow >> 1);
    if (overflow & 1) value = -value - 1; else value += max_value;
    return true;
  };
  for (long long i = 0; i < n; ++i) {
    const int r0 = t.index(idx0, i), r1 = t.index(idx1, i);
    if (r0 < 0 || r0 >= t.rows || r1 < 0 || r1 >= t.rows) return false;
    const int32_t* row0 = t.cdf + (long long)r0 * t.cdf_stride;
    const int32_t* row1 = t.cdf + (long long)r1 * t.cdf_stride;
    const int32_t m0 = t.cdf_length[r0] - 2, m1 = t.cdf_length[r1] - 2;
    const int32_t z0 = -t.offset[r0], z1 = -t.offset[r1];
    const bool hit0 = z0 >= 0 && z0 < m0 && d0.try_symbol(row0, z0, kPrecision);
    const bool hit1 = z1 >= 0 && z1 < m1 && d1.try_symbol(row1, z1, kPrecision);
    long long v0 = z0, v1 = z1;
    if (!hit0) v0 = d0.decode(row0, m0 + 1, kPrecision, t.lut ? t.lut + (long long)r0 * 256 : nullptr);
    if (!hit1) v1 = d1.decode(row1, m1 + 1, kPrecision, t.lut ? t.lut + (long long)r1 * 256 : nullptr);
    if (!hit0 && v0 == m0 && !escape(d0, v0, m0)) return false;
    if (!hit1 && v1 == m1 && !escape(d1, v1, m1)) return false;
    out0[i] = (int32_t)(v0 + t.offset[r0]);
    out1[i] = (int32_t)(v1 + t.offset[r1]);
  }
  return true;
}

// Fork-join helpers of ONE calling thread (thread_local): created on first use, parked on a condition variable between
// calls.  Spawning std::threads per call cost 0.3-0.7 ms for 8-16 threads -- as much as the work itself for the small host
// stages of a batch (first latent's decode, coordinate packing, point extraction), six of which run per batch.  Each caller
// (the block loops' driver and each of its host workers) owns its team, so concurrent callers never queue behind each
// other's jobs; the caller takes part in the loop itself and then waits for its helpers.
class HelperTeam {
 public:
  ~HelperTeam() { shutdown(); }

  void run(int n, int threads, const std::function<void(int)>& f) {
    if (pid_ != getpid()) abandon();   // forked child: the parent's helper threads do not exist here
    {
      std::unique_lock<std::mutex> lk(m_);
      while ((int)helpers_.size() < threads - 1) {
        const int id = (int)helpers_.size();
        helpers_.emplace_back([this, id, seen = gen_] { loop(id, seen); });
      }
      f_ = &f;
      n_ = n;
      next_.store(0, std::memory_order_relaxed);
      active_ = threads - 1;
      done_ = 0;
      ++gen_;
    }
    cv_start_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return done_ == active_; });
    f_ = nullptr;
  }

 private:
  void work() {
    for (;;) {
      const int i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_) break;
      (*f_)(i);
    }
  }
  void loop(int id, unsigned long long seen) {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_start_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (id >= active_) continue;   // this job wants fewer helpers
      }
      work();
      std::unique_lock<std::mutex> lk(m_);
      if (++done_ == active_) cv_done_.notify_one();
    }
  }
  void shutdown() {
    if (pid_ != getpid()) {
      abandon();
      return;
    }
    {
      std::unique_lock<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_start_.notify_all();
    for (auto& t : helpers_) t.join();
    helpers_.clear();
  }
  void abandon() {
    // the std::thread objects name threads of the parent process: leak them (neither join nor detach may touch them)
    new std::vector<std::thread>(std::move(helpers_));
    helpers_.clear();
    gen_ = 0;
    stop_ = false;
    pid_ = getpid();
  }

  std::mutex m_;
  std::condition_variable cv_start_, cv_done_;
  std::vector<std::thread> helpers_;
  const std::function<void(int)>* f_ = nullptr;
  int n_ = 0, active_ = 0, done_ = 0;
  std::atomic<int> next_{0};
  unsigned long long gen_ = 0;
  bool stop_ = false;
  pid_t pid_ = getpid();
};

template <class F>
void parallel_for(int n, int threads, F&& f) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads > 256) threads = 256;
  if (threads <= 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  static thread_local HelperTeam team;
  const std::function<void(int)> fn = [&f](int i) { f(i); };
  team.run(n, threads, fn);
}

}  // namespace

extern "C" int pccgeo_range_encode_host(const int32_t* symbols, const int32_t* indexes, const long long* sym_offsets,
                                        int nstreams, const int32_t* cdf, int cdf_stride, const int32_t* cdf_length,
                                        const int32_t* offset, int rows, int index_mode, long long channel_stride,
                                        uint8_t* out_bytes, long long out_capacity, long long* out_offsets, int threads) {
  if (!symbols || !sym_offsets || !cdf || !cdf_length || !offset || !out_offsets || nstreams < 0 || rows <= 0 ||
      (index_mode == 0 && !indexes) || (index_mode == 1 && channel_stride <= 0) || (index_mode != 0 && index_mode != 1)) {
    pccgeo::set_error("range_encode: bad argument");
    return PCCGEO_EINVAL;
  }
  Tables t{cdf, cdf_stride, cdf_length, offset, rows, index_mode, channel_stride};
  std::vector<std::vector<uint8_t>> outs(nstreams);
  std::atomic<int> bad{0};
  parallel_for(nstreams, threads, [&](int i) {
    const long long a = sym_offsets[i], b = sym_offsets[i + 1];
    Encoder enc((size_t)(b - a));  // coder state on this thread's stack: neighbouring streams' states must not share cache lines
    if (!encode_stream(t, symbols + a, index_mode == 0 ? indexes + a : nullptr, b - a, enc)) bad.store(1);
    outs[i] = std::move(enc.out);
  });
  if (bad.load()) {
    pccgeo::set_error("range_encode: table index out of range");
    return PCCGEO_EINVAL;
  }
  long long pos = 0;
  out_offsets[0] = 0;
  for (int i = 0; i < nstreams; ++i) {
    pos += (long long)outs[i].size();
    out_offsets[i + 1] = pos;
  }
  if (!out_bytes || pos > out_capacity) {
    pccgeo::set_error("range_encode: output needs %lld bytes, capacity %lld", pos, out_capacity);
    return PCCGEO_ENOSPC;
  }
  for (int i = 0; i < nstreams; ++i)
    if (!outs[i].empty()) std::memcpy(out_bytes + out_offsets[i], outs[i].data(), outs[i].size());
  return PCCGEO_OK;
}

extern "C" int pccgeo_range_decode_host(const uint8_t* bytes, const long long* byte_offsets, const int32_t* indexes,
                                        const long long* sym_offsets, int nstreams, const int32_t* cdf, int cdf_stride,
                                        const int32_t* cdf_length, const int32_t* offset, int rows, int index_mode,
                                        long long channel_stride, int32_t* symbols_out, int threads) {
  if (!byte_offsets || !sym_offsets || !cdf || !cdf_length || !offset || !symbols_out || nstreams < 0 || rows <= 0 ||
      (index_mode == 0 && !indexes) || (index_mode == 1 && channel_stride <= 0) || (index_mode != 0 && index_mode != 1)) {
    pccgeo::set_error("range_decode: bad argument");
    return PCCGEO_EINVAL;
  }
  static const uint8_t kEmpty = 0;
  const uint8_t* base = bytes ? bytes : &kEmpty;
  Tables t{cdf, cdf_stride, cdf_length, offset, rows, index_mode, channel_stride};
  // per-row search accelerators: lut[r][b] = largest s in [0, n) with cdf[s] <= b << 8
  std::vector<uint16_t> lut((size_t)rows * 256);
  for (int r = 0; r < rows; ++r) {
    const int32_t* row = cdf + (long long)r * cdf_stride;
    const int n = cdf_length[r] - 1;  // symbols incl. the escape slot
    int s = 0;
    for (int b = 0; b < 256; ++b) {
      const int32_t v = b << (kPrecision - 8);
      while (s + 1 < n && row[s + 1] <= v) ++s;
      lut[(size_t)r * 256 + b] = (uint16_t)s;
    }
  }
  t.lut = lut.data();
  std::atomic<int> bad{0};
  // streams are decoded in pairs (two interleaved dependency chains per thread) when neighbours have equal length
  const int npairs = (nstreams + 1) / 2;
  parallel_for(npairs, threads, [&](int pi) {
    const int i = 2 * pi, j = i + 1;
    const long long a = sym_offsets[i], b = sym_offsets[i + 1];
    const int32_t* ia = index_mode == 0 ? indexes + a : nullptr;
    if (j < nstreams && sym_offsets[j + 1] - sym_offsets[j] == b - a && index_mode == 0) {
      const long long c = sym_offsets[j];
      if (!decode_stream_x2(t, base + byte_offsets[i], base + byte_offsets[i + 1], base + byte_offsets[j], base + byte_offsets[j + 1],
                            ia, indexes + c, b - a, symbols_out + a, symbols_out + c))
        bad.store(1);
      return;
    }
    if (!decode_stream(t, base + byte_offsets[i], base + byte_offsets[i + 1], ia, b - a, symbols_out + a)) bad.store(1);
    if (j < nstreams) {
      const long long c = sym_offsets[j], d = sym_offsets[j + 1];
      if (!decode_stream(t, base + byte_offsets[j], base + byte_offsets[j + 1], index_mode == 0 ? indexes + c : nullptr, d - c,
                         symbols_out + c))
        bad.store(1);
    }
  });
  if (bad.load()) {
    pccgeo::set_error("range_decode: corrupt stream");
    return PCCGEO_EDATA;
  }
  return PCCGEO_OK;
}

extern "C" int pccgeo_pmf_to_quantized_cdf_host(const double* pmf, int len, int precision, int32_t* cdf) {
  if (!pmf || !cdf || len <= 0 || precision < 1 || precision > 24 || len > (1 << precision)) {
    pccgeo::set_error("pmf_to_quantized_cdf: bad argument");
    return PCCGEO_EINVAL;
  }
  const long long target = 1ll << precision;
  std::vector<long long> v(len);
  long long total = 0;
  for (int i = 0; i < len; ++i) {
    long long q = (long long)std::nearbyint(pmf[i] * (double)target);
    v[i] = q < 1 ? 1 : q;
    total += v[i];
  }
  typedef std::pair<double, int> Item;
  std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
  if (total > target) {
    for (int i = 0; i < len; ++i)
      if (v[i] > 1) heap.push(Item(pmf[i] * (std::log2((double)v[i]) - std::log2((double)(v[i] - 1))), i));
    while (total > target) {
      if (heap.empty()) {
        pccgeo::set_error("pmf_to_quantized_cdf: cannot repair the sum");
        return PCCGEO_EINVAL;
      }
      const int i = heap.top().second;
      heap.pop();
      --v[i];
      --total;
      if (v[i] > 1) heap.push(Item(pmf[i] * (std::log2((double)v[i]) - std::log2((double)(v[i] - 1))), i));
    }
  } else if (total < target) {
    for (int i = 0; i < len; ++i) heap.push(Item(-pmf[i] * (std::log2((double)(v[i] + 1)) - std::log2((double)v[i])), i));
    while (total < target) {
      const int i = heap.top().second;
      heap.pop();
      ++v[i];
      ++total;
      heap.push(Item(-pmf[i] * (std::log2((double)(v[i] + 1)) - std::log2((double)v[i])), i));
    }
  }
  cdf[0] = 0;
  long long acc = 0;
  for (int i = 0; i < len; ++i) {
    acc += v[i];
    cdf[i + 1] = (int32_t)acc;
  }
  return PCCGEO_OK;
}

// Packed occupancy words (one block = d*h*w/32 uint32, bit i of word w <-> voxel w*32+i in C order) -> float32 (z,y,x)
// rows in np.argwhere order, the host half of decompress_blocks / compress_blocks (reference src/model_types.py:209,234).
// offsets[b] (prefix sum of the per-block popcounts, computed here when counts == NULL) gives each block's first row.
extern "C" int pccgeo_bits_to_points_host(const uint32_t* bits, int n_blocks, int d, int h, int w, long long* offsets,
                                          float* points, long long capacity_points, int threads) {
  if (!bits || !offsets || n_blocks < 0 || d <= 0 || h <= 0 || w <= 0 || ((long long)d * h * w) % 32 != 0) {
    pccgeo::set_error("bits_to_points: bad argument");
    return PCCGEO_EINVAL;
  }
  const long long words = (long long)d * h * w / 32;
  std::vector<long long> cnt(n_blocks, 0);
  parallel_for(n_blocks, threads, [&](int b) {
    long long c = 0;
    const uint32_t* p = bits + (long long)b * words;
    for (long long i = 0; i < words; ++i) c += __builtin_popcount(p[i]);
    cnt[b] = c;
  });
  offsets[0] = 0;
  for (int b = 0; b < n_blocks; ++b) offsets[b + 1] = offsets[b] + cnt[b];
  if (!points) return PCCGEO_OK;  // size query
  if (offsets[n_blocks] > capacity_points) {
    pccgeo::set_error("bits_to_points: need %lld rows, capacity %lld", offsets[n_blocks], capacity_points);
    return PCCGEO_ENOSPC;
  }
  const int hw = h * w;
  parallel_for(n_blocks, threads, [&](int b) {
    const uint32_t* p = bits + (long long)b * words;
    float* o = points + offsets[b] * 3;
    for (long long i = 0; i < words; ++i) {
      uint32_t m = p[i];
      while (m) {
        const int bit = __builtin_ctz(m);
        m &= m - 1;
        const long long v = i * 32 + bit;
        const int z = (int)(v / hw), rem = (int)(v % hw);
        o[0] = (float)z; o[1] = (float)(rem / w); o[2] = (float)(rem % w);
        o += 3;
      }
    }
  });
  return PCCGEO_OK;
}

// Host half of sparse_to_dense / pc_to_tf (reference src/model_types.py:23-39,108-114) for a batch: per-block point
// arrays (rows of >= 3 float32 / float64 coordinates, any row pitch) -> one int16 (sum n_i, 4) array of
// (block, c0, c1, c2) rows for pccgeo_densify.  offsets[b] = first output row of block b (prefix sum of counts).
extern "C" int pccgeo_blocks_to_coords_host(const void* const* blocks, const long long* counts, const long long* row_bytes,
                                            int n_blocks, int is_f64, int16_t* out, int threads) {
  if (n_blocks < 0 || (n_blocks > 0 && (!blocks || !counts || !row_bytes || !out))) {
    pccgeo::set_error("blocks_to_coords: bad argument");
    return PCCGEO_EINVAL;
  }
  std::vector<long long> offs(n_blocks + 1, 0);
  for (int b = 0; b < n_blocks; ++b) {
    if (counts[b] < 0 || (counts[b] > 0 && !blocks[b])) {
      pccgeo::set_error("blocks_to_coords: block %d is null", b);
      return PCCGEO_EINVAL;
    }
    offs[b + 1] = offs[b] + counts[b];
  }
  parallel_for(n_blocks, threads, [&](int b) {
    const uint8_t* src = (const uint8_t*)blocks[b];
    int16_t* o = out + offs[b] * 4;
    for (long long i = 0; i < counts[b]; ++i, src += row_bytes[b], o += 4) {
      o[0] = (int16_t)b;
      if (is_f64) {
        const double* r = (const double*)src;
        o[1] = (int16_t)r[0]; o[2] = (int16_t)r[1]; o[3] = (int16_t)r[2];
      } else {
        const float* r = (const float*)src;
        o[1] = (int16_t)r[0]; o[2] = (int16_t)r[1]; o[3] = (int16_t)r[2];
      }
    }
  });
  return PCCGEO_OK;
}

// Host half of partition_octree (reference src/utils/octree_coding.py:103-111, the per-point Python loop that takes 7.6 s
// on longdress): stable counting sort of point rows by block index.  rows: (n, cols) float64; block_idx[i] in [0, n_blocks);
// origins: (n_blocks, 3) block origins subtracted from the first three columns; out: (n, cols) grouped rows, points of a
// block keep their input order; offsets: (n_blocks + 1) prefix sums.
extern "C" int pccgeo_group_points_host(const double* rows, const int32_t* block_idx, long long n, int cols, int n_blocks,
                                        const double* origins, double* out, long long* offsets) {
  if (n < 0 || cols < 3 || n_blocks < 0 || (n > 0 && (!rows || !block_idx || !origins || !out)) || !offsets) {
    pccgeo::set_error("group_points: bad argument");
    return PCCGEO_EINVAL;
  }
  for (int b = 0; b <= n_blocks; ++b) offsets[b] = 0;
  for (long long i = 0; i < n; ++i) {
    if (block_idx[i] < 0 || block_idx[i] >= n_blocks) {
      pccgeo::set_error("group_points: block index out of range");
      return PCCGEO_EINVAL;
    }
    ++offsets[block_idx[i] + 1];
  }
  for (int b = 0; b < n_blocks; ++b) offsets[b + 1] += offsets[b];
  std::vector<long long> cur(offsets, offsets + n_blocks);
  for (long long i = 0; i < n; ++i) {
    const int b = block_idx[i];
    double* o = out + cur[b]++ * cols;
    const double* r = rows + i * cols;
    o[0] = r[0] - origins[b * 3]; o[1] = r[1] - origins[b * 3 + 1]; o[2] = r[2] - origins[b * 3 + 2];
    for (int c = 3; c < cols; ++c) o[c] = r[c];
  }
  return PCCGEO_OK;
}
