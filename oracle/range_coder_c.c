/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of oracle/range_coder.py, statement for statement -- the
 * symbol-level contract of tensorflow-compression 1.3 `unbounded_index_range_encode/decode` (precision 16, overflow_width 4;
 * reference call sites src/utils/patch_gaussian_conditional.py:27-31, src/model_types.py:291-292,382-387,404-407) over this
 * project's byte-level 32-bit range coder (low / cache / cache_size carry propagation).  PARITY UNPINNED against tfc's own
 * bytes (its kernels are not available offline); pinned against oracle/range_coder.py byte for byte (tests/test_range_coder.py).
 *
 * Exists so that the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) and the tests have an entropy coder that is
 * independent of the product library: nothing here is shared with pcc_geo_cnn_v2_b200/csrc.  Built by oracle/build.py into
 * oracle/_build/liboracle_rc.so (gcc -O2); the product never loads it.
 */
#include <stdint.h>
#include <stddef.h>

#define TOP (1u << 24)

typedef struct {
    uint64_t low;       /* up to 33 bits */
    uint32_t range;
    uint8_t cache;
    int64_t cache_size;
    uint8_t *out;
    int64_t n, cap;     /* n counts every byte, also those beyond cap (size query) */
} enc_t;

static void put(enc_t *e, uint8_t b) {   /* byte -1 is the always-zero first byte of the machine: dropped */
    if (e->n >= 0 && e->n < e->cap) e->out[e->n] = b;
    e->n++;
}

static void shift_low(enc_t *e) {
    if ((uint32_t)e->low < 0xFF000000u || (e->low >> 32) != 0) {
        uint8_t carry = (uint8_t)(e->low >> 32);
        uint8_t temp = e->cache;
        do {
            put(e, (uint8_t)(temp + carry));
            temp = 0xFF;
        } while (--e->cache_size != 0);
        e->cache = (uint8_t)(e->low >> 24);
    }
    e->cache_size++;
    e->low = (e->low & 0x00FFFFFFu) << 8;
}

static void encode(enc_t *e, uint32_t lower, uint32_t upper, int precision) {
    uint32_t r = e->range >> precision;
    e->low += (uint64_t)r * lower;
    e->range = r * (upper - lower);
    while (e->range < TOP) {
        shift_low(e);
        e->range <<= 8;
    }
}

/* One stream.  Returns the number of bytes of the string (after dropping the always-zero first byte and the trailing zero
 * bytes); if it exceeds cap only the first cap bytes were written.  -1: a table index is out of range. */
int64_t oracle_rc_encode(const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *cdf, int64_t cdf_stride,
                         const int32_t *cdf_length, const int32_t *offset, int64_t n_rows, uint8_t *out, int64_t cap) {
    const int precision = 16, overflow_width = 4;
    const uint32_t max_overflow = (1u << overflow_width) - 1;
    enc_t e = {0, 0xFFFFFFFFu, 0, 1, out, -1, cap};
    for (int64_t i = 0; i < n; i++) {
        int32_t idx = indexes[i];
        if (idx < 0 || idx >= n_rows) return -1;
        const int32_t *row = cdf + idx * cdf_stride;
        int64_t max_value = (int64_t)cdf_length[idx] - 2;
        int64_t value = (int64_t)symbols[i] - offset[idx];
        uint64_t overflow = 0;
        if (value < 0) {
            overflow = (uint64_t)(-2 * value - 1);
            value = max_value;
        } else if (value >= max_value) {
            overflow = (uint64_t)(2 * (value - max_value));
            value = max_value;
        }
        encode(&e, (uint32_t)row[value], (uint32_t)row[value + 1], precision);
        if (value == max_value) {
            int widths = 0;
            while (widths < 16 && (overflow >> (widths * overflow_width)) != 0) widths++;
            uint32_t val = (uint32_t)widths;
            while (val >= max_overflow) {
                encode(&e, max_overflow, max_overflow + 1, overflow_width);
                val -= max_overflow;
            }
            encode(&e, val, val + 1, overflow_width);
            for (int j = 0; j < widths; j++) {
                val = (uint32_t)((overflow >> (j * overflow_width)) & max_overflow);
                encode(&e, val, val + 1, overflow_width);
            }
        }
    }
    /* finish: the value in [low, low+range) with the most trailing zero bits, flush, strip trailing zeros */
    uint64_t hi = e.low + e.range - 1;
    for (int nbits = 32; nbits >= 0; nbits--) {
        uint64_t mask = (nbits >= 64) ? ~0ull : ((1ull << nbits) - 1);
        uint64_t v = (e.low + mask) & ~mask;
        if (v <= hi) {
            e.low = v;
            break;
        }
    }
    for (int k = 0; k < 5; k++) shift_low(&e);
    int64_t len = e.n < 0 ? 0 : e.n;
    /* trailing zero bytes: only those inside the written part can be inspected; the caller re-calls with a larger buffer
     * when len > cap, so stripping here is exact whenever the string fits */
    if (len <= cap)
        while (len > 0 && out[len - 1] == 0) len--;
    return len;
}

typedef struct {
    const uint8_t *data;
    int64_t pos, len;
    uint32_t range, code;
} dec_t;

static uint32_t next_byte(dec_t *d) {
    uint32_t b = d->pos < d->len ? d->data[d->pos] : 0;
    d->pos++;
    return b;
}

static void renorm(dec_t *d) {
    while (d->range < TOP) {
        d->code = (d->code << 8) | next_byte(d);
        d->range <<= 8;
    }
}

static uint32_t decode_uniform(dec_t *d, int bits) {
    uint32_t r = d->range >> bits;
    uint32_t s = d->code / r;
    if (s > (1u << bits) - 1) s = (1u << bits) - 1;
    d->code -= r * s;
    d->range = r;
    renorm(d);
    return s;
}

/* One stream.  0 on success, -1: table index out of range, -2: malformed escape code. */
int oracle_rc_decode(const uint8_t *data, int64_t nbytes, const int32_t *indexes, int64_t n, const int32_t *cdf, int64_t cdf_stride,
                     const int32_t *cdf_length, const int32_t *offset, int64_t n_rows, int32_t *out) {
    const int precision = 16, overflow_width = 4;
    const uint32_t max_overflow = (1u << overflow_width) - 1;
    dec_t d = {data, 0, nbytes, 0xFFFFFFFFu, 0};
    for (int k = 0; k < 4; k++) d.code = (d.code << 8) | next_byte(&d);
    for (int64_t i = 0; i < n; i++) {
        int32_t idx = indexes[i];
        if (idx < 0 || idx >= n_rows) return -1;
        const int32_t *row = cdf + idx * cdf_stride;
        int64_t max_value = (int64_t)cdf_length[idx] - 2;
        uint32_t r = d.range >> precision;
        uint32_t target = d.code / r;
        if (target > (1u << precision) - 1) target = (1u << precision) - 1;
        int64_t lo = 0, hi = max_value + 1;     /* largest s with cdf[s] <= target */
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if ((uint32_t)row[mid] <= target) lo = mid; else hi = mid;
        }
        int64_t value = lo;
        d.code -= r * (uint32_t)row[value];
        d.range = r * (uint32_t)(row[value + 1] - row[value]);
        renorm(&d);
        if (value == max_value) {
            int64_t widths = 0;
            for (;;) {
                uint32_t val = decode_uniform(&d, overflow_width);
                widths += val;
                if (val != max_overflow) break;
                if (widths > 64) return -2;
            }
            if (widths > 16) return -2;
            uint64_t overflow = 0;
            for (int64_t j = 0; j < widths; j++) overflow |= (uint64_t)decode_uniform(&d, overflow_width) << (j * overflow_width);
            value = (int64_t)(overflow >> 1);
            if (overflow & 1) value = -value - 1; else value += max_value;
        }
        out[i] = (int32_t)(value + offset[idx]);
    }
    return 0;
}
