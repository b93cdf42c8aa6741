"""Oracle (test infrastructure): pure-Python restatement of the range-coding ops the reference reaches
through tensorflow-compression 1.3 (`unbounded_index_range_encode/decode`, precision=16,
overflow_width=4; call sites src/utils/patch_gaussian_conditional.py:27-31 and tfc's
EntropyBottleneck/GaussianConditional .compress/.decompress used at src/model_types.py:291-292,382-387,
404-407).

PARITY UNPINNED: tfc's C++ kernels are not available offline.  The *symbol-level* contract is restated
(value = symbol - offset[index]; in-range values coded with the row's 16-bit CDF; out-of-range values
coded as the escape symbol cdf_length-2 followed by an Elias-gamma-like code in 4-bit uniform chunks,
negatives on odd codes).  The *byte-level* arithmetic coder is this project's own 32-bit range coder
with carry propagation (documented in DESIGN.md) -- it is NOT claimed byte-compatible with tfc.
The product's C++/CUDA coders must match THIS file byte for byte.

Slow (pure Python): use on small inputs only.
"""
import numpy as np

TOP = 1 << 24
MASK32 = 0xFFFFFFFF


class RangeEncoder:
    def __init__(self):
        self.low = 0            # up to 33 bits
        self.range = MASK32
        self.cache = 0
        self.cache_size = 1
        self.out = bytearray()

    def _shift_low(self):
        if (self.low & MASK32) < 0xFF000000 or (self.low >> 32) != 0:
            carry = self.low >> 32
            temp = self.cache
            while True:
                self.out.append((temp + carry) & 0xFF)
                temp = 0xFF
                self.cache_size -= 1
                if self.cache_size == 0:
                    break
            self.cache = (self.low >> 24) & 0xFF
        self.cache_size += 1
        self.low = (self.low & 0x00FFFFFF) << 8

    def encode(self, lower, upper, precision):
        r = self.range >> precision
        self.low += r * lower
        self.range = r * (upper - lower)
        while self.range < TOP:
            self._shift_low()
            self.range = (self.range << 8) & MASK32

    def finish(self):
        # pick the value in [low, low+range) with the most trailing zero bits, then flush and strip the
        # trailing zero bytes (the decoder pads with zeros); the first emitted byte is always 0: drop it.
        hi = self.low + self.range - 1
        for nbits in range(32, -1, -1):
            mask = (1 << nbits) - 1
            v = (self.low + mask) & ~mask
            if v <= hi:
                self.low = v
                break
        for _ in range(5):
            self._shift_low()
        out = bytes(self.out[1:])
        return out.rstrip(b'\x00')


class RangeDecoder:
    def __init__(self, data):
        self.data = data
        self.pos = 0
        self.range = MASK32
        self.code = 0
        for _ in range(4):
            self.code = (self.code << 8) | self._next()

    def _next(self):
        b = self.data[self.pos] if self.pos < len(self.data) else 0
        self.pos += 1
        return b

    def decode(self, cdf, n, precision):
        """cdf: sequence with cdf[0]=0 .. cdf[n]=2**precision; returns symbol in [0,n)."""
        r = self.range >> precision
        value = min(self.code // r, (1 << precision) - 1)
        lo, hi = 0, n  # largest s with cdf[s] <= value
        while hi - lo > 1:
            mid = (lo + hi) >> 1
            if cdf[mid] <= value:
                lo = mid
            else:
                hi = mid
        s = lo
        self.code -= r * int(cdf[s])
        self.range = r * (int(cdf[s + 1]) - int(cdf[s]))
        while self.range < TOP:
            self.code = ((self.code << 8) | self._next()) & MASK32
            self.range = (self.range << 8) & MASK32
        return s

    def decode_uniform(self, bits):
        r = self.range >> bits
        s = min(self.code // r, (1 << bits) - 1)
        self.code -= r * s
        self.range = r
        while self.range < TOP:
            self.code = ((self.code << 8) | self._next()) & MASK32
            self.range = (self.range << 8) & MASK32
        return s


def unbounded_index_range_encode(symbols, indexes, cdf, cdf_length, offset, precision=16, overflow_width=4):
    symbols = np.asarray(symbols).reshape(-1)
    indexes = np.asarray(indexes).reshape(-1)
    enc = RangeEncoder()
    max_overflow = (1 << overflow_width) - 1
    for sym, idx in zip(symbols.tolist(), indexes.tolist()):
        row = cdf[idx]
        max_value = int(cdf_length[idx]) - 2
        value = sym - int(offset[idx])
        overflow = 0
        if value < 0:
            overflow = -2 * value - 1
            value = max_value
        elif value >= max_value:
            overflow = 2 * (value - max_value)
            value = max_value
        enc.encode(int(row[value]), int(row[value + 1]), precision)
        if value == max_value:
            widths = 0
            while (overflow >> (widths * overflow_width)) != 0:
                widths += 1
            val = widths
            while val >= max_overflow:
                enc.encode(max_overflow, max_overflow + 1, overflow_width)
                val -= max_overflow
            enc.encode(val, val + 1, overflow_width)
            for j in range(widths):
                val = (overflow >> (j * overflow_width)) & max_overflow
                enc.encode(val, val + 1, overflow_width)
    return enc.finish()


def unbounded_index_range_decode(data, indexes, cdf, cdf_length, offset, precision=16, overflow_width=4):
    indexes = np.asarray(indexes)
    flat = indexes.reshape(-1)
    dec = RangeDecoder(data)
    out = np.zeros(flat.shape, np.int32)
    max_overflow = (1 << overflow_width) - 1
    for i, idx in enumerate(flat.tolist()):
        row = cdf[idx]
        max_value = int(cdf_length[idx]) - 2
        value = dec.decode(row, max_value + 1, precision)
        if value == max_value:
            widths = 0
            while True:
                val = dec.decode_uniform(overflow_width)
                widths += val
                if val != max_overflow:
                    break
            overflow = 0
            for j in range(widths):
                overflow |= dec.decode_uniform(overflow_width) << (j * overflow_width)
            value = overflow >> 1
            if overflow & 1:
                value = -value - 1
            else:
                value += max_value
        out[i] = value + int(offset[idx])
    return out.reshape(indexes.shape)


# ---------------------------------------------------------------------------------------------------------
# the same coder in plain C (oracle/range_coder_c.c, built by oracle/build.py): lets the CPU baseline of bench.py and the
# tests code full-size blocks in milliseconds WITHOUT loading the product library.  Pinned against the Python statements
# above byte for byte (tests/test_range_coder.py).
# ---------------------------------------------------------------------------------------------------------
_clib = [None]


def _c():
    if _clib[0] is None:
        import ctypes
        from . import build as B
        lib = ctypes.CDLL(B.build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.oracle_rc_encode.restype = i64
        lib.oracle_rc_encode.argtypes = [p, p, i64, p, i64, p, p, i64, p, i64]
        lib.oracle_rc_decode.restype = ctypes.c_int
        lib.oracle_rc_decode.argtypes = [p, i64, p, i64, p, i64, p, p, i64, p]
        _clib[0] = lib
    return _clib[0]


def _tab(cdf, cdf_length, offset):
    cdf = np.ascontiguousarray(cdf, np.int32)
    return cdf, np.ascontiguousarray(cdf_length, np.int32), np.ascontiguousarray(offset, np.int32)


def encode_c(symbols, indexes, cdf, cdf_length, offset):
    """unbounded_index_range_encode through the C restatement -> bytes."""
    sym = np.ascontiguousarray(np.asarray(symbols).reshape(-1), np.int32)
    idx = np.ascontiguousarray(np.asarray(indexes).reshape(-1), np.int32)
    assert sym.size == idx.size
    cdf, cl, off = _tab(cdf, cdf_length, offset)
    cap = 2 * sym.size + 64
    while True:
        out = np.empty(cap, np.uint8)
        n = _c().oracle_rc_encode(sym.ctypes.data, idx.ctypes.data, sym.size, cdf.ctypes.data, cdf.shape[1], cl.ctypes.data,
                                  off.ctypes.data, cdf.shape[0], out.ctypes.data, cap)
        if n < 0:
            raise ValueError('table index out of range')
        if n <= cap:
            return out[:n].tobytes()
        cap = int(n) + 8


def decode_c(data, indexes, cdf, cdf_length, offset):
    """unbounded_index_range_decode through the C restatement -> int32 array shaped like indexes."""
    indexes = np.asarray(indexes)
    idx = np.ascontiguousarray(indexes.reshape(-1), np.int32)
    cdf, cl, off = _tab(cdf, cdf_length, offset)
    buf = np.frombuffer(bytes(data) + b'\x00', np.uint8)
    out = np.empty(idx.size, np.int32)
    rc = _c().oracle_rc_decode(buf.ctypes.data, len(data), idx.ctypes.data, idx.size, cdf.ctypes.data, cdf.shape[1], cl.ctypes.data,
                               off.ctypes.data, cdf.shape[0], out.ctypes.data)
    if rc != 0:
        raise ValueError(f'oracle_rc_decode failed ({rc})')
    return out.reshape(indexes.shape)
