"""Range coder: the pure-Python oracle round-trips, and the C++ host coder in libpccgeo matches it byte for byte
(same bytes out of encode, same symbols out of decode), including escape-coded overflow, empty streams,
per-channel (EntropyBottleneck) index mode and corrupt input."""
import numpy as np
import pytest

from oracle import entropy as E
from oracle import range_coder as RC
from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200._lib import PccGeoError


@pytest.fixture(scope='module')
def gc_tab():
    return E.gc_tables(E.make_scale_table())


def _random_stream(rng, tab, n, spread):
    idx = rng.integers(0, len(tab['cdf_length']), size=n).astype(np.int32)
    center = -tab['offset'][idx]
    sym = np.rint(rng.normal(size=n) * (center / 2.5) * spread).astype(np.int32)
    return sym, idx


@pytest.mark.parametrize('n,spread', [(0, 1.0), (1, 1.0), (257, 1.0), (2000, 3.0), (500, 40.0)])
def test_python_oracle_round_trip(gc_tab, n, spread):
    rng = np.random.default_rng(n)
    sym, idx = _random_stream(rng, gc_tab, n, spread)
    data = RC.unbounded_index_range_encode(sym, idx, gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
    out = RC.unbounded_index_range_decode(data, idx, gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
    assert np.array_equal(out, sym)


def test_known_answer_bytes():
    """Hand-checkable: a 2-symbol alphabet with p=(1/2,1/2) + escape slot; coding symbol 0 four times narrows
    [0,2^32) to [0, 2^32/16) -> the shortest representative is the empty string."""
    cdf = np.array([[0, 32768, 65535, 65536]], np.int32)
    cl, off = np.array([4], np.int32), np.array([0], np.int32)
    assert RC.unbounded_index_range_encode([0, 0, 0, 0], [0] * 4, cdf, cl, off) == b''
    s = RC.unbounded_index_range_encode([1, 1, 1, 1], [0] * 4, cdf, cl, off)
    assert RC.unbounded_index_range_decode(s, np.zeros(4, np.int32), cdf, cl, off).tolist() == [1, 1, 1, 1]
    assert len(s) <= 2


def test_cpp_matches_python_bytes(gc_tab):
    rng = np.random.default_rng(7)
    lens = [0, 1, 33, 1000, 4096, 5]
    syms, idxs = zip(*[_random_stream(rng, gc_tab, n, s) for n, s in zip(lens, [1, 1, 1, 2.0, 1.0, 60.0])])
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    got = ops.range_encode(np.concatenate(syms), offs, gc_tab, indexes=np.concatenate(idxs), threads=3)
    want = [RC.unbounded_index_range_encode(s, i, gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
            for s, i in zip(syms, idxs)]
    assert got == want
    dec = ops.range_decode(got, offs, gc_tab, indexes=np.concatenate(idxs), threads=2)
    assert np.array_equal(dec, np.concatenate(syms))


def test_cpp_channel_index_mode_matches_python():
    p = E.eb_init(4, np.random.default_rng(0))
    tab = E.eb_tables(p)
    rng = np.random.default_rng(1)
    sym = np.rint(rng.normal(size=(3, 4, 2, 2, 2)) * 8).astype(np.int32)   # |sym| > 10 -> escapes
    offs = np.arange(4, dtype=np.int64) * 32
    got = ops.range_encode(sym.reshape(-1), offs, tab, channel_stride=8)
    chan = np.broadcast_to(np.arange(4, dtype=np.int32).reshape(4, 1, 1, 1), (4, 2, 2, 2))
    want = [RC.unbounded_index_range_encode(sym[i], chan, tab['cdf'], tab['cdf_length'], tab['offset']) for i in range(3)]
    assert got == want
    dec = ops.range_decode(got, offs, tab, channel_stride=8)
    assert np.array_equal(dec.reshape(sym.shape), sym)


def test_cpp_rejects_bad_index_and_truncation(gc_tab):
    offs = np.array([0, 4], np.int64)
    with pytest.raises(PccGeoError):
        ops.range_encode(np.zeros(4, np.int32), offs, gc_tab, indexes=np.full(4, 64, np.int32))
    rng = np.random.default_rng(3)
    sym, idx = _random_stream(rng, gc_tab, 4000, 1.0)
    offs = np.array([0, 4000], np.int64)
    s = ops.range_encode(sym, offs, gc_tab, indexes=idx)[0]
    # a truncated string either decodes to different symbols or is flagged -- it must never crash
    try:
        out = ops.range_decode([s[:len(s) // 4]], offs, gc_tab, indexes=idx)
        assert not np.array_equal(out, sym)
    except PccGeoError:
        pass


def test_pmf_to_quantized_cdf_cpp_matches_oracle():
    rng = np.random.default_rng(5)
    for n in (2, 3, 24, 300, 1480):
        p = rng.random(n) ** 6
        p /= p.sum() * rng.uniform(0.9, 1.1)
        assert np.array_equal(ops.pmf_to_quantized_cdf(p), E.pmf_to_quantized_cdf(p))


def test_long_tail_of_near_certain_symbols_round_trips(gc_tab):
    """Regression: (a) thousands of ~probability-1 symbols cost nothing; (b) a tail of symbols whose CDF lower bound is 0
    (the leftmost bin of a narrow table) emits only zero bytes, which the encoder strips -- the decoder must read implicit
    zeros far past the end of the string (it once flagged more than 8 such bytes as a corrupt stream)."""
    rng = np.random.default_rng(11)
    head = 300
    sym = np.zeros(30000, np.int32)
    idx = np.zeros(30000, np.int32)                       # scale index 0 (sigma 0.11): p(0) ~ 1
    idx[:head] = rng.integers(20, 40, size=head)
    sym[:head] = np.rint(rng.normal(size=head) * 20).astype(np.int32)
    sym[-60:] = gc_tab['offset'][0]                       # value 0 of table 0: lower bound 0, ~2 zero bytes each
    offs = np.array([0, len(sym)], np.int64)
    s = ops.range_encode(sym, offs, gc_tab, indexes=idx)
    assert len(s[0]) < 1500                               # neither the certain symbols nor the zero-byte tail add length
    assert np.array_equal(ops.range_decode(s, offs, gc_tab, indexes=idx), sym)
    small = slice(0, 3000)
    want = RC.unbounded_index_range_encode(sym[small], idx[small], gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
    got = ops.range_encode(sym[small], np.array([0, 3000], np.int64), gc_tab, indexes=idx[small])[0]
    assert got == want
    assert np.array_equal(RC.unbounded_index_range_decode(want, idx[small], gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset']),
                          sym[small])


def test_random_tables_property():
    """Property test over random 16-bit tables (random pmfs through pmf_to_quantized_cdf, 2..300 symbols per row, random
    offsets), random streams incl. out-of-table symbols, both index modes: C++ encode -> decode is the identity, the device
    encoder's arithmetic (run on the host) gives the same bytes, and the Python oracle agrees on a sample."""
    from hypothesis import given, settings, strategies as st
    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1))
    def check(seed):
        rng = np.random.default_rng(seed)
        rows = int(rng.integers(1, 9))
        lens = rng.integers(2, 300, rows)
        maxlen = int(lens.max())
        cdf = np.zeros((rows, maxlen + 2), np.int32)
        for r in range(rows):
            pmf = rng.random(int(lens[r])) ** int(rng.integers(1, 6)) + 1e-9
            tail = rng.random() * 0.05 + 1e-6
            p = np.concatenate([pmf / pmf.sum() * (1 - tail), [tail]])
            cdf[r, :lens[r] + 2] = ops.pmf_to_quantized_cdf(p, 16)
        tab = {'cdf': cdf, 'cdf_length': (lens + 2).astype(np.int32), 'offset': rng.integers(-150, 5, rows).astype(np.int32)}
        ns, per = int(rng.integers(1, 5)), int(rng.integers(1, 400))
        idx = rng.integers(0, rows, (ns, per)).astype(np.int32)
        sym = (tab['offset'][idx] + rng.integers(-3, lens[idx] + 4)).astype(np.int32)
        sym[rng.random((ns, per)) < 0.02] += int(rng.integers(-40000, 40000))      # far escapes
        offs = np.arange(ns + 1, dtype=np.int64) * per
        strings = ops.range_encode(sym.reshape(-1), offs, tab, indexes=idx.reshape(-1), threads=2)
        assert np.array_equal(ops.range_decode(strings, offs, tab, indexes=idx.reshape(-1), threads=2).reshape(ns, per), sym)
        assert ops.range_encode_emulate(sym, ns, tab, indexes=idx) == strings
        assert RC.unbounded_index_range_encode(sym[0], idx[0], tab['cdf'], tab['cdf_length'], tab['offset']) == strings[0]
        if per % rows == 0:       # per-channel tables: index = (position / channel_stride) % rows
            cs = per // rows
            s2 = ops.range_encode(sym.reshape(-1), offs, tab, channel_stride=cs, threads=1)
            assert np.array_equal(ops.range_decode(s2, offs, tab, channel_stride=cs, threads=1).reshape(ns, per), sym)
            assert ops.range_encode_emulate(sym, ns, tab, channel_stride=cs) == s2

    check()


@pytest.mark.parametrize('n,spread', [(0, 1.0), (1, 1.0), (257, 1.0), (2000, 3.0), (500, 40.0), (3000, 1e6)])
def test_c_oracle_matches_python_oracle(gc_tab, n, spread):
    """oracle/range_coder_c.c (the coder of bench.py's CPU arm, independent of libpccgeo) == oracle/range_coder.py, byte for
    byte and symbol for symbol, escapes of every width included; and == the product's host coder."""
    rng = np.random.default_rng(n + 1)
    sym, idx = _random_stream(rng, gc_tab, n, spread)
    want = RC.unbounded_index_range_encode(sym, idx, gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
    got = RC.encode_c(sym, idx, gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
    assert got == want
    assert np.array_equal(RC.decode_c(got, idx, gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset']), sym)
    assert ops.range_encode(sym, np.array([0, n], np.int64), gc_tab, indexes=idx, threads=1) == [want]
    p = E.eb_init(6, np.random.default_rng(n))
    tab = E.eb_tables(p)
    zs = np.rint(rng.normal(size=(6, 5, 5, 5)) * 4 * spread ** 0.25).astype(np.int32)
    cidx = np.broadcast_to(np.arange(6, dtype=np.int32).reshape(6, 1, 1, 1), zs.shape)
    want = RC.unbounded_index_range_encode(zs, cidx, tab['cdf'], tab['cdf_length'], tab['offset'])
    assert RC.encode_c(zs, cidx, tab['cdf'], tab['cdf_length'], tab['offset']) == want
    assert np.array_equal(RC.decode_c(want, cidx, tab['cdf'], tab['cdf_length'], tab['offset']), zs)


def test_c_oracle_rejects_bad_index(gc_tab):
    with pytest.raises(ValueError):
        RC.encode_c([0], [len(gc_tab['cdf_length'])], gc_tab['cdf'], gc_tab['cdf_length'], gc_tab['offset'])
