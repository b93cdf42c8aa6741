"""The device range ENCODER's arithmetic (csrc/rc_device.cu: rc_map_symbol / rc_enc_symbol / rc_enc_finish, the inline functions
the kernels call) run on the host through pccgeo_range_encode_emulate_host, pinned against the production host coder
(range_coder.cpp) byte for byte.  No GPU needed; the kernels themselves are covered by tests/test_gpu_range_coder_device.py."""
import numpy as np
import pytest

from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table


@pytest.mark.parametrize('ns,per,spread', [(4, 1000, 1.0), (3, 4097, 30.0), (2, 31, 3.0), (5, 1, 1.0), (2, 5000, 2000.0)])
def test_emulated_device_encoder_matches_host_coder(ns, per, spread):
    t = gaussian_tables(make_scale_table())
    rng = np.random.default_rng(ns * 1000 + per)
    idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
    sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * spread).astype(np.int32)  # spread > 1: escapes
    offs = np.arange(ns + 1, dtype=np.int64) * per
    assert ops.range_encode_emulate(sym, ns, t, indexes=idx) == ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=2)


def test_emulated_device_encoder_per_channel_tables_and_zero_streams():
    t = gaussian_tables(make_scale_table())
    rows = 8
    tt = {k: v[:rows] for k, v in t.items()}
    rng = np.random.default_rng(5)
    sym = rng.integers(-6, 7, (3, rows * 16)).astype(np.int32)
    offs = np.arange(4, dtype=np.int64) * rows * 16
    assert ops.range_encode_emulate(sym, 3, tt, channel_stride=16) == ops.range_encode(sym.reshape(-1), offs, tt, channel_stride=16, threads=1)
    z = np.zeros((2, 100), np.int32)
    assert ops.range_encode_emulate(z, 2, t, indexes=z) == ops.range_encode(z.reshape(-1), np.array([0, 100, 200]), t, indexes=z.reshape(-1))


def test_emulated_device_encoder_rejects_bad_table_index():
    t = gaussian_tables(make_scale_table())
    with pytest.raises(Exception):
        ops.range_encode_emulate(np.zeros((1, 4), np.int32), 1, t, indexes=np.full((1, 4), 64, np.int32))


def test_device_encoder_carry_walks_back_through_memory():
    """0x12 FF FF | FF FF FF FF (register) + 1 -> 0x13 00 00 | 00 00 00 00"""
    from pcc_geo_cnn_v2_b200 import _lib as L
    buf = np.array([0x00, 0x12, 0xFF, 0xFF], np.uint8)
    assert L.lib().pccgeo_rc_carry_probe_host(L.ptr(buf), 4) == 0
    assert buf.tolist() == [0x00, 0x13, 0x00, 0x00]
    buf = np.array([0x00, 0x12, 0xFF, 0x7F], np.uint8)
    assert L.lib().pccgeo_rc_carry_probe_host(L.ptr(buf), 4) == 0
    assert buf.tolist() == [0x00, 0x12, 0xFF, 0x80]


def test_emulated_device_encoder_many_random_streams():
    t = gaussian_tables(make_scale_table())
    rng = np.random.default_rng(99)
    for spread in (0.02, 0.3, 1.0, 8.0):
        ns, per = 48, 3000
        idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
        sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * spread).astype(np.int32)
        offs = np.arange(ns + 1, dtype=np.int64) * per
        assert ops.range_encode_emulate(sym, ns, t, indexes=idx) == ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1))


def test_compact_tables_for_the_device_decoder():
    """pccgeo_range_compact_tables_host: rows back to back as 16-bit entries without their (implied) final 2^16; rejects
    rows that are not 16-bit CDFs."""
    from pcc_geo_cnn_v2_b200 import _lib as L
    t = gaussian_tables(make_scale_table())
    cdf, cl = np.ascontiguousarray(t['cdf'], np.int32), np.ascontiguousarray(t['cdf_length'], np.int32)
    rows = cdf.shape[0]
    total = L.lib().pccgeo_range_compact_tables_host(L.ptr(cdf), cdf.shape[1], L.ptr(cl), rows, None, None)
    assert total == int((cl - 1).sum()) and total * 2 + rows * 12 < 48 * 1024      # fits the default shared-memory window
    c16, start = np.zeros(total, np.uint16), np.zeros(rows, np.int32)
    assert L.lib().pccgeo_range_compact_tables_host(L.ptr(cdf), cdf.shape[1], L.ptr(cl), rows, L.ptr(c16), L.ptr(start)) == total
    for r in (0, 1, 17, rows - 1):
        n = int(cl[r]) - 1
        assert np.array_equal(c16[start[r]:start[r] + n], cdf[r, :n]) and cdf[r, n] == 65536
    bad = cdf.copy()
    bad[3, int(cl[3]) - 1] = 65535      # a row that does not end at 2^16
    assert L.lib().pccgeo_range_compact_tables_host(L.ptr(bad), bad.shape[1], L.ptr(cl), rows, None, None) == -1
    bad = cdf.copy()
    bad[5, 1] = bad[5, 2] + 1           # not monotone
    assert L.lib().pccgeo_range_compact_tables_host(L.ptr(bad), bad.shape[1], L.ptr(cl), rows, None, None) == -1
