"""Device range coder (csrc/rc_device.cu) against the host coder (range_coder.cpp): the strings must be the same bytes and
the decoded symbols the same integers, for both table-index modes, with escapes, ragged stream lengths and empty strings;
and the block loops must produce identical files and points with the coder on either side."""
import numpy as np
import pytest
import torch

from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table

pytestmark = pytest.mark.gpu


def _device_strings(sym, t, idx=None, channel_stride=0):
    packed, lengths, offsets, err = ops.range_encode_device(torch.from_numpy(sym).cuda(), ops.device_tables(t),
                                                            indexes=None if idx is None else torch.from_numpy(idx).cuda(),
                                                            channel_stride=channel_stride)
    assert int(err.item()) == 0
    lengths, offsets = lengths.cpu().numpy(), offsets.cpu().numpy()
    assert (lengths >= 0).all() and (np.diff(offsets) == lengths).all()
    buf = packed[:int(offsets[-1])].cpu().numpy().tobytes()
    return [buf[offsets[i]:offsets[i + 1]] for i in range(len(lengths))], packed, offsets


@pytest.mark.parametrize('ns,per,spread', [(4, 1000, 1.0), (3, 4097, 30.0), (2, 31, 3.0), (5, 1, 1.0), (2, 5000, 2000.0), (64, 16384, 0.6),
                                           (7, 333, 0.01)])
def test_indexed_tables_roundtrip_and_bytes(ns, per, spread):
    t = gaussian_tables(make_scale_table())
    rng = np.random.default_rng(ns * 1000 + per)
    idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
    sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * spread).astype(np.int32)
    offs = np.arange(ns + 1, dtype=np.int64) * per
    ref = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1))
    got, packed, boffs = _device_strings(sym, t, idx)
    assert got == ref
    out, err = ops.range_decode_device(packed, torch.from_numpy(boffs).cuda(), ns, per, ops.device_tables(t), indexes=torch.from_numpy(idx).cuda())
    assert int(err.item()) == 0
    assert np.array_equal(out.cpu().numpy(), sym)
    assert np.array_equal(ops.range_decode(ref, offs, t, indexes=idx.reshape(-1)).reshape(ns, per), sym)


@pytest.mark.parametrize('ns,per,spread', [(3, 2000, 1.0), (2, 257, 40.0), (4, 64, 0.05)])
def test_device_coder_against_the_oracle_coder_directly(ns, per, spread):
    """the device coder's bytes are the ORACLE's bytes (oracle/range_coder_c.c, no product code on the checking side), and the
    oracle decodes them back; round 1 only had this transitively (device == product host coder == Python oracle)"""
    from oracle import range_coder as RC
    t = gaussian_tables(make_scale_table())
    rng = np.random.default_rng(ns + per)
    idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
    sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * spread).astype(np.int32)
    got, packed, boffs = _device_strings(sym, t, idx)
    for i in range(ns):
        want = RC.encode_c(sym[i], idx[i], t['cdf'], t['cdf_length'], t['offset'])
        assert got[i] == bytes(want)
        assert np.array_equal(RC.decode_c(got[i], idx[i], t['cdf'], t['cdf_length'], t['offset']), sym[i])


def test_per_channel_tables_and_empty_strings():
    t = gaussian_tables(make_scale_table())
    rows = 8
    tt = {k: v[:rows] for k, v in t.items()}
    rng = np.random.default_rng(5)
    sym = rng.integers(-6, 7, (3, rows * 16)).astype(np.int32)
    sym[1] = 0
    offs = np.arange(4, dtype=np.int64) * rows * 16
    ref = ops.range_encode(sym.reshape(-1), offs, tt, channel_stride=16)
    got, packed, boffs = _device_strings(sym, tt, channel_stride=16)
    assert got == ref
    out, err = ops.range_decode_device(packed, torch.from_numpy(boffs).cuda(), 3, rows * 16, ops.device_tables(tt), channel_stride=16)
    assert int(err.item()) == 0 and np.array_equal(out.cpu().numpy(), sym)


def test_bad_table_index_sets_error_flag():
    t = gaussian_tables(make_scale_table())
    sym = torch.zeros((1, 40), dtype=torch.int32, device='cuda')
    idx = torch.full((1, 40), 64, dtype=torch.int32, device='cuda')
    assert int(ops.range_encode_device(sym, ops.device_tables(t), indexes=idx)[3].item()) == 1
    blob = torch.zeros(8, dtype=torch.uint8, device='cuda')
    boffs = torch.tensor([0, 4], dtype=torch.int64, device='cuda')
    assert int(ops.range_decode_device(blob, boffs, 1, 40, ops.device_tables(t), indexes=idx)[1].item()) == 1


@pytest.mark.parametrize('config,bias', [('c3p', -0.7), ('c1', 0.4)])
def test_block_loops_identical_with_device_coder(config, bias):
    """compress_blocks / decompress_blocks (V2 'c3p' with the hyperprior, V1 'c1') with the coder on the GPU vs in the host
    workers: same strings, same points; several ragged batches, two coder groups; and either side decodes the other's."""
    from pcc_geo_cnn_v2_b200 import synthetic
    from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType
    m = ModelConfigType[config].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=7, output_bias=bias))
    m.batch_size, m.coder_group_blocks = 3, 6
    blocks = synthetic.surface_blocks(11, size=64, seed=21) + [np.array([[0, 0, 0]], np.float32)]
    m.compress((1, 1, 64, 64, 64))
    res = {}
    for dev in (False, True, 'overlap'):   # 'overlap': the coder on a side stream under the next group's transforms
        m.device_coder, m.coder_overlap = bool(dev), dev == 'overlap'
        dl, meta, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
        dec, _ = m.decompress_blocks(None, dl[0], (64, 64, 64))
        res[dev] = (dl[0], [p.tobytes() for p in meta[0]['x_hat_list']], [p.tobytes() for p in dec])
    assert res[True][0] == res[False][0] and res['overlap'] == res[True]
    assert res[True][1] == res[False][1] and res[True][2] == res[False][2] and res[True][1] == res[True][2]
    assert sum(len(p) for p in res[True][2]) > 1000
    m.device_coder = False
    dec, _ = m.decompress_blocks(None, res[True][0], (64, 64, 64))
    assert [p.tobytes() for p in dec] == res[False][2]


def test_corrupt_streams_never_crash_the_decoder():
    """Random bytes, truncated and empty strings: the device decoder must return (garbage symbols, possibly the error flag),
    like the host decoder -- and agree with it symbol for symbol whenever neither flags an error."""
    t = gaussian_tables(make_scale_table())
    dt = ops.device_tables(t)
    rng = np.random.default_rng(11)
    ns, per = 12, 700
    idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
    strings = [rng.bytes(int(n)) for n in rng.integers(0, 900, ns)]
    strings[3] = b''
    strings[5] = b'\xff' * 400            # long escape runs: unary width code beyond any valid value
    offs = np.zeros(ns + 1, np.int64)
    offs[1:] = np.cumsum([len(s) for s in strings])
    blob = torch.from_numpy(np.frombuffer(b''.join(strings) + b'\0', np.uint8).copy()).cuda()
    out, err = ops.range_decode_device(blob, torch.from_numpy(offs).cuda(), ns, per, dt, indexes=torch.from_numpy(idx).cuda())
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    soffs = np.arange(ns + 1, dtype=np.int64) * per
    for i in range(ns):   # stream by stream: the host decoder raises on a corrupt escape code
        try:
            ref = ops.range_decode([strings[i]], soffs[:2], t, indexes=idx[i])
        except Exception:
            continue
        assert np.array_equal(got[i], ref), i
    assert int(err.item()) in (0, 1)
