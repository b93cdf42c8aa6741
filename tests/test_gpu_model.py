"""End-to-end parity of the CUDA path with the oracle and the committed golden fixtures, through the reference's
own call signatures (ModelConfigType[...].build(), compress_blocks / decompress_blocks, TransformType).

Bars: int32 symbols / indexes / byte strings / decoded points are compared exactly; because analysis latents are
*rounded*, a different fp32 summation order can legitimately flip a symbol that sits within ~1e-5 of a rounding
boundary, so symbol comparisons allow a stated, tiny flip budget (and strings are compared only where the symbols
agree).  Decoding the oracle's strings must reproduce the oracle's points except voxels the fixture marks fragile
(|x_hat - t| < 1e-4)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.model import OracleModel
from pcc_geo_cnn_v2_b200 import model_transforms as MT
from pcc_geo_cnn_v2_b200 import synthetic
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _model(config, seed, **kw):
    m = ModelConfigType[config].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=seed, **kw))
    return m


def _as_set(p):
    return {tuple(int(v) for v in r) for r in np.asarray(p)}


@pytest.fixture(autouse=True)
def _default_precision():
    MT.set_precision('bf16x3')
    yield
    MT.set_precision('bf16x3')


# reference src/test_model_transforms.py:14-73, run against the CUDA implementation (channels_last like the original)
@pytest.mark.parametrize('name,filters,inp,kw,expect', [
    ('AnalysisTransformV1', 1, (1, 8, 8, 8, 1), {}, (1, 1, 1, 1, 1)),
    ('SynthesisTransformV1', 2, (1, 1, 1, 1, 1), {}, (1, 8, 8, 8, 1)),
    ('AnalysisTransformV2', 2, (1, 8, 8, 8, 1), {}, (1, 1, 1, 1, 2)),
    ('AnalysisTransformV2', 2, (1, 8, 8, 8, 1), {'residual_mode': 'concat'}, (1, 1, 1, 1, 2)),
    ('SynthesisTransformV2', 2, (1, 1, 1, 1, 1), {}, (1, 8, 8, 8, 1)),
    ('SynthesisTransformV2', 2, (1, 1, 1, 1, 1), {'residual_mode': 'concat'}, (1, 8, 8, 8, 1)),
    ('AnalysisTransformProgressiveV2', 4, (1, 8, 8, 8, 1), {}, (1, 1, 1, 1, 4)),
    ('SynthesisTransformProgressiveV2', 4, (1, 1, 1, 1, 1), {}, (1, 8, 8, 8, 1)),
    ('HyperAnalysisTransform', 1, (1, 8, 8, 8, 1), {}, (1, 4, 4, 4, 1)),
    ('HyperSynthesisTransform', 1, (1, 1, 1, 1, 1), {}, (1, 2, 2, 2, 1)),
])
def test_reference_shape_contracts_on_gpu(name, filters, inp, kw, expect):
    layer = MT.TransformType[name].value(filters, data_format='channels_last', **kw)
    y = layer(torch.zeros(inp, device='cuda'))
    assert tuple(y.shape) == expect


def test_blocks_shape_contracts_on_gpu():
    x = torch.zeros((1, 8, 8, 8, 1), device='cuda')
    assert tuple(MT.AnalysisBlock(1, data_format='channels_last')(x).shape) == (1, 4, 4, 4, 1)
    assert tuple(MT.AnalysisBlock(1, data_format='channels_last', residual_mode='concat')(x).shape) == (1, 4, 4, 4, 2)
    y = torch.zeros((1, 1, 1, 1, 1), device='cuda')
    assert tuple(MT.SynthesisBlock(1, data_format='channels_last')(y).shape) == (1, 2, 2, 2, 1)
    assert tuple(MT.SynthesisBlock(1, data_format='channels_last', residual_mode='concat')(y).shape) == (1, 2, 2, 2, 2)
    with pytest.raises(TypeError):
        MT.AnalysisBlock(1, data_format='channels_last')(torch.zeros((1, 8, 8, 8, 1)))   # CPU tensor: no fallback


@pytest.mark.parametrize('fixture', sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, 'c[0-9]*_*.npz'))))
@pytest.mark.parametrize('precision', ['bf16x3', 'fp32'])
def test_golden_fixture(fixture, precision):
    g = np.load(os.path.join(GOLDEN, fixture))
    config, size, seed, nb = str(g['config']), int(g['size']), int(g['seed']), int(g['n_blocks'])
    MT.set_precision(precision)
    m = _model(config, seed, **{k: float(g[k]) for k in ('gain', 'synthesis_gain', 'output_bias') if k in g})
    m.compress((1, 1, size, size, size))
    blocks = [g[f'block{j}'].astype(np.float32) for j in range(nb)]
    v2 = f'z_sym0' in g
    if size == 64:   # the full-size fixtures exercise non-trivial latents and non-empty decodes
        assert all(np.count_nonzero(g[f'y_sym{j}']) > g[f'y_sym{j}'].size // 20 and len(g[f'points{j}']) > 500 for j in range(nb))
    # ---- encoder side: symbols / indexes / strings
    from pcc_geo_cnn_v2_b200 import ops
    from pcc_geo_cnn_v2_b200.entropy_models import GaussianConditional
    from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords
    x = ops.densify(torch.from_numpy(blocks_to_coords(blocks)).cuda(), nb, size, size, size)
    dev = m._encode_device(x)
    strings = m._encode_host(dev)
    ysym = dev['y_sym'].cpu().numpy()
    ytotal = yflips = ztotal = zflips = 0
    for j in range(nb):
        d = ysym[j] != g[f'y_sym{j}']
        assert np.abs(ysym[j].astype(np.int64) - g[f'y_sym{j}']).max() <= 1      # a flip moves a symbol by one step
        yflips += int(d.sum())
        ytotal += d.size
        same = not d.any()
        if v2:
            zs = dev['z_sym'][j].cpu().numpy()
            dz = zs != g[f'z_sym{j}']
            assert np.abs(zs.astype(np.int64) - g[f'z_sym{j}']).max() <= 1
            zflips += int(dz.sum())
            ztotal += dz.size
            if dz.any():
                continue                     # a flipped z changes sigma downstream: nothing else is comparable for this block
            idx_diff = dev['indexes'][j].cpu().numpy() != g[f'idx{j}']
            assert idx_diff.mean() <= 2e-3   # sigma within ~1e-5 of a table boundary may land in the neighbouring bin
            same &= not idx_diff.any()
        if same:
            for i, s in enumerate(strings[j]):
                assert s == g[f'string{j}_{i}'].tobytes(), f'string {i} of block {j} differs'
    # flip budgets: fp32 kernels differ from the oracle only by summation order; bf16x3 carries ~1e-5 relative error
    budget = 2e-5 if precision == 'fp32' else 4e-4
    assert yflips <= max(1, int(budget * ytotal)), f'{yflips}/{ytotal} y symbols differ from the oracle'
    # the hyper-latent of the high-gain 64^3 V1 fixtures sits behind k9/k5 layers (4 000-term sums): a few more boundary hits
    zbudget = budget if precision == 'fp32' or 'gain' not in g else 1e-3
    assert zflips <= max(1, int(zbudget * max(ztotal, 1))), f'{zflips}/{ztotal} z symbols differ from the oracle'
    # ---- decoder side: oracle-made strings -> points, for the blocks whose scale indexes agree with the oracle's (an
    # index flip desynchronises the y stream of ANY two implementations -- the reference pins this step to the CPU and
    # retries for the same reason, patch_gaussian_conditional.py:105-116, decompress_octree.py:69-131)
    # Blocks with an index flip are NOT skipped: they are decoded stage by stage with the oracle's indexes (range decoder ->
    # the oracle's symbols exactly; synthesis + threshold -> the oracle's points), so every block of every fixture is decoded
    # from the oracle's strings in every precision.
    from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables
    from pcc_geo_cnn_v2_b200.model_types import threshold_f32
    data, keep = [], []
    for j in range(nb):
        strs = tuple(g[f'string{j}_{i}'].tobytes() for i in range(2 if v2 else 1))
        ok = True
        if v2:
            zshape = (m.num_filters,) + (size // 16,) * 3
            zsym = m.entropy_bottleneck.decode_symbols([strs[1]], zshape)
            assert np.array_equal(zsym[0], g[f'z_sym{j}'])                    # integer path: exact
            z_hat = ops.eb_dequantize(torch.from_numpy(zsym).cuda(), m.entropy_bottleneck.device_params())
            idx = GaussianConditional(m.hyper_synthesis_transform(z_hat), m.scale_table).indexes()
            idx_diff = idx[0].cpu().numpy() != g[f'idx{j}']
            assert idx_diff.mean() <= 2e-3
            ok = not idx_diff.any()
            gidx = g[f'idx{j}'].astype(np.int32).reshape(-1)
            ysym = ops.range_decode([strs[0]], np.array([0, gidx.size], np.int64), gaussian_tables(m.scale_table), indexes=gidx, threads=1)
            assert np.array_equal(ysym.reshape(g[f'y_sym{j}'].shape), g[f'y_sym{j}'])
            y_hat = torch.from_numpy(g[f'y_sym{j}'][None]).float().cuda()
            thr = torch.from_numpy(threshold_f32(m.thresholds, np.full(1, 128))).cuda()
            _, bits, _ = m.synthesis_transform.packed(y_hat, thr, want_f32=False)
            pts = ops.bits_to_points(bits.cpu().numpy(), (size, size, size), 1)[0]
            assert (_as_set(g[f'points{j}']) ^ _as_set(pts)) <= _as_set(g[f'fragile{j}']), f'block {j} (stage-wise decode)'
        if ok:
            data.append((strs, 128))
            keep.append(j)
    # fp32 kernels reproduce every scale index of (nearly) every block; in bf16x3 a block of 32 768 indexes usually has a couple
    # of boundary hits (measured on B200: 7 of 32 blocks keep all of them, every block keeps >= 99.98 %), which is why the
    # stage-wise decode above -- not this end-to-end pass -- is what covers every block in the default precision
    if precision == 'fp32':
        assert len(keep) >= nb - 1, f'{len(keep)}/{nb} blocks reproduce every scale index of the oracle'
    m.decompress()
    dec, _ = m.decompress_blocks(None, data, (size, size, size)) if data else ([], [])
    for j, pts in zip(keep, dec):
        want, got = _as_set(g[f'points{j}']), _as_set(pts)
        assert pts.dtype == np.float32
        assert (want ^ got) <= _as_set(g[f'fragile{j}']), f'{len(want ^ got)} voxels differ beyond the fragile set'
        assert pts.tolist() == sorted(pts.tolist())      # argwhere (C) order


@pytest.mark.parametrize('config,size,bias', [('c3p', 64, -0.7), ('c1', 64, 0.4), ('c2', 32, 0.47), ('c3', 64, -0.7)])
def test_encode_decode_self_consistency(config, size, bias):
    """decompress_octree.py --debug contract (decompress_octree.py:94-119): the decoder's point set equals the
    encoder's exactly, block by block, here over a ragged batch (incl. a 1-point block)."""
    m = _model(config, 7, output_bias=bias)   # per-config output bias: the decoded point sets must not be empty
    m.batch_size = 4
    blocks = synthetic.surface_blocks(6, size=size, seed=9) + [np.array([[0, 0, 0]], np.float32)]
    m.compress((1, 1, size, size, size))
    data_list, metadata, _ = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True)
    assert len(data_list) == 1 and len(data_list[0]) == len(blocks)
    assert all(t == 128 and isinstance(s, tuple) and all(isinstance(b, bytes) for b in s) for s, t in data_list[0])
    m.decompress()
    dec, _ = m.decompress_blocks(None, data_list[0], (size, size, size))
    enc_pts = metadata[0]['x_hat_list']
    for a, b in zip(enc_pts, dec):
        assert np.array_equal(a, b)
    assert sum(len(p) for p in dec) > 100, [len(p) for p in dec]


@pytest.mark.parametrize('config,size,bias', [('c3p', 64, -0.7), ('c1', 64, 0.4), ('c2', 64, 0.47)])
def test_graph_pipeline_equals_eager_pipeline(config, size, bias):
    """The CUDA-graph stage replays of the block loops produce the same bytes and points as the eager launches, over
    several ragged batches in flight (static buffers are shared between batches: a stage must never read a tensor that a
    later batch's replay has already overwritten), and survive a parameter change (graphs are re-captured).  The output
    bias is chosen per config so that the decoded point sets are non-empty and differ from block to block."""
    m = _model(config, 7, output_bias=bias)
    m.batch_size = 3
    blocks = synthetic.surface_blocks(8, size=size, seed=21)
    m.compress((1, 1, size, size, size))
    out = {}
    for graphs in (True, False, True):
        m.use_graphs = graphs
        dl, meta, _ = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True)
        dec, _ = m.decompress_blocks(None, dl[0], (size, size, size))
        cur = ([s for s, _ in dl[0]], [p.tobytes() for p in meta[0]['x_hat_list']], [p.tobytes() for p in dec])
        assert out.setdefault('ref', cur) == cur
        assert min(len(p) for p in dec) > 0 and len({len(p) for p in dec}) > 4, [len(p) for p in dec]
    assert m._graphs, 'the graph path did not run'
    # new parameters -> stale graphs must not be replayed
    m.use_graphs = True
    m.set_weights(synthetic.trained_like_weights(m, seed=8))
    dl2, meta2, _ = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True)
    m.use_graphs = False
    dl3, meta3, _ = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True)
    assert [s for s, _ in dl2[0]] == [s for s, _ in dl3[0]] and [s for s, _ in dl2[0]] != out['ref'][0]
    assert all(np.array_equal(a, b) for a, b in zip(meta2[0]['x_hat_list'], meta3[0]['x_hat_list']))


def test_train_forward_matches_oracle():
    """tr_train.py path forward values (model_types.py:327-355): loss = lambda*FL + mbpov with the SAME noise."""
    m = _model('c3p', 21)
    o = OracleModel('c3p')
    w = m.get_weights()
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, w['entropy_bottleneck'])
    from oracle.model import sparse_to_dense
    blocks = synthetic.surface_blocks(2, size=32, seed=3)
    x = np.concatenate([sparse_to_dense(b, (1, 1, 32, 32, 32)) for b in blocks])
    g = torch.Generator().manual_seed(0)
    ny = torch.rand((2, 64, 4, 4, 4), generator=g) - 0.5
    nz = torch.rand((2, 64, 2, 2, 2), generator=g) - 0.5
    ref = o.train_forward(x, 2, 0.75, 1e-4, ny, nz)
    m.train(torch.from_numpy(x).cuda(), 2, 0.75, 1e-4, noise_y=ny.cuda(), noise_z=nz.cuda())
    for got, want in ((m.train_fl, ref['fl']), (m.train_mbpov, ref['mbpov']), (m.train_loss, ref['loss'])):
        assert abs(float(got) - float(want)) < 1e-4 * abs(float(want)), (float(got), float(want))
    rel = (m.y.cpu() - ref['y']).abs().max() / ref['y'].abs().max()
    assert float(rel) < 3e-5


def test_training_host_glue_matches_reference_semantics():
    """input_fn / pc_to_tf / process_x / quantize_tensor / binary_classification_summaries / add_channels
    (model_types.py:23-62,91-105,121-125) against their numpy statements."""
    from pcc_geo_cnn_v2_b200 import model_types as MT
    rng = np.random.default_rng(3)
    pts = [np.unique(rng.integers(0, 16, size=(n, 3)), axis=0) for n in (40, 7, 120, 1, 64)]
    for fmt, shape in (('channels_first', (1, 16, 16, 16)), ('channels_last', (16, 16, 16, 1))):
        dense = MT.process_x(MT.pc_to_tf(pts[0], shape, fmt), shape)
        want = np.zeros((16, 16, 16), np.float32)
        want[pts[0][:, 0], pts[0][:, 1], pts[0][:, 2]] = 1
        assert dense.shape == shape and np.array_equal(dense.cpu().numpy().reshape(16, 16, 16), want)
        assert MT.add_channels([4, 5, 6], 64, fmt) == ([64, 4, 5, 6] if fmt == 'channels_first' else [4, 5, 6, 64])
    np.random.seed(42)
    got = list(MT.input_fn(pts, 2, (1, 16, 16, 16), 'channels_first', repeat=False, shuffle=True))
    assert [tuple(b.shape) for b in got] == [(2, 1, 16, 16, 16)] * 2 + [(1, 1, 16, 16, 16)]
    np.random.seed(42)
    order = np.random.permutation(len(pts))
    sums = [float(b.sum()) for b in got]
    assert sums == [float(len(pts[order[0]]) + len(pts[order[1]])), float(len(pts[order[2]]) + len(pts[order[3]])), float(len(pts[order[4]]))]
    it = MT.input_fn(pts, 3, (1, 16, 16, 16), 'channels_first', repeat=True, shuffle=False)
    got = [next(it) for _ in range(3)]   # shuffle -> repeat -> batch: batches are always full and span the epoch boundary
    assert [tuple(b.shape) for b in got] == [(3, 1, 16, 16, 16)] * 3
    ln = [len(p) for p in pts]
    assert [float(b.sum()) for b in got] == [float(ln[0] + ln[1] + ln[2]), float(ln[3] + ln[4] + ln[0]), float(ln[1] + ln[2] + ln[3])]
    x = torch.tensor([[-0.2, 0.4, 0.5, 0.51, 1.7]], device='cuda')
    assert MT.quantize_tensor(x).tolist() == [[0, 0, 0, 1, 1]]
    xq = torch.tensor([1, 1, 0, 0, 1, 0], device='cuda', dtype=torch.uint8)
    xt = torch.tensor([1, 0, 0, 1, 1, 0], device='cuda', dtype=torch.uint8)
    bc = MT.binary_classification_summaries(xq, xt)
    assert abs(float(bc['bc/precision']) - 2 / 3) < 1e-12 and abs(float(bc['bc/recall']) - 2 / 3) < 1e-12
    assert abs(float(bc['bc/accuracy']) - 4 / 6) < 1e-12 and abs(float(bc['bc/specificity']) - 2 / 3) < 1e-12


@pytest.mark.parametrize('config,size,bias', [('c3p', 64, -0.7), ('c1', 32, 0.4)])
def test_debug_contract_of_decompress_octree(config, size, bias):
    """compress_octree.py --debug saves the encoder's per-block debug tensors and decoded blocks; decompress_octree.py --debug
    (decompress_octree.py:64-131) re-decodes and requires, key by key, the decoder's debug tensors to match the encoder's
    (numbers: atol 1e-3 / rtol 1e-7 via assert_allclose; objects: equality) and the decoded blocks to be equal."""
    m = _model(config, 7, output_bias=bias)
    m.batch_size = 3
    blocks = synthetic.surface_blocks(5, size=size, seed=12)
    m.compress((1, 1, size, size, size))
    assert m.x.shape == (1, 1, size, size, size) and len(m.strings) == (2 if config == 'c3p' else 1)
    data_list, meta, dbg_enc = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True, debug=True)
    plain, meta_plain, dbg_none = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True, debug=False)
    assert dbg_none == [None] * len(blocks)
    assert [d[0] for d in plain[0]] == [d[0] for d in data_list[0]]          # debug mode changes neither bytes nor points
    assert all(np.array_equal(a, b) for a, b in zip(meta[0]['x_hat_list'], meta_plain[0]['x_hat_list']))
    m.decompress()
    assert m.x_shape_t.shape == (3,) and len(m.strings_t) == len(m.strings)
    dec, dbg_dec = m.decompress_blocks(None, data_list[0], (size, size, size), debug=True)
    assert len(dbg_enc) == len(dbg_dec) == len(blocks)
    want_keys = {'y_hat', 'x_hat'} | ({'z_hat', 'sigma_hat', 'decompress/indexes', 'decompress/symbols', 'decompress/outputs',
                                       'decompress/strings', 'decompress/build/scale_table', 'decompress/quantized_cdf'}
                                      if config == 'c3p' else set())
    for j, (de, dd) in enumerate(zip(dbg_enc, dbg_dec)):
        assert want_keys <= set(dd) and set(dd) == set(de)
        for key in dd:                                                       # the loop of decompress_octree.py:94-119
            v1, v2 = np.asarray(dd[key]), np.asarray(de[key])
            if v1.dtype == object:
                np.testing.assert_equal(v1, v2, err_msg=f'Values did not match for key {key}')
            else:
                assert np.issubdtype(v1.dtype, np.number)
                np.testing.assert_allclose(v1, v2, rtol=1e-7, atol=0.001, err_msg=f'Values did not match for key {key}')
        assert dd['x_hat'].shape == (1, 1, size, size, size) and dd['y_hat'].shape[0] == 1
        np.testing.assert_equal(dec[j], meta[0]['x_hat_list'][j])
        # the debug x_hat is the tensor the points come from
        assert np.array_equal(np.argwhere(np.clip(dd['x_hat'][0, 0], 0, 1) > m.thresholds[128]).astype(np.float32), dec[j])


@pytest.mark.parametrize('config,device_coder', [('c3p', False), ('c3p', True), ('c1', False)])
def test_block_loops_do_not_depend_on_the_stream_schedule(config, device_coder, monkeypatch):
    """The block loops run their stage graphs on two compute streams and their copies on dedicated copy streams, with batches
    alternating between two sets of static buffers.  The bytes and the points must be what the
    single-stream schedule (PCCGEO_SIDE_COPIES=0) produces, over many ragged batches, on a fresh model each time (first-use
    captures and buffer allocations are part of what is being checked), and twice in a row on the same model."""
    size = 32
    blocks = synthetic.surface_blocks(45, size=size, seed=21) + [np.array([[1, 2, 3]], np.float32)]
    results = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('PCCGEO_SIDE_COPIES', mode)
        m = _model(config, 5, output_bias={'c3p': -0.7, 'c1': 0.4}[config])
        m.batch_size = 8                       # 46 blocks -> 5 full batches + one of 6
        m.device_coder = device_coder
        m.compress((1, 1, size, size, size))
        m.decompress()
        runs = []
        for _ in range(2):
            data_list, metadata, _ = m.compress_blocks(None, blocks, None, None, size, 0, fixed_threshold=True)
            dec, _ = m.decompress_blocks(None, data_list[0], (size, size, size))
            runs.append(([s for s, _ in data_list[0]], metadata[0]['x_hat_list'], dec))
        assert runs[0][0] == runs[1][0]
        assert all(np.array_equal(a, b) for a, b in zip(runs[0][2], runs[1][2]))
        results[mode] = runs[0]
    assert results['0'][0] == results['1'][0]                                             # strings
    for enc0, enc1, dec0, dec1 in zip(results['0'][1], results['1'][1], results['0'][2], results['1'][2]):
        assert np.array_equal(enc0, enc1) and np.array_equal(dec0, dec1) and np.array_equal(enc1, dec1)
