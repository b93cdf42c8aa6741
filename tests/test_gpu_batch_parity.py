"""Whole-batch parity on the headline configuration (BASELINE.json configs[1]: c3p, batch 32, 64^3 blocks): the 24 ModelNet40
blocks of tests/golden/modelnet_blocks.npz + 8 synthetic surface blocks go through the CUDA path in ONE batch and are compared
with the oracle block by block -- the reference's own gate is per-block exact equality of the decoder's tensors with the
encoder's (src/decompress_octree.py:94-119).

Every decode stage is checked against the oracle for EVERY block (nothing is skipped):
  * range decoding of the oracle's strings with the oracle's indexes (host and device coder)  -> the oracle's symbols, exactly;
  * hyper-synthesis on the oracle's z_hat -> scale indexes: agreement rate per block (an index is a threshold on sigma, so
    a value within rounding distance of a table boundary may land in the neighbouring bin);
  * synthesis + clip + threshold on the oracle's y_hat -> the oracle's points outside the fragile set (|x_hat - t| < 1e-4).
and the full decompress_blocks() path decodes the oracle's strings for every block whose indexes agree with the oracle's
(any two implementations desynchronise on an index flip; the reference pins that step to the CPU and retries for the same
reason, src/utils/patch_gaussian_conditional.py:105-116, src/decompress_octree.py:69-131); the share of such blocks is asserted.
Also here: value parity of residual_mode='concat' (src/model_transforms.py:37-38).  (64^3 goldens of the V1 k9/k5 configs:
tests/golden/c1_64.npz, c2_64.npz through test_gpu_model.py::test_golden_fixture.)
"""
import os

import numpy as np
import pytest
import torch

from oracle import transforms as T
from oracle.model import OracleModel, sparse_to_dense
from pcc_geo_cnn_v2_b200 import model_transforms as MT
from pcc_geo_cnn_v2_b200 import ops, synthetic
from pcc_geo_cnn_v2_b200.entropy_models import GaussianConditional, gaussian_tables
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords, threshold_f32

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SIZE = 64


@pytest.fixture(autouse=True)
def _default_precision():
    MT.set_precision('bf16x3')
    yield
    MT.set_precision('bf16x3')


def _as_set(p):
    return {tuple(int(v) for v in r) for r in np.asarray(p)}


def _headline_blocks():
    g = np.load(os.path.join(GOLDEN, 'modelnet_blocks.npz'))
    real = [g[f'block{i}'].astype(np.float32) for i in range(len(g['names']))]
    return real + synthetic.surface_blocks(8, size=SIZE, seed=5)


def _oracle_of(m, config):
    o = OracleModel(config)
    w = m.get_weights()
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, w['entropy_bottleneck'])
    return o


# min_keep: share of blocks whose EVERY scale index equals the oracle's.  An index is a threshold on sigma against a 64-level table
# (log spacing 0.123), so an element flips with probability ~ 2*delta/0.123 for a relative sigma difference delta, and a block of
# 32 768 indexes keeps them all with probability (1 - p)^32768: fp32 kernels vs the fp32 oneDNN oracle (delta ~ 3e-7, summation
# order only) measure 26/32 blocks on B200; bf16x3 (delta ~ 1e-5) keeps few blocks but >= 99.8 % of the indexes of every block.
@pytest.mark.parametrize('precision,sym_budget,min_keep', [('bf16x3', 4e-4, 0.0), ('fp32', 2e-5, 0.6)])
def test_headline_batch_against_oracle_block_by_block(precision, sym_budget, min_keep):
    torch.set_num_threads(os.cpu_count() or 1)
    MT.set_precision(precision)
    blocks = _headline_blocks()
    nb = len(blocks)
    assert nb == 32
    m = ModelConfigType['c3p'].build(batch_size=nb)
    m.set_weights(synthetic.trained_like_weights(m, seed=42))
    o = _oracle_of(m, 'c3p')
    t128 = o.thresholds[128]

    # ---- the oracle, block by block (CPU)
    ref = []
    with torch.no_grad():
        for b in blocks:
            strings, x_hat, t = o.compress(sparse_to_dense(b, (1, 1, SIZE, SIZE, SIZE)))
            xh = x_hat[0, 0].numpy()
            ref.append({'strings': strings, 'y_sym': t['y_symbols'][0].numpy(), 'z_sym': t['z_symbols'][0].numpy(),
                        'idx': t['indexes'][0].numpy(), 'z_hat': t['z_hat'], 'y_hat': t['y_hat'],
                        'points': np.argwhere(np.clip(xh, 0, 1) > t128), 'fragile': np.argwhere(np.abs(xh - t128) < 1e-4)})

    # ---- encoder side, ONE batch of 32 through the public block loop and through the device pass it wraps
    m.compress((1, 1, SIZE, SIZE, SIZE))
    data_list, meta, _ = m.compress_blocks(None, blocks, None, None, SIZE, 0, fixed_threshold=True)
    x = ops.densify(torch.from_numpy(blocks_to_coords(blocks)).cuda(), nb, SIZE, SIZE, SIZE)
    dev = m._encode_device(x)
    ysym, zsym, idx = dev['y_sym'].cpu().numpy(), dev['z_sym'].cpu().numpy(), dev['indexes'].cpu().numpy()
    yflips = zflips = 0
    enc_idx_rate, strings_equal = [], 0
    for j, r in enumerate(ref):
        dy, dz = ysym[j] != r['y_sym'], zsym[j] != r['z_sym']
        assert np.abs(ysym[j].astype(np.int64) - r['y_sym']).max() <= 1 and np.abs(zsym[j].astype(np.int64) - r['z_sym']).max() <= 1
        yflips += int(dy.sum())
        zflips += int(dz.sum())
        if dz.any():
            continue      # a flipped z changes sigma everywhere downstream: indexes / strings are not comparable for this block
        di = idx[j] != r['idx']
        enc_idx_rate.append(1.0 - float(di.mean()))
        if not dy.any() and not di.any():
            assert data_list[0][j][0] == r['strings'], f'block {j}: same symbols and indexes but different bytes'
            strings_equal += 1
    assert yflips <= max(1, int(sym_budget * ysym.size)), f'{yflips}/{ysym.size} y symbols differ from the oracle'
    assert zflips <= max(1, int(sym_budget * zsym.size)), f'{zflips}/{zsym.size} z symbols differ from the oracle'
    assert len(enc_idx_rate) >= nb // 2 and min(enc_idx_rate) >= 1 - 2e-3, enc_idx_rate   # blocks without a z flip

    # ---- decoder side, stage by stage, EVERY block, fed with the oracle's own data
    gtab, etab = gaussian_tables(m.scale_table), m.entropy_bottleneck.tables
    per_y, per_z = ref[0]['y_sym'].size, ref[0]['z_sym'].size
    y_strings, z_strings = [r['strings'][0] for r in ref], [r['strings'][1] for r in ref]
    ref_idx = np.stack([r['idx'] for r in ref]).astype(np.int32)
    offs_y, offs_z = np.arange(nb + 1, dtype=np.int64) * per_y, np.arange(nb + 1, dtype=np.int64) * per_z
    # (1) range decoding: bytes -> symbols, integer path, exact (host coder through the C ABI)
    zs = ops.range_decode(z_strings, offs_z, etab, channel_stride=per_z // 64, threads=4).reshape(nb, -1)
    ys = ops.range_decode(y_strings, offs_y, gtab, indexes=ref_idx.reshape(-1), threads=4).reshape(nb, -1)
    for j, r in enumerate(ref):
        assert np.array_equal(zs[j], r['z_sym'].reshape(-1)) and np.array_equal(ys[j], r['y_sym'].reshape(-1)), j
    # ... and the device coder
    blob, boffs = m._upload_strings(y_strings)
    out, err = ops.range_decode_device(blob, boffs, nb, per_y, ops.device_tables(gtab), indexes=torch.from_numpy(ref_idx).cuda())
    assert int(err.item()) == 0 and np.array_equal(out.cpu().numpy().reshape(nb, -1), ys)
    # (2) hyper-synthesis on the oracle's z_hat -> indexes
    z_hat = torch.cat([r['z_hat'] for r in ref]).float().cuda()
    got_idx = GaussianConditional(m.hyper_synthesis_transform(z_hat), m.scale_table).indexes().cpu().numpy()
    dec_rate = [1.0 - float((got_idx[j] != r['idx']).mean()) for j, r in enumerate(ref)]
    assert min(dec_rate) >= 1 - 2e-3, dec_rate
    keep = [j for j in range(nb) if dec_rate[j] == 1.0]
    # (3) synthesis + clip + threshold + pack + point extraction on the oracle's y_hat
    y_hat = torch.cat([r['y_hat'] for r in ref]).float().cuda()
    thr = torch.from_numpy(threshold_f32(m.thresholds, np.full(nb, 128))).cuda()
    _, bits, counts = m.synthesis_transform.packed(y_hat, thr, want_f32=False)
    pts = ops.bits_to_points(bits.cpu().numpy(), (SIZE, SIZE, SIZE), 4)
    worst = 0
    for j, r in enumerate(ref):
        diff = _as_set(pts[j]) ^ _as_set(r['points'])
        assert diff <= _as_set(r['fragile']), f'block {j}: {len(diff)} voxels differ beyond the fragile set'
        assert int(counts[j]) == len(pts[j])
        worst = max(worst, len(diff))

    # ---- the public decoder on the oracle's strings (blocks whose indexes agree) and on our own strings (all blocks)
    print(f'\n[{precision}] y flips {yflips}/{ysym.size}, z flips {zflips}/{zsym.size}, encoder index agreement min '
          f'{min(enc_idx_rate):.6f}, decoder index agreement min {min(dec_rate):.6f}, blocks with all indexes equal '
          f'{len(keep)}/{nb}, byte-identical strings {strings_equal}/{nb}, worst fragile-set diff {worst}')
    assert len(keep) >= min_keep * nb, f'only {len(keep)}/{nb} blocks reproduce every scale index of the oracle'
    m.decompress()
    dec, _ = m.decompress_blocks(None, [(ref[j]['strings'], 128) for j in keep], (SIZE, SIZE, SIZE))
    for j, p in zip(keep, dec):
        assert (_as_set(p) ^ _as_set(ref[j]['points'])) <= _as_set(ref[j]['fragile']), j
        assert p.dtype == np.float32 and p.tolist() == sorted(p.tolist())
    own, _ = m.decompress_blocks(None, data_list[0], (SIZE, SIZE, SIZE))
    for a, b in zip(meta[0]['x_hat_list'], own):
        assert np.array_equal(a, b)             # decompress_octree.py:94-119: decoder == encoder, exactly


@pytest.mark.parametrize('precision,tol', [('bf16x3', 3e-5), ('fp32', 1e-5)])
@pytest.mark.parametrize('name,filters,cin,shape', [
    ('AnalysisBlock', 16, 1, (32, 32, 32)), ('SynthesisBlock', 16, 32, (8, 16, 8)),
    ('AnalysisTransformV2', 16, 1, (32, 32, 32)), ('SynthesisTransformV2', 16, 16, (4, 4, 4)),
    ('AnalysisTransformProgressiveV2', 32, 1, (32, 32, 32)), ('SynthesisTransformProgressiveV2', 32, 32, (4, 4, 4))])
def test_residual_concat_values_match_oracle(name, filters, cin, shape, precision, tol):
    """ResidualLayer(residual_mode='concat') = concat((t, t1), channel axis) (src/model_transforms.py:37-38): values, not
    only shapes, against the oracle in float64, for the blocks and the three V2 transform families built with concat."""
    MT.set_precision(precision)
    rng = np.random.default_rng(11)
    spec = T.build_transform(name, filters, 'concat')
    w = T.init_weights(spec, cin, rng, bias_scale=0.05, dtype=torch.float64)
    # Glorot weights shrink the signal layer by layer: scale them up so that every layer's ReLU sees both signs at O(1)
    for l in w:
        l['kernel'] = l['kernel'] * 2.0
    cls = getattr(MT, name)
    layer = cls(filters, data_format='channels_first', residual_mode='concat')
    x = torch.from_numpy((rng.random((2, cin) + shape) < (0.1 if cin == 1 else 0.5)).astype(np.float32) * rng.normal(1.0, 0.3, (2, cin) + shape).astype(np.float32))
    if cin == 1:
        x = (x != 0).float()
    # lazily-built layers (Keras semantics): one call creates the variables with the concat channel counts, then set them
    y0 = layer(x.cuda())
    leaves = layer.leaf_layers()
    assert len(leaves) == len(w)
    for l, ww in zip(leaves, w):
        assert tuple(l.kernel.shape) == tuple(ww['kernel'].shape), (l.kernel.shape, ww['kernel'].shape)
        l.set_weights(ww['kernel'].numpy().astype(np.float32), None if ww['bias'] is None else ww['bias'].numpy().astype(np.float32))
    if hasattr(layer, '_trace'):
        del layer._trace
    w32 = [{'kernel': torch.from_numpy(l.kernel).double(), 'bias': None if l.bias is None else torch.from_numpy(l.bias).double()}
           for l in leaves]
    want = T.apply_transform(spec, w32, x.double())
    got = layer(x.cuda()).cpu().double()
    assert tuple(got.shape) == tuple(want.shape) == tuple(y0.shape)
    scale = float(want.abs().max())
    assert scale > 1e-3
    assert float((got - want).abs().max()) <= tol * scale, float((got - want).abs().max()) / scale
