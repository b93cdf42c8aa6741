"""Parity of the convolution kernels with the oracle (torch-CPU restatement of Keras Conv3D/Conv3DTranspose
'same'), through the C ABI.  fp32 CUDA-core kernel: abs/rel 1e-5 of the output scale.  tcgen05 kernel: bf16x3
mode within 3e-5 of the output scale (fp32-class), bf16 mode within 2e-2."""
import numpy as np
import pytest
import torch

from oracle import transforms as T
from pcc_geo_cnn_v2_b200 import ops

pytestmark = pytest.mark.gpu


def _case(rng, transposed, k, s, cin, cout, shape, n=2):
    x = torch.from_numpy(rng.normal(size=(n, cin) + shape).astype(np.float32))
    x = x * (torch.rand_like(x) < 0.5)  # sparse-ish, like post-ReLU activations
    kshape = (k, k, k, cout, cin) if transposed else (k, k, k, cin, cout)
    kern = torch.from_numpy((rng.normal(size=kshape) / np.sqrt(k ** 3 * cin)).astype(np.float32))
    bias = torch.from_numpy(rng.normal(size=(cout,)).astype(np.float32) * 0.1)
    return x, kern, bias


def _tap_major(kern, transposed):
    k = kern.shape[0]
    w = kern.permute(0, 1, 2, 4, 3) if transposed else kern
    return w.reshape(k ** 3, w.shape[3], w.shape[4]).contiguous()


def _oracle(x, kern, bias, s, relu, transposed, res=None):
    fn = T.conv3d_transpose_same if transposed else T.conv3d_same
    y = fn(x.double(), kern.double(), None if bias is None else bias.double(), s, relu)
    return y if res is None else y + res.double()


CASES = [
    # transposed, k, stride, cin, cout, spatial
    (False, 3, 1, 16, 16, (8, 8, 8)), (False, 3, 2, 1, 16, (16, 16, 16)), (False, 3, 2, 16, 32, (8, 8, 8)),
    (False, 3, 1, 64, 64, (4, 4, 4)), (False, 9, 2, 1, 32, (16, 16, 16)), (False, 5, 2, 32, 32, (8, 8, 8)),
    (False, 3, 2, 4, 6, (7, 7, 7)), (False, 3, 1, 3, 5, (1, 1, 1)),
    (True, 3, 1, 16, 16, (8, 8, 8)), (True, 3, 2, 64, 32, (4, 4, 4)), (True, 3, 1, 16, 1, (16, 16, 16)),
    (True, 5, 2, 32, 32, (4, 4, 4)), (True, 9, 2, 32, 1, (8, 8, 8)), (True, 3, 2, 5, 3, (3, 3, 3)), (True, 9, 2, 2, 2, (1, 1, 1)),
]


@pytest.mark.parametrize('transposed,k,s,cin,cout,shape', CASES)
def test_direct_conv_matches_oracle(transposed, k, s, cin, cout, shape):
    rng = np.random.default_rng(hash((transposed, k, s, cin, cout)) % 2 ** 31)
    x, kern, bias = _case(rng, transposed, k, s, cin, cout, shape)
    want = _oracle(x, kern, bias, s, True, transposed)
    got = ops.conv3d_f32(x.cuda(), _tap_major(kern, transposed).cuda(), bias.cuda(), cout, k, s, transposed, True)
    assert tuple(got.shape) == tuple(want.shape)
    scale = float(want.abs().max()) + 1e-6
    assert float((got.cpu().double() - want).abs().max()) < 1e-5 * scale
    # no bias, no relu, residual
    res = torch.from_numpy(rng.normal(size=tuple(want.shape)).astype(np.float32))
    want2 = _oracle(x, kern, None, s, False, transposed, res)
    got2 = ops.conv3d_f32(x.cuda(), _tap_major(kern, transposed).cuda(), None, cout, k, s, transposed, False, res.cuda())
    assert float((got2.cpu().double() - want2).abs().max()) < 1e-5 * (float(want2.abs().max()) + 1e-6)


def test_blocked_layout_round_trip():
    x = torch.randn(3, 20, 4, 16, 8, device='cuda')
    for terms, tol in ((2, 2e-5), (1, 1e-2)):
        xb = ops.f32_to_blocked(x, terms)
        assert xb.numel() == terms * 3 * 32 * 4 * 16 * 8
        back = ops.blocked_to_f32(xb, tuple(x.shape), terms)
        assert float((back - x).abs().max()) <= tol * float(x.abs().max())


UMMA_CASES = [
    # transposed, cin, cout, (D,H,W), n
    (False, 16, 16, (4, 16, 8), 1), (False, 16, 16, (16, 16, 16), 2), (True, 16, 16, (8, 32, 16), 1),
    (False, 32, 32, (16, 16, 16), 2), (True, 32, 32, (5, 16, 8), 3), (False, 16, 32, (3, 16, 8), 1),
    (True, 32, 16, (20, 16, 24), 1), (False, 16, 16, (1, 16, 8), 1), (False, 16, 16, (40, 32, 32), 5),
]


@pytest.mark.parametrize('terms,tol', [(2, 3e-5), (1, 2e-2)])
@pytest.mark.parametrize('transposed,cin,cout,shape,n', UMMA_CASES)
def test_umma_conv_matches_oracle(transposed, cin, cout, shape, n, terms, tol):
    rng = np.random.default_rng(hash((transposed, cin, cout, shape, n)) % 2 ** 31)
    x, kern, bias = _case(rng, transposed, 3, 1, cin, cout, shape, n)
    res = torch.from_numpy(rng.normal(size=(n, cout) + shape).astype(np.float32))
    want = _oracle(x, kern, bias, 1, True, transposed, res)
    wp = ops.umma_pack_weights(_tap_major(kern, transposed).numpy(), cin, cout, 1, transposed, terms)
    xb = ops.f32_to_blocked(x.cuda(), terms)
    rb = ops.f32_to_blocked(res.cuda(), terms)
    yb, shp = ops.conv3d_umma(xb, tuple(x.shape), wp, bias.cuda(), cout, 1, transposed, True, terms, rb)
    got = ops.blocked_to_f32(yb, shp, terms)
    torch.cuda.synchronize()
    assert shp == tuple(want.shape)
    err = float((got.cpu().double() - want).abs().max())
    scale = float(want.abs().max())
    assert err < tol * scale, f'max err {err:.3e} vs scale {scale:.3e}'


@pytest.mark.parametrize('terms', [2, 1])
def test_umma_full_size_layer_matches_fp32_kernel(terms):
    """BASELINE-size check (64^3, 16->16, batch 4): tensor-core kernel against the fp32 CUDA-core kernel."""
    rng = np.random.default_rng(11)
    x, kern, bias = _case(rng, True, 3, 1, 16, 16, (64, 64, 64), 4)
    xd = x.cuda()
    ref = ops.conv3d_f32(xd, _tap_major(kern, True).cuda(), bias.cuda(), 16, 3, 1, True, True)
    wp = ops.umma_pack_weights(_tap_major(kern, True).numpy(), 16, 16, 1, True, terms)
    yb, shp = ops.conv3d_umma(ops.f32_to_blocked(xd, terms), tuple(x.shape), wp, bias.cuda(), 16, 1, True, True, terms)
    got = ops.blocked_to_f32(yb, shp, terms)
    err = float((got - ref).abs().max())
    assert err < (3e-5 if terms == 2 else 2e-2) * float(ref.abs().max())
    # deterministic: same launch twice -> identical bits
    yb2, _ = ops.conv3d_umma(ops.f32_to_blocked(xd, terms), tuple(x.shape), wp, bias.cuda(), 16, 1, True, True, terms)
    assert torch.equal(yb, yb2)


HL_CASES = [
    # transposed, cin, cout, (D,H,W), n
    (False, 16, 16, (4, 16, 8), 1), (True, 16, 16, (16, 16, 16), 2), (False, 12, 9, (3, 16, 8), 1), (True, 16, 16, (1, 16, 8), 1),
    (False, 16, 16, (40, 32, 32), 5), (True, 16, 16, (20, 16, 24), 3), (True, 16, 16, (64, 64, 64), 2),
]


@pytest.mark.parametrize('transposed,cin,cout,shape,n', HL_CASES)
def test_umma_hl_conv_matches_oracle(transposed, cin, cout, shape, n):
    """hi/lo-stacked form of the TMA kernel (two N=96 MMAs per tap; all four partial products): fp32-class accuracy vs the
    float64 oracle, with bias, ReLU and the fused residual; bit-identical from launch to launch; at least as accurate as the
    three-product kernel it replaces."""
    rng = np.random.default_rng(hash((transposed, cin, cout, shape, n)) % 2 ** 31)
    x, kern, bias = _case(rng, transposed, 3, 1, cin, cout, shape, n)
    res = torch.from_numpy(rng.normal(size=(n, cout) + shape).astype(np.float32))
    want = _oracle(x, kern, bias, 1, True, transposed, res)
    wp = ops.umma_hl_pack_weights(_tap_major(kern, transposed).numpy(), cin, cout, transposed)
    xb = ops.f32_to_blocked(x.cuda(), 2)
    rb = ops.f32_to_blocked(res.cuda(), 2)
    yb, shp = ops.conv3d_umma_hl(xb, tuple(x.shape), wp, bias.cuda(), cout, transposed, True, rb)
    got = ops.blocked_to_f32(yb, shp, 2)
    assert shp == tuple(want.shape)
    scale = float(want.abs().max())
    err = float((got.cpu().double() - want).abs().max())
    assert err < 3e-5 * scale, f'max err {err:.3e} vs scale {scale:.3e}'
    yb2, _ = ops.conv3d_umma_hl(xb, tuple(x.shape), wp, bias.cuda(), cout, transposed, True, rb)
    assert torch.equal(yb, yb2)
    w3 = ops.umma_pack_weights(_tap_major(kern, transposed).numpy(), cin, cout, 1, transposed, 2)
    y3, _ = ops.conv3d_umma(xb, tuple(x.shape), w3, bias.cuda(), cout, 1, transposed, True, 2, rb)
    err3 = float((ops.blocked_to_f32(y3, shp, 2).cpu().double() - want).abs().max())
    assert err <= 1.5 * err3 + 1e-7 * scale, (err, err3)
    # no bias / no relu / no residual
    want2 = _oracle(x, kern, None, 1, False, transposed)
    y2, _ = ops.conv3d_umma_hl(xb, tuple(x.shape), wp, None, cout, transposed, False)
    assert float((ops.blocked_to_f32(y2, shp, 2).cpu().double() - want2).abs().max()) < 3e-5 * float(want2.abs().max())


ZY_CASES = [
    # transposed, cin, cout, (D,H,W), n
    (False, 16, 16, (4, 16, 8), 1), (True, 16, 16, (16, 16, 16), 2), (False, 12, 9, (3, 5, 8), 3), (True, 16, 16, (1, 1, 8), 1),
    (False, 16, 16, (2, 23, 16), 17), (True, 16, 16, (7, 64, 24), 16), (True, 16, 16, (20, 16, 24), 3), (False, 16, 16, (64, 64, 64), 32),
    (True, 16, 16, (5, 3, 8), 33),
    # small batches: M tiles of 8 blocks x 16 x / 4 blocks x 32 x (the 128^3 blocks run at batch 4)
    (False, 16, 16, (4, 12, 32), 4), (True, 16, 16, (3, 8, 64), 5), (False, 16, 16, (2, 10, 128), 4), (True, 16, 16, (3, 9, 32), 8),
]


@pytest.mark.parametrize('terms,tol', [(2, 3e-5), (1, 2e-2)])
@pytest.mark.parametrize('transposed,cin,cout,shape,n', ZY_CASES)
def test_umma_zy_conv_matches_oracle(transposed, cin, cout, shape, n, terms, tol):
    """zy-ring kernel (y and z taps accumulated in a 2-D TMEM ring, N = 144 MMAs, tiles of 16 blocks x 8 x x <= 10 rows): fp32-class
    accuracy vs the float64 oracle with bias, ReLU and the fused residual, over ragged batches (not a multiple of 16), ragged row
    tiles, single-plane / single-row volumes and the headline shape; bit-identical from launch to launch."""
    rng = np.random.default_rng(hash((transposed, cin, cout, shape, n)) % 2 ** 31)
    big = n * np.prod(shape) > 2 ** 21
    x, kern, bias = _case(rng, transposed, 3, 1, cin, cout, shape, n)
    res = torch.from_numpy(rng.normal(size=(n, cout) + shape).astype(np.float32))
    wp = ops.umma_zy_pack_weights(_tap_major(kern, transposed).numpy(), cin, cout, transposed, terms)
    xb = ops.f32_to_blocked(x.cuda(), terms)
    rb = ops.f32_to_blocked(res.cuda(), terms)
    yb, shp = ops.conv3d_umma_zy(xb, tuple(x.shape), wp, bias.cuda(), cout, True, terms, rb)
    got = ops.blocked_to_f32(yb, shp, terms)
    if big:   # the headline shape: the fp32 CUDA-core kernel is the checker (the float64 CPU oracle would take minutes)
        want = ops.conv3d_f32(x.cuda(), _tap_major(kern, transposed).cuda(), bias.cuda(), cout, 3, 1, transposed, True, res.cuda()).cpu().double()
    else:
        want = _oracle(x, kern, bias, 1, True, transposed, res)
    assert shp == tuple(want.shape)
    scale = float(want.abs().max())
    err = float((got.cpu().double() - want).abs().max())
    assert err < tol * scale, f'max err {err:.3e} vs scale {scale:.3e}'
    yb2, _ = ops.conv3d_umma_zy(xb, tuple(x.shape), wp, bias.cuda(), cout, True, terms, rb)
    assert torch.equal(yb, yb2)
    if not big:   # no bias / no relu / no residual
        want2 = _oracle(x, kern, None, 1, False, transposed)
        y2, _ = ops.conv3d_umma_zy(xb, tuple(x.shape), wp, None, cout, False, terms)
        assert float((ops.blocked_to_f32(y2, shp, terms).cpu().double() - want2).abs().max()) < tol * float(want2.abs().max())


UMMA_UP2_CASES = [
    # cin, cout, (D,H,W) of the input, n
    (32, 16, (4, 16, 8), 1), (32, 16, (16, 16, 16), 2), (16, 16, (1, 16, 8), 1), (32, 16, (7, 32, 24), 3), (24, 12, (5, 16, 8), 2),
    (32, 16, (32, 32, 32), 2),
]


@pytest.mark.parametrize('terms,tol', [(2, 3e-5), (1, 2e-2)])
@pytest.mark.parametrize('cin,cout,shape,n', UMMA_UP2_CASES)
def test_umma_stride2_transposed_matches_oracle(cin, cout, shape, n, terms, tol):
    """SynthesisBlock's first layer (Conv3DTranspose k3 s2 'same', model_transforms.py:78) on the TMA kernel's UP=2 form."""
    rng = np.random.default_rng(hash((cin, cout, shape, n)) % 2 ** 31)
    x, kern, bias = _case(rng, True, 3, 2, cin, cout, shape, n)
    want = _oracle(x, kern, bias, 2, True, True)
    wp = ops.umma_pack_weights(_tap_major(kern, True).numpy(), cin, cout, 2, True, terms)
    xb = ops.f32_to_blocked(x.cuda(), terms)
    yb, shp = ops.conv3d_umma(xb, tuple(x.shape), wp, bias.cuda(), cout, 2, True, True, terms)
    got = ops.blocked_to_f32(yb, shp, terms)
    torch.cuda.synchronize()
    assert shp == tuple(want.shape)
    err = float((got.cpu().double() - want).abs().max())
    scale = float(want.abs().max())
    assert err < tol * scale, f'max err {err:.3e} vs scale {scale:.3e}'
    # no bias, no relu; deterministic
    yb2, _ = ops.conv3d_umma(xb, tuple(x.shape), wp, None, cout, 2, True, False, terms)
    want2 = _oracle(x, kern, None, 2, False, True)
    assert float((ops.blocked_to_f32(yb2, shp, terms).cpu().double() - want2).abs().max()) < tol * float(want2.abs().max())
    yb3, _ = ops.conv3d_umma(xb, tuple(x.shape), wp, None, cout, 2, True, False, terms)
    assert torch.equal(yb2, yb3)


UMMA_YS_CASES = [
    # transposed, cin, cout, (D,H,W), n
    (False, 16, 16, (4, 28, 8), 1), (True, 16, 16, (8, 32, 16), 2), (False, 16, 16, (1, 64, 8), 1), (True, 16, 16, (2, 40, 24), 3),
    (False, 12, 10, (5, 30, 8), 2), (True, 16, 16, (11, 14, 8), 1), (True, 16, 16, (64, 64, 64), 2),
]


@pytest.mark.parametrize('terms,tol', [(2, 3e-5), (1, 2e-2)])
@pytest.mark.parametrize('transposed,cin,cout,shape,n', UMMA_YS_CASES)
def test_umma_ystacked_conv_matches_oracle(transposed, cin, cout, shape, n, terms, tol):
    """16-filter stride-1 layers of AnalysisBlock / SynthesisBlock (model_transforms.py:62-81) on the y-stacked kernel,
    incl. heights that are not multiples of its 14-row tile, fused bias + ReLU + residual."""
    rng = np.random.default_rng(hash((transposed, cin, cout, shape, n)) % 2 ** 31)
    x, kern, bias = _case(rng, transposed, 3, 1, cin, cout, shape, n)
    res = torch.from_numpy(rng.normal(size=(n, cout) + shape).astype(np.float32))
    want = _oracle(x, kern, bias, 1, True, transposed, res)
    wp = ops.umma_ys_pack_weights(_tap_major(kern, transposed).numpy(), cin, cout, transposed, terms)
    xb = ops.f32_to_blocked(x.cuda(), terms)
    rb = ops.f32_to_blocked(res.cuda(), terms)
    yb, shp = ops.conv3d_umma_ys(xb, tuple(x.shape), wp, bias.cuda(), cout, True, terms, rb)
    got = ops.blocked_to_f32(yb, shp, terms)
    torch.cuda.synchronize()
    err = float((got.cpu().double() - want).abs().max())
    scale = float(want.abs().max())
    assert err < tol * scale, f'max err {err:.3e} vs scale {scale:.3e}'
    yb2, _ = ops.conv3d_umma_ys(xb, tuple(x.shape), wp, None, cout, False, terms)
    want2 = _oracle(x, kern, None, 1, False, transposed)
    assert float((ops.blocked_to_f32(yb2, shp, terms).cpu().double() - want2).abs().max()) < tol * float(want2.abs().max())
    yb3, _ = ops.conv3d_umma_ys(xb, tuple(x.shape), wp, None, cout, False, terms)
    assert torch.equal(yb2, yb3)   # deterministic


OUT1_CASES = [
    # transposed, cin, (D,H,W), n
    (True, 16, (4, 16, 8), 1), (True, 16, (16, 16, 16), 2), (False, 16, (5, 32, 24), 2), (True, 16, (1, 16, 8), 1),
    (True, 16, (2, 16, 8), 3), (True, 12, (9, 16, 16), 1), (True, 16, (64, 64, 64), 2),
]


@pytest.mark.parametrize('terms,tol', [(2, 3e-5), (1, 2e-2)])
@pytest.mark.parametrize('transposed,cin,shape,n', OUT1_CASES)
def test_out1_conv_matches_oracle_and_packs_bits(transposed, cin, shape, n, terms, tol):
    """Last synthesis layer (model_transforms.py:107,135) + clip/threshold/pack (model_types.py:201-202,233-234)."""
    rng = np.random.default_rng(hash((transposed, cin, shape, n)) % 2 ** 31)
    x, kern, bias = _case(rng, transposed, 3, 1, cin, 1, shape, n)
    bias = bias + 0.3
    want = _oracle(x, kern, bias, 1, True, transposed)
    wp = ops.out1_pack_weights(_tap_major(kern, transposed).numpy(), cin, transposed, terms)
    xb = ops.f32_to_blocked(x.cuda(), terms)
    thr = torch.from_numpy(rng.uniform(0.1, 0.9, size=n).astype(np.float32)).cuda()
    xh, bits, counts = ops.conv3d_out1(xb, tuple(x.shape), wp, bias.cuda(), True, terms, True, thr)
    torch.cuda.synchronize()
    assert tuple(xh.shape) == tuple(want.shape)
    err = float((xh.cpu().double() - want).abs().max())
    scale = float(want.abs().max())
    assert err < tol * scale, f'max err {err:.3e} vs scale {scale:.3e}'
    # the packed output is exactly threshold_pack of the kernel's own x_hat (bit-exact integer work)
    bits2, counts2 = ops.threshold_pack(xh, thr)
    assert torch.equal(bits, bits2) and torch.equal(counts, counts2)
    # bits-only and f32-only calls agree with the combined one; deterministic
    _, bits3, counts3 = ops.conv3d_out1(xb, tuple(x.shape), wp, bias.cuda(), True, terms, False, thr)
    xh3, _, _ = ops.conv3d_out1(xb, tuple(x.shape), wp, bias.cuda(), True, terms, True, None)
    assert torch.equal(bits, bits3) and torch.equal(counts, counts3) and torch.equal(xh, xh3)
    # no bias / no relu
    xh4, _, _ = ops.conv3d_out1(xb, tuple(x.shape), wp, None, False, terms, True, None)
    want4 = _oracle(x, kern, None, 1, False, transposed)
    assert float((xh4.cpu().double() - want4).abs().max()) < tol * float(want4.abs().max())


GEMM_CASES = [
    # transposed, k, stride, cin, cout, (D,H,W), n
    (False, 3, 1, 64, 64, (8, 8, 8), 3), (False, 3, 1, 64, 64, (4, 4, 4), 5), (True, 3, 1, 64, 64, (16, 16, 16), 1),
    (False, 3, 2, 16, 32, (16, 16, 16), 2), (False, 3, 2, 32, 64, (8, 8, 8), 3), (False, 3, 2, 64, 64, (8, 8, 8), 2),
    (True, 3, 2, 64, 64, (4, 4, 4), 3), (True, 3, 2, 64, 32, (8, 8, 8), 2), (True, 3, 2, 32, 16, (16, 16, 16), 1),
    (True, 3, 1, 16, 1, (8, 16, 8), 2), (False, 5, 2, 32, 32, (8, 8, 8), 2), (True, 5, 2, 32, 32, (4, 4, 4), 2),
    (True, 9, 2, 32, 1, (4, 4, 4), 1), (False, 3, 1, 16, 16, (3, 5, 7), 2), (True, 3, 2, 16, 16, (3, 5, 7), 1),
]


@pytest.mark.parametrize('terms,tol', [(2, 3e-5), (1, 2e-2)])
@pytest.mark.parametrize('transposed,k,s,cin,cout,shape,n', GEMM_CASES)
def test_gemm_conv_matches_oracle(transposed, k, s, cin, cout, shape, n, terms, tol):
    rng = np.random.default_rng(hash((transposed, k, s, cin, cout, shape, n)) % 2 ** 31)
    x, kern, bias = _case(rng, transposed, k, s, cin, cout, shape, n)
    want0 = _oracle(x, kern, bias, s, True, transposed)
    res = torch.from_numpy(rng.normal(size=tuple(want0.shape)).astype(np.float32))
    want = want0 + res.double()
    wimg = ops.gemm_pack_weights(_tap_major(kern, transposed).numpy(), cin, cout, k, s, transposed, terms)
    xb = ops.f32_to_blocked(x.cuda(), terms)
    rb = ops.f32_to_blocked(res.cuda(), terms)
    yb, shp = ops.conv3d_gemm(xb, tuple(x.shape), wimg, bias.cuda(), cout, s, transposed, True, terms, rb)
    got = ops.blocked_to_f32(yb, shp, terms)
    torch.cuda.synchronize()
    assert shp == tuple(want.shape)
    err = float((got.cpu().double() - want).abs().max())
    scale = float(want.abs().max())
    assert err < tol * scale, f'max err {err:.3e} vs scale {scale:.3e}'
    # no bias / relu / residual
    yb2, _ = ops.conv3d_gemm(xb, tuple(x.shape), wimg, None, cout, s, transposed, False, terms)
    got2 = ops.blocked_to_f32(yb2, shp, terms)
    want2 = _oracle(x, kern, None, s, False, transposed)
    assert float((got2.cpu().double() - want2).abs().max()) < tol * float(want2.abs().max())


@pytest.mark.parametrize('terms,tol', [(2, 1e-5), (1, 1e-2)])
@pytest.mark.parametrize('cout,shape,n', [(16, (16, 16, 16), 2), (16, (64, 64, 64), 3), (12, (8, 6, 4), 5), (16, (2, 2, 2), 1)])
def test_fused_first_layer_matches_oracle(cout, shape, n, terms, tol):
    """Conv3D(F, 3, strides 2, 'same') + bias + ReLU on the one-channel occupancy volume, written straight into the blocked
    bf16 layout (conv3d_first.cu), against the float64 oracle -- on occupancy-like {0,1} input and on a dense real-valued one."""
    rng = np.random.default_rng(cout + shape[0])
    for dense in (False, True):
        x = torch.from_numpy(rng.normal(size=(n, 1) + shape).astype(np.float32)) if dense else \
            torch.from_numpy((rng.random((n, 1) + shape) < 0.05).astype(np.float32))
        kern = torch.from_numpy((rng.normal(size=(3, 3, 3, 1, cout)) / np.sqrt(27)).astype(np.float32))
        bias = torch.from_numpy(rng.normal(size=(cout,)).astype(np.float32) * 0.1)
        want = _oracle(x, kern, bias, 2, True, False)
        yb, shp = ops.conv3d_first(x.cuda(), _tap_major(kern, False).cuda(), bias.cuda(), cout, True, terms)
        got = ops.blocked_to_f32(yb, shp, terms)
        assert shp == tuple(want.shape)
        scale = float(want.abs().max()) + 1e-6
        assert float((got.cpu().double() - want).abs().max()) < tol * scale
