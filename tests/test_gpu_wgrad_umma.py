"""tcgen05 weight gradient (conv3d_wgrad_umma.cu) of the stride-1 3x3x3 layers against float64 autograd on the oracle's
convolutions (small volumes) and against the fp32 CUDA-core kernel -- itself within 1e-5 of the oracle, test_gpu_train.py -- at the
sizes of the c3p training step (reference src/model_types.py:364-369: what Adam.minimize differentiates).

Tolerances: bf16x3 operands carry ~2^-16 relative error per product and the sum runs in fp32 inside TMEM: 2e-4 of the gradient's
scale; single bf16 (terms = 1): 2e-2."""
import numpy as np
import pytest
import torch

from oracle import transforms as T
from pcc_geo_cnn_v2_b200 import ops

pytestmark = pytest.mark.gpu
DT = torch.float64


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _umma(xd, gd, transposed, terms):
    n, c, d, h, w = xd.shape
    assert ops.wgrad_umma_eligible(n, c, c, 3, 1, d, h, w, terms)
    xb, gb = ops.f32_to_blocked(xd, terms), ops.f32_to_blocked(gd, terms)
    return ops.conv3d_wgrad_umma(xb, gb, tuple(xd.shape), transposed, terms)


@pytest.mark.parametrize('c,shape,n', [(16, (16, 16, 16), 2), (16, (5, 7, 16), 3), (32, (6, 16, 16), 2), (64, (4, 5, 16), 2), (16, (3, 4, 32), 1),
                                       (64, (4, 5, 8), 2), (32, (3, 8, 8), 3), (64, (8, 8, 8), 2)])   # W = 8: K chunks span two rows
@pytest.mark.parametrize('transposed', [False, True])
def test_wgrad_umma_matches_float64_autograd(c, shape, n, transposed):
    rng = np.random.default_rng(c + 7 * int(transposed) + shape[0])
    x = torch.tensor(rng.normal(size=(n, c) + shape), dtype=DT, requires_grad=True)
    kern = torch.tensor(rng.normal(size=(3, 3, 3, c, c)) / np.sqrt(27 * c), dtype=DT, requires_grad=True)
    y = (T.conv3d_transpose_same if transposed else T.conv3d_same)(x, kern, None, 1, False)
    gy = rng.normal(size=tuple(y.shape)).astype(np.float32)
    y.backward(torch.tensor(gy, dtype=DT))
    want = kern.grad.numpy().transpose(0, 1, 2, 4, 3) if transposed else kern.grad.numpy()
    xd, gd = x.detach().float().cuda(), torch.from_numpy(gy).cuda()
    got = _umma(xd, gd, transposed, 2).cpu().numpy().reshape(3, 3, 3, c, c)
    assert _rel(got, want) < 2e-4, _rel(got, want)
    got1 = _umma(xd, gd, transposed, 1).cpu().numpy().reshape(3, 3, 3, c, c)
    assert _rel(got1, want) < 2e-2, _rel(got1, want)


@pytest.mark.parametrize('c,s,n', [(16, 64, 32), (16, 32, 32), (32, 32, 32), (32, 16, 32), (64, 16, 32), (16, 64, 5), (64, 8, 32)])
def test_wgrad_umma_matches_fp32_kernel_at_training_sizes(c, s, n):
    g = torch.Generator(device='cuda').manual_seed(c * 100 + s)
    xd = torch.relu(torch.randn((n, c, s, s, s), device='cuda', generator=g))        # post-ReLU activations: non-negative, sparse
    gd = torch.randn((n, c, s, s, s), device='cuda', generator=g) * 1e-3
    want = ops.conv3d_wgrad_f32(xd, gd, c, 3, 1, False).cpu().numpy()
    got = _umma(xd, gd, False, 2).cpu().numpy()
    assert _rel(got, want) < 2e-4, _rel(got, want)
    again = _umma(xd, gd, False, 2).cpu().numpy()
    assert np.array_equal(got, again)            # deterministic: fixed reduction order


@pytest.mark.parametrize('transposed,cin,cout,s_small,n', [(True, 32, 16, 16, 3), (False, 16, 32, 16, 2), (True, 64, 32, 16, 2), (True, 32, 16, 32, 32),
                                                           (False, 32, 64, 8, 2), (True, 64, 64, 8, 3)])
def test_stride2_layers_by_phase_decomposition(transposed, cin, cout, s_small, n):
    """the trainer's stride-2 weight gradient (eight phase volumes of the large tensor as channels -> stride-1 tcgen05 launches) against
    the fp32 kernel, which test_gpu_train.py pins to float64 autograd"""
    from types import SimpleNamespace
    from pcc_geo_cnn_v2_b200.training import Trainer
    g = torch.Generator(device='cuda').manual_seed(cin + cout + s_small)
    s_in = s_small if transposed else 2 * s_small
    s_out = 2 * s_small if transposed else s_small
    xd = torch.relu(torch.randn((n, cin, s_in, s_in, s_in), device='cuda', generator=g))
    gd = torch.randn((n, cout, s_out, s_out, s_out), device='cuda', generator=g) * 1e-3
    layer = SimpleNamespace(transposed=transposed, filters=cout, k=3, stride=2)
    got = Trainer._wgrad_stride2(layer, xd, gd, None, 2)
    assert got is not None
    want = ops.conv3d_wgrad_f32(xd, gd, cout, 3, 2, transposed).cpu().numpy()
    assert _rel(got.cpu().numpy(), want) < 2e-4, _rel(got.cpu().numpy(), want)


def test_wgrad_umma_rejects_unsupported_geometry():
    assert not ops.wgrad_umma_eligible(2, 16, 32, 3, 1, 16, 16, 16)
    assert not ops.wgrad_umma_eligible(2, 16, 16, 3, 2, 16, 16, 16)
    assert not ops.wgrad_umma_eligible(2, 16, 16, 3, 1, 8, 8, 8)
    assert not ops.wgrad_umma_eligible(2, 64, 64, 3, 1, 4, 4, 4)
    assert not ops.wgrad_umma_eligible(2, 8, 8, 3, 1, 16, 16, 16)
    with pytest.raises(ValueError):
        ops.conv3d_wgrad_umma(torch.zeros(8, device='cuda', dtype=torch.bfloat16), torch.zeros(8, device='cuda', dtype=torch.bfloat16),
                              (1, 64, 4, 4, 4), False, 2)


@pytest.mark.parametrize('shape', [(3, 16, 5, 6, 8), (2, 12, 4, 4, 4), (32, 64, 8, 8, 8)])
@pytest.mark.parametrize('terms', [2, 1])
def test_blocked_backward_helpers(shape, terms):
    """ReLU mask, residual add and bias gradient on the blocked bf16 layout (what the tensor-core training path uses between the layers of
    its backward pass) against the fp32 kernels of the same steps"""
    g = torch.Generator(device='cuda').manual_seed(shape[1])
    gd = torch.randn(shape, device='cuda', generator=g)
    y = torch.relu(torch.randn(shape, device='cuda', generator=g))
    y[0, 0].zero_()
    gb, yb = ops.f32_to_blocked(gd, terms), ops.f32_to_blocked(y, terms)
    masked = ops.relu_mask_blocked(gb, yb, shape, terms)
    want = ops.f32_to_blocked(ops.relu_bwd(gd, y), terms)
    assert torch.equal(ops.blocked_to_f32(masked, shape, terms), ops.blocked_to_f32(want, shape, terms))
    other = torch.randn(shape, device='cuda', generator=g)
    ob = ops.f32_to_blocked(other, terms)
    tol = 2e-5 if terms == 2 else 2e-2
    s = ops.blocked_to_f32(ops.add_blocked(gb, ob, shape, terms), shape, terms)
    ref = ops.blocked_to_f32(gb, shape, terms) + ops.blocked_to_f32(ob, shape, terms)
    assert float((s - ref).abs().max()) <= tol * float(ref.abs().max())
    db = ops.bias_grad_blocked(gb, shape, terms).cpu().numpy()
    ref_db = ops.blocked_to_f32(gb, shape, terms).double().sum((0, 2, 3, 4)).cpu().numpy()
    assert db.shape == (shape[1],)
    assert np.abs(db - ref_db).max() <= 1e-5 * max(1.0, np.abs(ref_db).max())
