"""Pins the oracle's transform restatement: the reference's own shape tests (src/test_model_transforms.py:27-73,
ported), TF 'SAME' padding known answers computed by hand, and the conv / transposed-conv adjoint identity."""
import numpy as np
import pytest
import torch

from oracle import transforms as T


def _run(name, filters, x_shape_cl, mode='add'):
    """reference tests feed channels_last zeros; the oracle is channels_first."""
    n, d, h, w, c = x_shape_cl
    spec = T.build_transform(name, filters, mode)
    weights = T.init_weights(spec, c, np.random.default_rng(0))
    y = T.apply_transform(spec, weights, torch.zeros((n, c, d, h, w)))
    return tuple(y.permute(0, 2, 3, 4, 1).shape)


X, Y = (1, 8, 8, 8, 1), (1, 1, 1, 1, 1)


@pytest.mark.parametrize('name,filters,inp,mode,expect', [
    ('AnalysisTransformV1', 1, X, 'add', (1, 1, 1, 1, 1)),                  # test_model_transforms.py:27-29
    ('SynthesisTransformV1', 2, Y, 'add', (1, 8, 8, 8, 1)),                 # :31-33
    ('AnalysisBlock', 1, X, 'add', (1, 4, 4, 4, 1)),                        # :35-37
    ('AnalysisBlock', 1, X, 'concat', (1, 4, 4, 4, 2)),                     # :38-39
    ('SynthesisBlock', 1, Y, 'add', (1, 2, 2, 2, 1)),                       # :41-43
    ('SynthesisBlock', 1, Y, 'concat', (1, 2, 2, 2, 2)),                    # :44-45
    ('AnalysisTransformV2', 2, X, 'add', (1, 1, 1, 1, 2)),                  # :47-49
    ('AnalysisTransformV2', 2, X, 'concat', (1, 1, 1, 1, 2)),               # :50-51
    ('SynthesisTransformV2', 2, Y, 'add', (1, 8, 8, 8, 1)),                 # :53-55
    ('SynthesisTransformV2', 2, Y, 'concat', (1, 8, 8, 8, 1)),              # :56-57
    ('AnalysisTransformProgressiveV2', 4, X, 'add', (1, 1, 1, 1, 4)),       # :59-61
    ('SynthesisTransformProgressiveV2', 4, Y, 'add', (1, 8, 8, 8, 1)),      # :63-65
    ('HyperAnalysisTransform', 1, X, 'add', (1, 4, 4, 4, 1)),               # :67-69
    ('HyperSynthesisTransform', 1, Y, 'add', (1, 2, 2, 2, 1)),              # :71-73
])
def test_reference_shape_contracts(name, filters, inp, mode, expect):
    assert _run(name, filters, inp, mode) == expect


def test_same_pads_known_answers():
    # TF SAME: out=ceil(n/s), total=max((out-1)s+k-n,0), before=total//2  
    assert T.same_pads(64, 3, 1) == (64, 1, 1)
    assert T.same_pads(64, 3, 2) == (32, 0, 1)
    assert T.same_pads(64, 5, 2) == (32, 1, 2)
    assert T.same_pads(64, 9, 2) == (32, 3, 4)
    assert T.same_pads(1, 3, 2) == (1, 1, 1)
    assert T.same_pads(7, 3, 2) == (4, 1, 1)


def test_conv_stride2_alignment_known_answer():
    """k3 s2 on even sizes pads (0,1): out[o] = sum_k in[2o+k] w[k] -- a delta at input index 2 with an all-ones
    kernel must light outputs 0 (tap 2) and 1 (tap 0) only."""
    x = torch.zeros(1, 1, 8, 8, 8)
    x[0, 0, 2, 2, 2] = 1.0
    k = torch.ones(3, 3, 3, 1, 1)
    y = T.conv3d_same(x, k, None, 2, False)[0, 0]
    assert y.shape == (4, 4, 4)
    nz = {tuple(i) for i in torch.nonzero(y).tolist()}
    assert nz == {(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)}


def test_conv_transpose_stride2_subpixel_known_answer():
    """Conv3DTranspose k3 s2 SAME: out[2i+k] += in[i] w[k] (crop 0 front / 1 back).  A delta at i=1 with
    w = (1,10,100) per axis gives out[2]=1, out[3]=10, out[4]=100 along each axis."""
    x = torch.zeros(1, 1, 4, 4, 4)
    x[0, 0, 1, 1, 1] = 1.0
    w1 = torch.tensor([1.0, 10.0, 100.0])
    k = (w1[:, None, None] * w1[None, :, None] * w1[None, None, :]).reshape(3, 3, 3, 1, 1)
    y = T.conv3d_transpose_same(x, k, None, 2, False)[0, 0]
    assert y.shape == (8, 8, 8)
    assert y[2, 2, 2] == 1 and y[3, 2, 2] == 10 and y[4, 2, 2] == 100 and y[4, 4, 4] == 1e6
    assert y[1].abs().sum() == 0 and y[5].abs().sum() == 0
    # last input voxel: tap 2 would land at 2*3+2 = 8 -> cropped
    x2 = torch.zeros(1, 1, 4, 4, 4)
    x2[0, 0, 3, 3, 3] = 1.0
    y2 = T.conv3d_transpose_same(x2, k, None, 2, False)[0, 0]
    assert y2[6, 6, 6] == 1 and y2[7, 7, 7] == 1000


@pytest.mark.parametrize('k,s', [(3, 1), (3, 2), (5, 2), (9, 2)])
def test_transpose_is_adjoint_of_conv(k, s):
    g = torch.Generator().manual_seed(k * 10 + s)
    x = torch.randn(2, 3, 4, 4, 4, dtype=torch.float64, generator=g)
    kern = torch.randn(k, k, k, 2, 3, dtype=torch.float64, generator=g)
    yt = T.conv3d_transpose_same(x, kern, None, s, False)
    y = torch.randn(yt.shape, dtype=torch.float64, generator=g)
    c = T.conv3d_same(y, kern, None, s, False)
    assert abs(float((yt * y).sum() - (x * c).sum())) < 1e-9


def test_residual_adds_post_relu_tensors():
    """ResidualLayer: out = relu(conv0(x)) + relu(conv2(relu(conv1(.)))) -- no activation after the add."""
    spec = T.build_transform('AnalysisBlock', 2)
    w = T.init_weights(spec, 1, np.random.default_rng(3), bias_scale=0.3)
    x = torch.rand(1, 1, 8, 8, 8)
    t1 = T.conv3d_same(x, w[0]['kernel'], w[0]['bias'], 2, True)
    t = T.conv3d_same(t1, w[1]['kernel'], w[1]['bias'], 1, True)
    t = T.conv3d_same(t, w[2]['kernel'], w[2]['bias'], 1, True)
    assert torch.allclose(T.apply_transform(spec, w, x), t1 + t)
