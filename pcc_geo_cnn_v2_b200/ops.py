"""Thin, allocation-explicit wrappers over the C ABI (include/pccgeo.h).  torch is used only for device
memory and the current stream; every compute step is a libpccgeo kernel."""
import numpy as np
import torch

from . import _lib as L


def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), 'expects a contiguous fp32 CUDA tensor'
    return t


def same_out(n, stride, transposed):
    return n * stride if transposed else -(-n // stride)


def conv3d_f32(x, w_tap, bias, cout, k, stride, transposed, relu, residual=None, out=None):
    """Keras Conv3D / Conv3DTranspose ('same') + bias + ReLU + residual on fp32 (N,C,D,H,W).
    w_tap: device fp32 (k^3, Cin, Cout)."""
    L.require_cuda()
    _f32c(x)
    n, cin, d, h, w = x.shape
    shp = (n, cout, same_out(d, stride, transposed), same_out(h, stride, transposed), same_out(w, stride, transposed))
    if out is None:
        out = torch.empty(shp, device=x.device, dtype=torch.float32)
    assert tuple(out.shape) == shp and out.is_contiguous()
    if residual is not None:
        assert tuple(residual.shape) == shp
        _f32c(residual)
    L.check(L.lib().pccgeo_conv3d_f32(L.ptr(x), L.ptr(w_tap), L.ptr(bias), L.ptr(residual), L.ptr(out),
                                      n, cin, d, h, w, cout, k, stride, int(transposed), int(relu), L.stream_ptr()),
            'conv3d_f32')
    return out


# ---- blocked bf16 layout for the tcgen05 path ---------------------------------------------------------
def round_up(a, m):
    return -(-a // m) * m


def blocked_numel(n, c, d, h, w, terms):
    return terms * n * round_up(c, 16) * d * h * w


def upload(arr):
    """numpy -> device through pinned staging, without blocking the host: a pageable `.cuda()` copy first waits for everything queued
    on the stream, which stalls a training step once per packed weight image.  torch's pinned-memory cache recycles the
    staging block after the copy has run."""
    L.require_cuda()
    src = torch.from_numpy(np.ascontiguousarray(arr))
    pin = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    pin.copy_(src)
    return pin.to('cuda', non_blocking=True)


def f32_to_blocked(x, terms, out=None):
    L.require_cuda()
    _f32c(x)
    n, c, d, h, w = x.shape
    if out is None:
        out = torch.empty(blocked_numel(n, c, d, h, w, terms), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_f32_to_blocked(L.ptr(x), L.ptr(out), n, c, d, h, w, terms, L.stream_ptr()), 'f32_to_blocked')
    return out


def f32_phases_to_blocked(x, c, terms):
    """x fp32 (N, Cb, 2d, 2h, 2w) -> [blocked bf16 image of shape (N, c, d, h, w)] * (8*Cb/c): the eight stride-2 phase volumes of x as
    channels (phase-major, then source channel), cut into chunks of c channels.  One pass, no intermediate copies."""
    L.require_cuda()
    _f32c(x)
    n, cb, d2, h2, w2 = x.shape
    d, h, w = d2 // 2, h2 // 2, w2 // 2
    nch = 8 * cb // c
    per = blocked_numel(n, c, d, h, w, terms)
    out = torch.empty(nch * per, device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_f32_phases_to_blocked(L.ptr(x), L.ptr(out), n, cb, c, d, h, w, terms, L.stream_ptr()), 'f32_phases_to_blocked')
    return [out[j * per:(j + 1) * per] for j in range(nch)]


def blocked_to_f32(xb, shape, terms, out=None):
    L.require_cuda()
    n, c, d, h, w = shape
    if out is None:
        out = torch.empty(shape, device=xb.device, dtype=torch.float32)
    L.check(L.lib().pccgeo_blocked_to_f32(L.ptr(xb), L.ptr(out), n, c, d, h, w, terms, L.stream_ptr()), 'blocked_to_f32')
    return out


def conv3d_first(x, w_tap, bias, cout, relu, terms, out=None):
    """One-channel fp32 volume (N,1,D,H,W) -> stride-2 3x3x3 conv + bias + ReLU in the blocked bf16 layout.  Returns (yb, out_shape)."""
    L.require_cuda()
    x = _f32c(x)
    n, c, d, h, w = x.shape
    assert c == 1
    shp = (n, cout, d // 2, h // 2, w // 2)
    if out is None:
        out = torch.empty(blocked_numel(*shp, terms), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_conv3d_first(L.ptr(x), L.ptr(w_tap), L.ptr(bias), L.ptr(out), n, d, h, w, cout, int(relu), terms, L.stream_ptr()),
            'conv3d_first')
    return out, shp


def umma_pack_weights(w_tap_host, cin, cout, stride, transposed, terms):
    """numpy fp32 (27, Cin, Cout) -> device uint8 image for pccgeo_conv3d_umma."""
    w = np.ascontiguousarray(w_tap_host, np.float32)
    size = L.lib().pccgeo_umma_pack_weights_host(L.ptr(w), None, cin, cout, stride, int(transposed), terms)
    if size <= 0:
        L.check(int(size) if size < 0 else -1, 'umma_pack_weights')
    img = np.zeros(size, np.uint8)
    rc = L.lib().pccgeo_umma_pack_weights_host(L.ptr(w), L.ptr(img), cin, cout, stride, int(transposed), terms)
    if rc < 0:
        L.check(int(rc), 'umma_pack_weights')
    return upload(img)


def conv3d_umma(xb, in_shape, wpacked, bias, cout, stride, transposed, relu, terms, residual_b=None, out=None):
    """Blocked-layout tensor-core conv.  in_shape = logical (N, Cin, D, H, W).  Returns (yb, out_shape)."""
    L.require_cuda()
    n, cin, d, h, w = in_shape
    shp = (n, cout, same_out(d, stride, transposed), same_out(h, stride, transposed), same_out(w, stride, transposed))
    if out is None:
        out = torch.empty(blocked_numel(*shp, terms), device=xb.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_conv3d_umma(L.ptr(xb), L.ptr(wpacked), L.ptr(bias), L.ptr(residual_b), L.ptr(out),
                                       n, cin, d, h, w, cout, stride, int(transposed), int(relu), terms,
                                       L.stream_ptr()), 'conv3d_umma')
    return out, shp


def umma_hl_pack_weights(w_tap_host, cin, cout, transposed):
    """numpy fp32 (27, Cin, Cout) -> device uint8 image for pccgeo_conv3d_umma_hl (hi/lo weight halves stacked in N)."""
    w = np.ascontiguousarray(w_tap_host, np.float32)
    size = L.lib().pccgeo_umma_hl_pack_weights_host(L.ptr(w), None, cin, cout, int(transposed))
    if size < 0:
        L.check(int(size), 'umma_hl_pack_weights')
    img = np.zeros(int(size), np.uint8)
    rc = L.lib().pccgeo_umma_hl_pack_weights_host(L.ptr(w), L.ptr(img), cin, cout, int(transposed))
    if rc < 0:
        L.check(int(rc), 'umma_hl_pack_weights')
    return upload(img)


def conv3d_umma_hl(xb, in_shape, wpacked, bias, cout, transposed, relu, residual_b=None, out=None):
    """Two-term blocked-layout conv, stride 1, <= 16 channels in and out, on the hi/lo-stacked kernel.  Returns (yb, out_shape)."""
    L.require_cuda()
    n, cin, d, h, w = in_shape
    shp = (n, cout, d, h, w)
    if out is None:
        out = torch.empty(blocked_numel(*shp, 2), device=xb.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_conv3d_umma_hl(L.ptr(xb), L.ptr(wpacked), L.ptr(bias), L.ptr(residual_b), L.ptr(out),
                                          n, cin, d, h, w, cout, int(transposed), int(relu), L.stream_ptr()), 'conv3d_umma_hl')
    return out, shp


def umma_zy_pack_weights(w_tap_host, cin, cout, transposed, terms):
    """numpy fp32 (27, Cin, Cout) -> device uint8 image for pccgeo_conv3d_umma_zy (three z-rotations of the y/z-stacked taps)."""
    w = np.ascontiguousarray(w_tap_host, np.float32)
    size = L.lib().pccgeo_umma_zy_pack_weights_host(L.ptr(w), None, cin, cout, int(transposed), terms)
    if size < 0:
        L.check(int(size), 'umma_zy_pack_weights')
    img = np.zeros(int(size), np.uint8)
    rc = L.lib().pccgeo_umma_zy_pack_weights_host(L.ptr(w), L.ptr(img), cin, cout, int(transposed), terms)
    if rc < 0:
        L.check(int(rc), 'umma_zy_pack_weights')
    return upload(img)


def conv3d_umma_zy(xb, in_shape, wpacked, bias, cout, relu, terms, residual_b=None, out=None):
    """Blocked-layout conv, stride 1, <= 16 channels in and out, on the zy-ring kernel.  Returns (yb, out_shape)."""
    L.require_cuda()
    n, cin, d, h, w = in_shape
    shp = (n, cout, d, h, w)
    if out is None:
        out = torch.empty(blocked_numel(*shp, terms), device=xb.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_conv3d_umma_zy(L.ptr(xb), L.ptr(wpacked), L.ptr(bias), L.ptr(residual_b), L.ptr(out),
                                          n, cin, d, h, w, cout, int(relu), terms, L.stream_ptr()), 'conv3d_umma_zy')
    return out, shp


def umma_ys_pack_weights(w_tap_host, cin, cout, transposed, terms):
    """numpy fp32 (27, Cin, Cout) -> device uint8 image for pccgeo_conv3d_umma_ys."""
    w = np.ascontiguousarray(w_tap_host, np.float32)
    size = L.lib().pccgeo_umma_ys_pack_weights_host(L.ptr(w), None, cin, cout, int(transposed), terms)
    if size <= 0:
        L.check(int(size) if size < 0 else -1, 'umma_ys_pack_weights')
    img = np.zeros(size, np.uint8)
    rc = L.lib().pccgeo_umma_ys_pack_weights_host(L.ptr(w), L.ptr(img), cin, cout, int(transposed), terms)
    if rc < 0:
        L.check(int(rc), 'umma_ys_pack_weights')
    return upload(img)


def conv3d_umma_ys(xb, in_shape, wpacked, bias, cout, relu, terms, residual_b=None, out=None):
    """y-stacked blocked-layout tensor-core conv (stride 1, <= 16 channels).  Returns (yb, out_shape)."""
    L.require_cuda()
    n, cin, d, h, w = in_shape
    shp = (n, cout, d, h, w)
    if out is None:
        out = torch.empty(blocked_numel(*shp, terms), device=xb.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_conv3d_umma_ys(L.ptr(xb), L.ptr(wpacked), L.ptr(bias), L.ptr(residual_b), L.ptr(out),
                                          n, cin, d, h, w, cout, int(relu), terms, L.stream_ptr()), 'conv3d_umma_ys')
    return out, shp


def out1_pack_weights(w_tap_host, cin, transposed, terms):
    """numpy fp32 (27, Cin, 1) -> device uint8 image for pccgeo_conv3d_out1."""
    w = np.ascontiguousarray(w_tap_host, np.float32)
    size = L.lib().pccgeo_out1_pack_weights_host(L.ptr(w), None, cin, int(transposed), terms)
    if size <= 0:
        L.check(int(size) if size < 0 else -1, 'out1_pack_weights')
    img = np.zeros(size, np.uint8)
    rc = L.lib().pccgeo_out1_pack_weights_host(L.ptr(w), L.ptr(img), cin, int(transposed), terms)
    if rc < 0:
        L.check(int(rc), 'out1_pack_weights')
    return upload(img)


def conv3d_out1(xb, in_shape, wpacked, bias, relu, terms, want_f32=True, thresholds=None):
    """Single-output-channel 3x3x3 stride-1 layer on a blocked tensor, fused with threshold + bit-pack.
    Returns (x_hat fp32 (N,1,D,H,W) or None, bits int32 (N, DHW/32) or None, counts int32 (N,) or None)."""
    L.require_cuda()
    n, cin, d, h, w = in_shape
    xh = torch.empty((n, 1, d, h, w), device=xb.device, dtype=torch.float32) if want_f32 else None
    bits = counts = None
    if thresholds is not None:
        assert thresholds.is_cuda and thresholds.dtype == torch.float32 and thresholds.numel() == n
        bits = torch.empty((n, d * h * w // 32), device=xb.device, dtype=torch.int32)
        counts = torch.empty(n, device=xb.device, dtype=torch.int32)
    L.check(L.lib().pccgeo_conv3d_out1(L.ptr(xb), L.ptr(wpacked), L.ptr(bias), L.ptr(xh), L.ptr(bits), L.ptr(thresholds),
                                       L.ptr(counts), n, cin, d, h, w, int(relu), terms, L.stream_ptr()), 'conv3d_out1')
    return xh, bits, counts


_gemm_img_size = {}


def gemm_pack_weights(w_tap_host, cin, cout, k, stride, transposed, terms):
    """numpy fp32 (k^3, Cin, Cout) -> (device uint8 image, host header bytes) for pccgeo_conv3d_gemm."""
    w = np.ascontiguousarray(w_tap_host, np.float32)
    geo = (cin, cout, k, stride, int(transposed), terms)
    size = _gemm_img_size.get(geo)
    if size is None:
        size = L.lib().pccgeo_gemm_pack_weights_host(L.ptr(w), None, cin, cout, k, stride, int(transposed), terms)
        if size <= 0:
            L.check(int(size) if size < 0 else -1, 'gemm_pack_weights')
        _gemm_img_size[geo] = size
    img = np.empty(size, np.uint8)   # the packer clears the image itself
    rc = L.lib().pccgeo_gemm_pack_weights_host(L.ptr(w), L.ptr(img), cin, cout, k, stride, int(transposed), terms)
    if rc < 0:
        L.check(int(rc), 'gemm_pack_weights')
    return upload(img), np.ascontiguousarray(img[:128].copy())


def conv3d_gemm(xb, in_shape, wimg, bias, cout, stride, transposed, relu, terms, residual_b=None, out=None):
    """General blocked-layout tensor-core conv.  wimg = (device image, host header) from gemm_pack_weights."""
    L.require_cuda()
    n, cin, d, h, w = in_shape
    shp = (n, cout, same_out(d, stride, transposed), same_out(h, stride, transposed), same_out(w, stride, transposed))
    if out is None:
        out = torch.empty(blocked_numel(*shp, terms), device=xb.device, dtype=torch.bfloat16)
    L.check(L.lib().pccgeo_conv3d_gemm(L.ptr(xb), L.ptr(wimg[0]), L.ptr(wimg[1]), L.ptr(bias), L.ptr(residual_b), L.ptr(out),
                                       n, cin, d, h, w, cout, int(relu), L.stream_ptr()), 'conv3d_gemm')
    return out, shp


# ---- entropy models ------------------------------------------------------------------------------------
_ws = {}


def reduce_ws(device):
    key = str(device)
    if key not in _ws:
        _ws[key] = torch.empty(int(L.lib().pccgeo_reduce_ws_doubles()), device=device, dtype=torch.float64)
    return _ws[key]


def eb_quantize(x, eb_params, want_symbols=True, want_xhat=True):
    L.require_cuda()
    _f32c(x)
    n, c = x.shape[:2]
    sp = x[0, 0].numel()
    sym = torch.empty(x.shape, device=x.device, dtype=torch.int32) if want_symbols else None
    xh = torch.empty_like(x) if want_xhat else None
    L.check(L.lib().pccgeo_eb_quantize(L.ptr(x), L.ptr(eb_params), L.ptr(sym), L.ptr(xh), n, c, sp, L.stream_ptr()), 'eb_quantize')
    return sym, xh


def eb_dequantize(sym, eb_params):
    L.require_cuda()
    assert sym.is_cuda and sym.dtype == torch.int32 and sym.is_contiguous()
    n, c = sym.shape[:2]
    sp = sym[0, 0].numel()
    out = torch.empty(sym.shape, device=sym.device, dtype=torch.float32)
    L.check(L.lib().pccgeo_eb_dequantize(L.ptr(sym), L.ptr(eb_params), L.ptr(out), n, c, sp, L.stream_ptr()), 'eb_dequantize')
    return out


def eb_likelihood(values, eb_params, want_likelihood=True, want_sum=True):
    L.require_cuda()
    _f32c(values)
    n, c = values.shape[:2]
    sp = values[0, 0].numel()
    lik = torch.empty_like(values) if want_likelihood else None
    s = torch.empty(1, device=values.device, dtype=torch.float64) if want_sum else None
    L.check(L.lib().pccgeo_eb_likelihood(L.ptr(values), L.ptr(eb_params), L.ptr(lik), L.ptr(s), L.ptr(reduce_ws(values.device)),
                                         n, c, sp, L.stream_ptr()), 'eb_likelihood')
    return lik, s


def gc_quantize(y, sigma, scale_table, want_symbols=True, want_yhat=True, want_indexes=True):
    L.require_cuda()
    ref = y if y is not None else sigma
    _f32c(ref)
    sym = torch.empty(ref.shape, device=ref.device, dtype=torch.int32) if (want_symbols and y is not None) else None
    yh = torch.empty(ref.shape, device=ref.device, dtype=torch.float32) if (want_yhat and y is not None) else None
    idx = torch.empty(ref.shape, device=ref.device, dtype=torch.int32) if want_indexes else None
    L.check(L.lib().pccgeo_gc_quantize(L.ptr(y), L.ptr(sigma), L.ptr(scale_table), int(scale_table.numel()),
                                       L.ptr(sym), L.ptr(yh), L.ptr(idx), ref.numel(), L.stream_ptr()), 'gc_quantize')
    return sym, yh, idx


def gc_likelihood(values, sigma, scale_min, want_likelihood=True, want_sum=True):
    L.require_cuda()
    _f32c(values)
    _f32c(sigma)
    lik = torch.empty_like(values) if want_likelihood else None
    s = torch.empty(1, device=values.device, dtype=torch.float64) if want_sum else None
    L.check(L.lib().pccgeo_gc_likelihood(L.ptr(values), L.ptr(sigma), float(scale_min), L.ptr(lik), L.ptr(s),
                                         L.ptr(reduce_ws(values.device)), values.numel(), L.stream_ptr()), 'gc_likelihood')
    return lik, s


def i32_to_f32(sym):
    L.require_cuda()
    out = torch.empty(sym.shape, device=sym.device, dtype=torch.float32)
    L.check(L.lib().pccgeo_i32_to_f32(L.ptr(sym), L.ptr(out), sym.numel(), L.stream_ptr()), 'i32_to_f32')
    return out


# ---- voxel helpers -------------------------------------------------------------------------------------
def densify(coords_i16, n, d, h, w, out=None, block0=None):
    """coords_i16: CUDA int16 (npts, 4) rows (block, z, y, x) -> fp32 (n,1,d,h,w) occupancy.  block0: the rows carry global
    block indexes (octree partition on the device); blocks [block0, block0 + n) are written."""
    L.require_cuda()
    if out is None:
        out = torch.zeros((n, 1, d, h, w), device='cuda', dtype=torch.float32)
    else:
        out.zero_()
    npts = 0 if coords_i16 is None else coords_i16.shape[0]
    if npts:
        assert coords_i16.is_cuda and coords_i16.dtype == torch.int16 and coords_i16.is_contiguous()
    if block0 is not None:
        L.check(L.lib().pccgeo_densify_from(L.ptr(coords_i16) if npts else None, npts, int(block0), L.ptr(out), n, d, h, w, L.stream_ptr()),
                'densify_from')
        return out
    L.check(L.lib().pccgeo_densify(L.ptr(coords_i16) if npts else None, npts, L.ptr(out), n, d, h, w, L.stream_ptr()), 'densify')
    return out


def threshold_pack(x_hat, thresholds):
    """x_hat fp32 (n,1,d,h,w), thresholds fp32 (n,) -> (bits uint32 as int32 tensor (n, vox/32), counts int32 (n,))."""
    L.require_cuda()
    _f32c(x_hat)
    n = x_hat.shape[0]
    vpb = x_hat[0].numel()
    bits = torch.empty((n, vpb // 32), device=x_hat.device, dtype=torch.int32)
    counts = torch.empty(n, device=x_hat.device, dtype=torch.int32)
    L.check(L.lib().pccgeo_threshold_pack(L.ptr(x_hat), L.ptr(thresholds), L.ptr(bits), L.ptr(counts), n, vpb, L.stream_ptr()),
            'threshold_pack')
    return bits, counts


def focal_loss_sum(x_true, x_pred, gamma, alpha):
    L.require_cuda()
    _f32c(x_true)
    _f32c(x_pred)
    out = torch.empty(1, device=x_true.device, dtype=torch.float64)
    L.check(L.lib().pccgeo_focal_loss(L.ptr(x_true), L.ptr(x_pred), float(gamma), float(alpha), L.ptr(out),
                                      L.ptr(reduce_ws(x_true.device)), x_true.numel(), L.stream_ptr()), 'focal_loss')
    return out


# ---- training path ---------------------------------------------------------------------------------------
def relu_bwd(dy, y):
    out = torch.empty_like(dy)
    L.check(L.lib().pccgeo_relu_bwd(L.ptr(dy), L.ptr(y), L.ptr(out), dy.numel(), L.stream_ptr()), 'relu_bwd')
    return out


def axpby(a, b=None, alpha=1.0, beta=1.0):
    out = torch.empty_like(a)
    L.check(L.lib().pccgeo_axpby(L.ptr(a), L.ptr(b), float(alpha), float(beta), L.ptr(out), a.numel(), L.stream_ptr()), 'axpby')
    return out


def focal_loss_bwd(x_true, x_pred, gamma, alpha, scale):
    out = torch.empty_like(x_pred)
    L.check(L.lib().pccgeo_focal_loss_bwd(L.ptr(_f32c(x_true)), L.ptr(_f32c(x_pred)), float(gamma), float(alpha), float(scale),
                                          L.ptr(out), x_pred.numel(), L.stream_ptr()), 'focal_loss_bwd')
    return out


def gc_likelihood_bwd(values, sigma, scale_min, c):
    dv, ds = torch.empty_like(values), torch.empty_like(values)
    L.check(L.lib().pccgeo_gc_likelihood_bwd(L.ptr(_f32c(values)), L.ptr(_f32c(sigma)), float(scale_min), float(c), L.ptr(dv),
                                             L.ptr(ds), values.numel(), L.stream_ptr()), 'gc_likelihood_bwd')
    return dv, ds


def eb_likelihood_bwd(values, eb_params, c):
    """-> (d values, d params (C,44) w.r.t. softplus'ed matrices / biases / tanh'ed factors of the packed block)"""
    _f32c(values)
    n, ch = values.shape[:2]
    sp = values[0, 0].numel()
    dv = torch.empty_like(values)
    dp = torch.empty((ch, 44), device=values.device, dtype=torch.float32)
    ws = torch.empty(int(L.lib().pccgeo_eb_bwd_ws_doubles(ch)), device=values.device, dtype=torch.float64)
    L.check(L.lib().pccgeo_eb_likelihood_bwd(L.ptr(values), L.ptr(eb_params), float(c), L.ptr(dv), L.ptr(dp), L.ptr(ws), n, ch, sp,
                                             L.stream_ptr()), 'eb_likelihood_bwd')
    return dv, dp


def wgrad_umma_eligible(n, cin, cout, k, stride, d, h, w, terms=2):
    """True when the tcgen05 weight-gradient kernel covers this layer (3x3x3, stride 1, 16/32/64 channels in == out, W in 16/32/64)"""
    if k != 3 or stride != 1 or cin != cout:
        return False
    return int(L.lib().pccgeo_wgrad_umma_ws_floats(cin, n, d, h, w, terms)) > 0


def conv3d_wgrad_umma(xb, gb, shape, transposed, terms):
    """-> dW tap-major (27, C, C) from the blocked bf16 input xb and pre-activation gradient gb (both `terms` terms) of a
    3x3x3 stride-1 'same' conv / transposed conv with C channels in and out; shape = (N, C, D, H, W)"""
    L.require_cuda()
    n, c, d, h, w = shape
    nws = int(L.lib().pccgeo_wgrad_umma_ws_floats(c, n, d, h, w, terms))
    if nws <= 0:
        raise ValueError(f'conv3d_wgrad_umma: unsupported geometry {shape}')
    dw = torch.empty((27, c, c), device=xb.device, dtype=torch.float32)
    ws = torch.empty(nws, device=xb.device, dtype=torch.float32)
    L.check(L.lib().pccgeo_conv3d_wgrad_umma(L.ptr(xb), L.ptr(gb), L.ptr(dw), L.ptr(ws), n, c, d, h, w, int(transposed), terms,
                                             L.stream_ptr()), 'conv3d_wgrad_umma')
    return dw


def conv3d_wgrad_f32(x, g, cout, k, stride, transposed):
    """-> dW tap-major (k^3, Cin, Cout) of a 'same' conv / transposed conv with input x and pre-activation gradient g"""
    _f32c(x)
    _f32c(g)
    n, cin, d, h, w = x.shape
    dw = torch.empty((k ** 3, cin, cout), device=x.device, dtype=torch.float32)
    ws = torch.empty(int(L.lib().pccgeo_wgrad_ws_floats(cin, cout, k)), device=x.device, dtype=torch.float32)
    L.check(L.lib().pccgeo_conv3d_wgrad_f32(L.ptr(x), L.ptr(g), L.ptr(dw), L.ptr(ws), n, cin, d, h, w, cout, k, stride,
                                            int(transposed), L.stream_ptr()), 'conv3d_wgrad_f32')
    return dw


def bias_grad_f32(g):
    _f32c(g)
    n, c = g.shape[:2]
    db = torch.empty(c, device=g.device, dtype=torch.float32)
    L.check(L.lib().pccgeo_bias_grad_f32(L.ptr(g), L.ptr(db), L.ptr(reduce_ws(g.device)), n, c, g[0, 0].numel(), L.stream_ptr()),
            'bias_grad_f32')
    return db


def _groups(shape):
    n, c, d, h, w = shape
    return n * (round_up(c, 16) // 8) * d * h * w


def relu_mask_blocked(gb, yb, shape, terms):
    """blocked gradient gb masked by the blocked activation yb (> 0), out of place; shape = logical (N, C, D, H, W) of both"""
    out = torch.empty_like(gb)
    L.check(L.lib().pccgeo_relu_mask_blocked(L.ptr(gb), L.ptr(yb), L.ptr(out), _groups(shape), terms, L.stream_ptr()), 'relu_mask_blocked')
    return out


def add_blocked(ab, bb, shape, terms):
    out = torch.empty_like(ab)
    L.check(L.lib().pccgeo_add_blocked(L.ptr(ab), L.ptr(bb), L.ptr(out), _groups(shape), terms, L.stream_ptr()), 'add_blocked')
    return out


def bias_grad_blocked(gb, shape, terms):
    n, c, d, h, w = shape
    db = torch.empty(c, device=gb.device, dtype=torch.float32)
    ws = torch.empty(int(L.lib().pccgeo_bias_grad_blocked_ws_doubles(c)), device=gb.device, dtype=torch.float64)
    L.check(L.lib().pccgeo_bias_grad_blocked(L.ptr(gb), L.ptr(db), L.ptr(ws), n, c, d * h * w, terms, L.stream_ptr()), 'bias_grad_blocked')
    return db


def adam_step(theta, grad, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    L.check(L.lib().pccgeo_adam_step(L.ptr(theta), L.ptr(grad), L.ptr(m), L.ptr(v), float(lr), float(beta1), float(beta2), float(eps),
                                     int(step), theta.numel(), L.stream_ptr()), 'adam_step')


# ---- host range coder ------------------------------------------------------------------------------------
def range_encode(symbols, sym_offsets, tables, indexes=None, channel_stride=0, threads=0):
    """symbols int32 (host numpy, concatenated streams); returns list of bytes, one per stream."""
    import os
    symbols = np.ascontiguousarray(symbols, np.int32)
    offs = np.ascontiguousarray(sym_offsets, np.int64)
    ns = len(offs) - 1
    cdf = np.ascontiguousarray(tables['cdf'], np.int32)
    cl = np.ascontiguousarray(tables['cdf_length'], np.int32)
    of = np.ascontiguousarray(tables['offset'], np.int32)
    mode = 0 if indexes is not None else 1
    if indexes is not None:
        indexes = np.ascontiguousarray(indexes, np.int32)
    cap = int(symbols.size) * 4 + 64 * ns + 64
    out_offs = np.zeros(ns + 1, np.int64)
    threads = threads or min(ns, os.cpu_count() or 1) or 1
    while True:
        out = np.empty(cap, np.uint8)
        rc = L.lib().pccgeo_range_encode_host(L.ptr(symbols), L.ptr(indexes), L.ptr(offs), ns, L.ptr(cdf), cdf.shape[1],
                                              L.ptr(cl), L.ptr(of), cdf.shape[0], mode, int(channel_stride),
                                              L.ptr(out), cap, L.ptr(out_offs), threads)
        if rc == L.PCCGEO_ENOSPC and int(out_offs[-1]) > cap:
            cap = int(out_offs[-1])   # escape-heavy streams (up to ~5 bytes per symbol): the call reports the size it needs
            continue
        L.check(rc, 'range_encode')
        break
    view, o = memoryview(out), out_offs.tolist()
    return [bytes(view[o[i]:o[i + 1]]) for i in range(ns)]


def range_decode(strings, sym_offsets, tables, indexes=None, channel_stride=0, threads=0):
    """strings: list of bytes; returns int32 numpy of the concatenated symbols."""
    import os
    offs = np.ascontiguousarray(sym_offsets, np.int64)
    ns = len(offs) - 1
    assert len(strings) == ns
    boffs = np.zeros(ns + 1, np.int64)
    boffs[1:] = np.cumsum([len(s) for s in strings])
    blob = np.empty(int(boffs[-1]) + 1, np.uint8)   # one gather copy (+ a pad byte: never a null pointer)
    for s, o in zip(strings, boffs.tolist()):
        if s:
            blob[o:o + len(s)] = np.frombuffer(s, np.uint8)
    blob[-1] = 0
    cdf = np.ascontiguousarray(tables['cdf'], np.int32)
    cl = np.ascontiguousarray(tables['cdf_length'], np.int32)
    of = np.ascontiguousarray(tables['offset'], np.int32)
    mode = 0 if indexes is not None else 1
    if indexes is not None:
        indexes = np.ascontiguousarray(indexes, np.int32)
    out = np.empty(int(offs[-1]), np.int32)
    threads = threads or min(ns, os.cpu_count() or 1) or 1
    L.check(L.lib().pccgeo_range_decode_host(L.ptr(blob), L.ptr(boffs), L.ptr(indexes), L.ptr(offs), ns, L.ptr(cdf),
                                             cdf.shape[1], L.ptr(cl), L.ptr(of), cdf.shape[0], mode, int(channel_stride),
                                             L.ptr(out), threads), 'range_decode')
    return out


# ---- device range coder ----------------------------------------------------------------------------------
_dev_tables = {}


def device_tables(tables):
    """{'cdf','cdf_length','offset'} host tables -> the same + the decoder's compact 16-bit rows as CUDA tensors, uploaded once per table set and device
    (keyed by the identity of the host cdf array: entropy models build a new dict when their tables change)."""
    key = (id(tables['cdf']), torch.cuda.current_device())
    hit = _dev_tables.get(key)
    if hit is not None and hit[0] is tables['cdf']:
        return hit[1]
    cdf = np.ascontiguousarray(tables['cdf'], np.int32)
    cl = np.ascontiguousarray(tables['cdf_length'], np.int32)
    of = np.ascontiguousarray(tables['offset'], np.int32)
    total = int(L.lib().pccgeo_range_compact_tables_host(L.ptr(cdf), cdf.shape[1], L.ptr(cl), cdf.shape[0], None, None))
    if total < 0:
        raise L.PccGeoError('range_compact_tables: ' + L.lib().pccgeo_last_error().decode('utf-8', 'replace'))
    cdf16, starts = np.zeros(total, np.uint16), np.zeros(cdf.shape[0], np.int32)
    L.lib().pccgeo_range_compact_tables_host(L.ptr(cdf), cdf.shape[1], L.ptr(cl), cdf.shape[0], L.ptr(cdf16), L.ptr(starts))
    dev = {'cdf': torch.from_numpy(cdf).cuda(), 'cdf_length': torch.from_numpy(cl).cuda(), 'offset': torch.from_numpy(of).cuda(),
           'cdf16': torch.from_numpy(cdf16.view(np.int16)).cuda(), 'row_start': torch.from_numpy(starts).cuda(), 'entries': total,
           'rows': cdf.shape[0], 'stride': cdf.shape[1]}
    if len(_dev_tables) > 64:
        _dev_tables.clear()
    _dev_tables[key] = (tables['cdf'], dev)
    return dev


def range_encode_device(symbols, dtab, indexes=None, channel_stride=0, packed_capacity=None):
    """symbols int32 CUDA (nstreams, ...) -> (packed uint8, lengths int32 (nstreams), offsets int64 (nstreams+1), err int32 (1))
    on the device, the bytes of pccgeo_range_encode_host.  Async on the current stream."""
    L.require_cuda()
    assert symbols.is_cuda and symbols.dtype == torch.int32 and symbols.is_contiguous()
    ns = symbols.shape[0]
    per = symbols.numel() // ns
    mode = 0 if indexes is not None else 1
    if indexes is not None:
        assert indexes.is_cuda and indexes.dtype == torch.int32 and indexes.is_contiguous() and indexes.numel() == symbols.numel()
    cap = int(packed_capacity) if packed_capacity else ns * (per * 4 + 64)
    ws = torch.empty(int(L.lib().pccgeo_rc_encode_ws_bytes(ns, per)), dtype=torch.uint8, device='cuda')
    packed = torch.empty(cap, dtype=torch.uint8, device='cuda')
    lengths = torch.empty(ns, dtype=torch.int32, device='cuda')
    offsets = torch.empty(ns + 1, dtype=torch.int64, device='cuda')
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    L.check(L.lib().pccgeo_range_encode_device(L.ptr(symbols), L.ptr(indexes), ns, per, L.ptr(dtab['cdf']), dtab['stride'],
                                               L.ptr(dtab['cdf_length']), L.ptr(dtab['offset']), dtab['rows'], mode,
                                               int(channel_stride), L.ptr(ws), L.ptr(packed), cap, L.ptr(lengths), L.ptr(offsets),
                                               L.ptr(err), L.stream_ptr()), 'range_encode_device')
    return packed, lengths, offsets, err


def range_decode_device(bytes_dev, byte_offsets, nstreams, per_stream, dtab, indexes=None, channel_stride=0, out=None, err=None):
    """bytes_dev uint8 CUDA, byte_offsets int64 CUDA (nstreams+1) -> (symbols int32 CUDA (nstreams, per_stream), err)."""
    L.require_cuda()
    assert bytes_dev.is_cuda and bytes_dev.dtype == torch.uint8 and byte_offsets.is_cuda and byte_offsets.dtype == torch.int64
    mode = 0 if indexes is not None else 1
    if indexes is not None:
        assert indexes.is_cuda and indexes.dtype == torch.int32 and indexes.is_contiguous() and indexes.numel() == nstreams * per_stream
    if out is None:
        out = torch.empty((nstreams, per_stream), dtype=torch.int32, device='cuda')
    if err is None:
        err = torch.zeros(1, dtype=torch.int32, device='cuda')
    L.check(L.lib().pccgeo_range_decode_device(L.ptr(bytes_dev), L.ptr(byte_offsets), L.ptr(indexes), nstreams, per_stream,
                                               L.ptr(dtab['cdf16']), L.ptr(dtab['row_start']), L.ptr(dtab['cdf_length']),
                                               L.ptr(dtab['offset']), dtab['rows'], dtab['entries'], mode, int(channel_stride),
                                               L.ptr(out), L.ptr(err), L.stream_ptr()), 'range_decode_device')
    return out, err


def range_encode_emulate(symbols, nstreams, tables, indexes=None, channel_stride=0):
    """Test hook: the device encoder's arithmetic run on the host (see include/pccgeo.h) -> list of bytes."""
    symbols = np.ascontiguousarray(symbols, np.int32)
    per = symbols.size // nstreams
    cdf = np.ascontiguousarray(tables['cdf'], np.int32)
    cl = np.ascontiguousarray(tables['cdf_length'], np.int32)
    of = np.ascontiguousarray(tables['offset'], np.int32)
    mode = 0 if indexes is not None else 1
    if indexes is not None:
        indexes = np.ascontiguousarray(indexes, np.int32)
    cap = symbols.size * 4 + 64 * nstreams
    packed = np.zeros(cap, np.uint8)
    lengths = np.zeros(nstreams, np.int32)
    offs = np.zeros(nstreams + 1, np.int64)
    L.check(L.lib().pccgeo_range_encode_emulate_host(L.ptr(symbols), L.ptr(indexes), nstreams, per, L.ptr(cdf), cdf.shape[1], L.ptr(cl),
                                                     L.ptr(of), cdf.shape[0], mode, int(channel_stride), L.ptr(packed), cap,
                                                     L.ptr(lengths), L.ptr(offs)), 'range_encode_emulate')
    buf = packed.tobytes()
    return [buf[offs[i]:offs[i] + max(int(lengths[i]), 0)] for i in range(nstreams)]


def bits_to_points(bits_host, dims, threads=0):
    """bits_host: numpy int32/uint32 (n, d*h*w/32) packed occupancy -> list of float32 (m_i, 3) arrays, argwhere order."""
    import os
    bits = np.ascontiguousarray(bits_host).view(np.uint32)
    n = bits.shape[0]
    d, h, w = (int(v) for v in dims)
    offs = np.zeros(n + 1, np.int64)
    threads = threads or min(n, os.cpu_count() or 1) or 1
    L.check(L.lib().pccgeo_bits_to_points_host(L.ptr(bits), n, d, h, w, L.ptr(offs), None, 0, threads), 'bits_to_points')
    pts = np.empty((int(offs[-1]), 3), np.float32)
    L.check(L.lib().pccgeo_bits_to_points_host(L.ptr(bits), n, d, h, w, L.ptr(offs), L.ptr(pts), int(offs[-1]), threads),
            'bits_to_points')
    return [pts[offs[i]:offs[i + 1]] for i in range(n)]


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(pmf, np.float64)
    cdf = np.zeros(len(pmf) + 1, np.int32)
    L.check(L.lib().pccgeo_pmf_to_quantized_cdf_host(L.ptr(pmf), len(pmf), precision, L.ptr(cdf)), 'pmf_to_quantized_cdf')
    return cdf
