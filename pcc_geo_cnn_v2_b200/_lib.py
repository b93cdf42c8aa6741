"""ctypes binding of libpccgeo.so (the C ABI declared in include/pccgeo.h).

The product path has no CPU fallback: if the library cannot be loaded, or a device entry point is called
without a CUDA device, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpccgeo.so')

_lib = None

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes); mirrors include/pccgeo.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    'pccgeo_last_error': (C.c_char_p, []),
    'pccgeo_version': (i32, []),
    'pccgeo_launch_count': (i64, []),
    'pccgeo_set_option': (i32, [C.c_char_p, i64]),
    'pccgeo_conv3d_f32': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_f32_to_blocked': (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_f32_phases_to_blocked': (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_blocked_to_f32': (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_umma_pack_weights_host': (i64, [vp, vp, i32, i32, i32, i32, i32]),
    'pccgeo_conv3d_umma': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_conv3d_first': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_umma_hl_pack_weights_host': (i64, [vp, vp, i32, i32, i32]),
    'pccgeo_conv3d_umma_hl': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_umma_zy_pack_weights_host': (i64, [vp, vp, i32, i32, i32, i32]),
    'pccgeo_conv3d_umma_zy': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_umma_ys_pack_weights_host': (i64, [vp, vp, i32, i32, i32, i32]),
    'pccgeo_conv3d_umma_ys': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_out1_pack_weights_host': (i64, [vp, vp, i32, i32, i32]),
    'pccgeo_conv3d_out1': (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_gemm_pack_weights_host': (i64, [vp, vp, i32, i32, i32, i32, i32, i32]),
    'pccgeo_conv3d_gemm': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_eb_quantize': (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    'pccgeo_eb_dequantize': (i32, [vp, vp, vp, i32, i32, i32, vp]),
    'pccgeo_eb_likelihood': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    'pccgeo_gc_quantize': (i32, [vp, vp, vp, i32, vp, vp, vp, i64, vp]),
    'pccgeo_gc_likelihood': (i32, [vp, vp, f32, vp, vp, vp, i64, vp]),
    'pccgeo_i32_to_f32': (i32, [vp, vp, i64, vp]),
    'pccgeo_reduce_ws_doubles': (C.c_size_t, []),
    'pccgeo_densify': (i32, [vp, i64, vp, i32, i32, i32, i32, vp]),
    'pccgeo_threshold_pack': (i32, [vp, vp, vp, vp, i32, i64, vp]),
    'pccgeo_focal_loss': (i32, [vp, vp, f32, f32, vp, vp, i64, vp]),
    'pccgeo_relu_bwd': (i32, [vp, vp, vp, i64, vp]),
    'pccgeo_axpby': (i32, [vp, vp, f32, f32, vp, i64, vp]),
    'pccgeo_focal_loss_bwd': (i32, [vp, vp, f32, f32, f32, vp, i64, vp]),
    'pccgeo_gc_likelihood_bwd': (i32, [vp, vp, f32, f32, vp, vp, i64, vp]),
    'pccgeo_eb_bwd_ws_doubles': (C.c_size_t, [i32]),
    'pccgeo_eb_likelihood_bwd': (i32, [vp, vp, f32, vp, vp, vp, i32, i32, i32, vp]),
    'pccgeo_wgrad_ws_floats': (C.c_size_t, [i32, i32, i32]),
    'pccgeo_conv3d_wgrad_f32': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_wgrad_umma_ws_floats': (i64, [i32, i32, i32, i32, i32, i32]),
    'pccgeo_conv3d_wgrad_umma': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_relu_mask_blocked': (i32, [vp, vp, vp, i64, i32, vp]),
    'pccgeo_add_blocked': (i32, [vp, vp, vp, i64, i32, vp]),
    'pccgeo_bias_grad_blocked_ws_doubles': (C.c_size_t, [i32]),
    'pccgeo_bias_grad_blocked': (i32, [vp, vp, vp, i32, i32, i64, i32, vp]),
    'pccgeo_bias_grad_f32': (i32, [vp, vp, vp, i32, i32, i64, vp]),
    'pccgeo_adam_step': (i32, [vp, vp, vp, vp, f32, f32, f32, f32, i64, i64, vp]),
    'pccgeo_range_encode_host': (i32, [vp, vp, vp, i32, vp, i32, vp, vp, i32, i32, i64, vp, i64, vp, i32]),
    'pccgeo_range_decode_host': (i32, [vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, i32, i64, vp, i32]),
    'pccgeo_pmf_to_quantized_cdf_host': (i32, [vp, i32, i32, vp]),
    'pccgeo_rc_encode_ws_bytes': (C.c_size_t, [i32, i64]),
    'pccgeo_range_encode_device': (i32, [vp, vp, i32, i64, vp, i32, vp, vp, i32, i32, i64, vp, vp, i64, vp, vp, vp, vp]),
    'pccgeo_range_decode_device': (i32, [vp, vp, vp, i32, i64, vp, vp, vp, vp, i32, i32, i32, i64, vp, vp, vp]),
    'pccgeo_rc_carry_probe_host': (C.c_uint, [vp, i32]),
    'pccgeo_range_compact_tables_host': (i64, [vp, i32, vp, i32, vp, vp]),
    'pccgeo_range_encode_emulate_host': (i32, [vp, vp, i32, i64, vp, i32, vp, vp, i32, i32, i64, vp, i64, vp, vp]),
    'pccgeo_threshold_opt_ws_bytes': (C.c_size_t, [i32, i32, i32, i32]),
    'pccgeo_threshold_hist': (i32, [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    'pccgeo_threshold_sum_ab': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    'pccgeo_blocks_to_coords_host': (i32, [vp, vp, vp, i32, i32, vp, i32]),
    'pccgeo_densify_from': (i32, [vp, i64, i32, vp, i32, i32, i32, i32, vp]),
    'pccgeo_octree_ws_bytes': (C.c_size_t, [i64, i32]),
    'pccgeo_octree_ws2_bytes': (C.c_size_t, [i64, i32]),
    'pccgeo_octree_partition_keys': (i32, [vp, i64, i32, C.c_double, i32, vp, vp]),
    'pccgeo_octree_partition_scatter': (i32, [vp, i64, i32, C.c_double, i32, i32, vp, vp, vp, vp, vp, vp]),
    'pccgeo_group_points_host': (i32, [vp, vp, i64, i32, i32, vp, vp, vp]),
    'pccgeo_bits_to_points_host': (i32, [vp, i32, i32, i32, i32, vp, vp, i64, i32]),
}


PCCGEO_OK, PCCGEO_EINVAL, PCCGEO_ECUDA, PCCGEO_ENOSPC = 0, -1, -2, -3   # include/pccgeo.h


class PccGeoError(RuntimeError):
    pass


def lib():
    """Load (building first if the .so is absent and nvcc is available) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    h = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


def check(rc, what=''):
    if rc != 0:
        msg = lib().pccgeo_last_error().decode('utf-8', 'replace')
        raise PccGeoError(f'{what} failed (code {rc}): {msg}')


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return t.data_ptr()
    return t.ctypes.data


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream.  torch.cuda.current_stream() builds a Stream object per call (~10 us: 900 calls per
    training step); the raw accessor behind it costs ~0.3 us."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


_cuda_ok = False


def require_cuda():
    global _cuda_ok
    if _cuda_ok:     # a positive answer does not change; torch.cuda.is_available() costs microseconds per call (NVML query)
        return
    import torch
    if not torch.cuda.is_available():
        raise PccGeoError('pcc_geo_cnn_v2_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    _cuda_ok = True
