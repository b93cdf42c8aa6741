"""Run one layer on the gather -> tcgen05 kernel in isolation (for ncu / timing):
    python tools/run_gemm_layer.py [B] [Cin] [Cout] [size] [stride] [transposed] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402

B, C, CO, S, stride, tr, reps = [int(a) for a in (sys.argv[1:] + ['32', '64', '64', '16', '1', '1', '5'][len(sys.argv) - 1:])]
if os.environ.get('PCCGEO_GATHER_MODE'):
    from pcc_geo_cnn_v2_b200 import _lib
    _lib.check(_lib.lib().pccgeo_set_option(b'gemm_gather_mode', int(os.environ['PCCGEO_GATHER_MODE'])), 'set_option')
rng = np.random.default_rng(0)
x = torch.randn(B, C, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, C, CO)) / np.sqrt(27 * C)).astype(np.float32)
bias = torch.zeros(CO, device='cuda')
wp = ops.gemm_pack_weights(w, C, CO, 3, stride, bool(tr), 2)
xb = ops.f32_to_blocked(x, 2)
for _ in range(3):
    yb, shp = ops.conv3d_gemm(xb, tuple(x.shape), wp, bias, CO, stride, bool(tr), True, 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv3d_gemm(xb, tuple(x.shape), wp, bias, CO, stride, bool(tr), True, 2, None, yb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
so = S * stride if tr else S // stride
vox = B * (S ** 3 if (tr or stride == 1) else so ** 3)
print(f"gather_mode={os.environ.get('PCCGEO_GATHER_MODE', 'default')} " f'B={B} {C}->{CO} S={S} stride={stride} transposed={tr}: {ms:.4f} ms/launch, {2 * 27 * C * CO * vox / ms / 1e9:.1f} TFLOP/s algorithmic')
