"""Run one tensor-core conv layer in isolation (for ncu / timing):
    python tools/run_layer.py [B] [Cin] [size] [terms] [reps] [Cout] [stride]      (transposed conv; stride 1 or 2)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402

B, C, S, terms, reps, CO, stride = [int(a) for a in (sys.argv[1:] + ['8', '16', '64', '2', '3', '0', '1'][len(sys.argv) - 1:])]
CO = CO or C
rng = np.random.default_rng(0)
x = torch.randn(B, C, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, C, CO)) / np.sqrt(27 * C)).astype(np.float32)
bias = torch.zeros(CO, device='cuda')
wp = ops.umma_pack_weights(w, C, CO, stride, True, terms)
xb = ops.f32_to_blocked(x, terms)
yb = torch.empty(ops.blocked_numel(B, CO, S * stride, S * stride, S * stride, terms), device='cuda', dtype=torch.bfloat16)
for _ in range(reps):
    ops.conv3d_umma(xb, tuple(x.shape), wp, bias, CO, stride, True, True, terms, None, yb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv3d_umma(xb, tuple(x.shape), wp, bias, CO, stride, True, True, terms, None, yb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'B={B} {C}->{CO} S={S} stride={stride} terms={terms}: {ms:.3f} ms/launch, '
      f'{2 * 27 * C * CO * B * S ** 3 / ms / 1e9:.1f} TFLOP/s algorithmic')
