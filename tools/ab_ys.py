"""A/B: y-stacked vs z-stacked kernel on one 16->16 layer: python tools/ab_ys.py [B] [S]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402

B, S = [int(a) for a in (sys.argv[1:] + ['32', '64'][len(sys.argv) - 1:])]
terms = 2
rng = np.random.default_rng(0)
x = torch.randn(B, 16, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, 16, 16)) / np.sqrt(27 * 16)).astype(np.float32)
bias = torch.zeros(16, device='cuda')
xb = ops.f32_to_blocked(x, terms)
yb = torch.empty_like(xb)
wz = ops.umma_pack_weights(w, 16, 16, 1, True, terms)
wy = ops.umma_ys_pack_weights(w, 16, 16, True, terms)
fz = lambda: ops.conv3d_umma(xb, tuple(x.shape), wz, bias, 16, 1, True, True, terms, None, yb)
fy = lambda: ops.conv3d_umma_ys(xb, tuple(x.shape), wy, bias, 16, True, terms, None, yb)
outs = {}
for name, f in (('z-stacked', fz), ('y-stacked', fy), ('z-stacked', fz), ('y-stacked', fy)):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    outs[name] = ops.blocked_to_f32(yb, tuple(x.shape), terms).clone()
    print(f'{name}: {e0.elapsed_time(e1) / 20:.4f} ms', flush=True)
d = (outs['z-stacked'] - outs['y-stacked']).abs().max().item()
print('max abs diff between the two kernels:', d, 'scale', outs['z-stacked'].abs().max().item())
