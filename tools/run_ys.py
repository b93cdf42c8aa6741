import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops
B, S, terms = 32, 64, 2
rng = np.random.default_rng(0)
x = torch.randn(B, 16, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, 16, 16)) / np.sqrt(27 * 16)).astype(np.float32)
bias = torch.zeros(16, device='cuda')
xb = ops.f32_to_blocked(x, terms); yb = torch.empty_like(xb)
wy = ops.umma_ys_pack_weights(w, 16, 16, True, terms)
for _ in range(5):
    ops.conv3d_umma_ys(xb, tuple(x.shape), wy, bias, 16, True, terms, None, yb)
torch.cuda.synchronize()
