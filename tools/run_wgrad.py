"""Time the weight-gradient kernel on one layer: python tools/run_wgrad.py [B] [Cin] [Cout] [S] [stride] [transposed]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops
B, Ci, Co, S, stride, tr = [int(a) for a in (sys.argv[1:] + ['32', '16', '16', '64', '1', '1'][len(sys.argv) - 1:])]
x = torch.randn(B, Ci, S, S, S, device='cuda')
So = S * stride if tr else S // stride
g = torch.randn(B, Co, So, So, So, device='cuda')
for _ in range(2):
    ops.conv3d_wgrad_f32(x, g, Co, 3, stride, bool(tr))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.conv3d_wgrad_f32(x, g, Co, 3, stride, bool(tr))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f'wgrad {Ci}->{Co} S={S} B={B}: {ms:.2f} ms, {27 * Ci * Co * B * S ** 3 / ms / 1e9:.2f} TFMA/s')
