"""Time the fused last synthesis layer (16->1 + ReLU + threshold/pack) in isolation: python tools/run_out1.py [B] [S] [terms] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402

B, S, terms, reps = [int(a) for a in (sys.argv[1:] + ['32', '64', '2', '10'][len(sys.argv) - 1:])]
rng = np.random.default_rng(0)
x = torch.randn(B, 16, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, 16, 1)) / np.sqrt(27 * 16)).astype(np.float32)
bias = torch.zeros(1, device='cuda')
wp = ops.out1_pack_weights(w, 16, True, terms)
xb = ops.f32_to_blocked(x, terms)
thr = torch.full((B,), 0.5, device='cuda')
for mode in ('bits', 'f32', 'both'):
    f = lambda: ops.conv3d_out1(xb, tuple(x.shape), wp, bias, True, terms, mode != 'bits', thr if mode != 'f32' else None)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = B * 16 * S ** 3 * 2 * terms / 1e9
    print(f'out1 {mode}: B={B} S={S} terms={terms}: {ms:.3f} ms/launch, input {gb / ms * 1e3:.0f} GB/s')
