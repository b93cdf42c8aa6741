"""A/B/C of the 16->16 layer kernels at the headline size (batch 32, 64^3, two terms): the three-product TMA kernel vs its
hi/lo-stacked form, timed alone (burst, 20 launches) and in a sustained 1.5 s loop (power-capped clocks).
    python tools/ab_hl.py [B] [size]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402

B, S = [int(a) for a in (sys.argv[1:] + ['32', '64'][len(sys.argv) - 1:])]
rng = np.random.default_rng(0)
x = torch.randn(B, 16, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, 16, 16)) / np.sqrt(27 * 16)).astype(np.float32)
bias = torch.zeros(16, device='cuda')
xb = ops.f32_to_blocked(x, 2)
yb = torch.empty_like(xb)
w3 = ops.umma_pack_weights(w, 16, 16, 1, True, 2)
whl = ops.umma_hl_pack_weights(w, 16, 16, True)
wzy = ops.umma_zy_pack_weights(w, 16, 16, True, 2)
runs = {'three-product (N=48 x 27)': lambda: ops.conv3d_umma(xb, tuple(x.shape), w3, bias, 16, 1, True, True, 2, None, yb),
        'hi/lo-stacked (N=96 x 18)': lambda: ops.conv3d_umma_hl(xb, tuple(x.shape), whl, bias, 16, True, True, None, yb),
        'zy-ring (N=144 x 9 per row)': lambda: ops.conv3d_umma_zy(xb, tuple(x.shape), wzy, bias, 16, True, 2, None, yb)}
flop = 2 * 27 * 16 * 16 * B * S ** 3


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, fn in runs.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    time.sleep(0.5)
    burst = timed(fn, 20)
    sus = timed(fn, 4000)
    print(f'{name}: burst {burst:.4f} ms ({flop / burst / 1e9:.0f} TFLOP/s algorithmic), sustained {sus:.4f} ms ({flop / sus / 1e9:.0f} TFLOP/s)')
