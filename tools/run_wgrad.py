"""Time the tcgen05 weight-gradient kernel against the fp32 kernel: python tools/run_wgrad.py C S [N] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcc_geo_cnn_v2_b200 import ops

c, s = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 32
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
x = torch.relu(torch.randn((n, c, s, s, s), device='cuda'))
g = torch.randn((n, c, s, s, s), device='cuda') * 1e-3
xb, gb = ops.f32_to_blocked(x, 2), ops.f32_to_blocked(g, 2)
flop = 2.0 * n * s ** 3 * 27 * c * c


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for terms in (2, 1):
    xb, gb = ops.f32_to_blocked(x, terms), ops.f32_to_blocked(g, terms)
    t = timed(lambda: ops.conv3d_wgrad_umma(xb, gb, tuple(x.shape), False, terms))
    print(f'wgrad umma C={c} S={s} N={n} terms={terms}: {t:.4f} ms  {flop / t / 1e9:.1f} TFLOP/s (algorithmic)')
t = timed(lambda: ops.f32_to_blocked(x, 2))
print(f'f32_to_blocked: {t:.4f} ms')
if reps <= 20:
    t = timed(lambda: ops.conv3d_wgrad_f32(x, g, c, 3, 1, False))
    print(f'wgrad fp32: {t:.4f} ms  {flop / t / 1e9:.1f} TFLOP/s')
    a = ops.conv3d_wgrad_umma(xb, gb, tuple(x.shape), False, 1)
    b = ops.conv3d_wgrad_f32(x, g, c, 3, 1, False)
    print('rel diff (terms=1):', float((a - b).abs().max() / b.abs().max()))
