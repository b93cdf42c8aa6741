"""Time one tr_train.py step (c3p, batch B, 64^3) and its kernel breakdown: python tools/train_profile.py [B]"""
import collections
import os
import sys
import time

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import ops, synthetic  # noqa: E402
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.train_tensor_cores = len(sys.argv) > 2 and sys.argv[2] == 'tc'
blocks = synthetic.surface_blocks(min(B, 8), size=64, seed=5)
blocks = [blocks[i % len(blocks)] for i in range(B)]
x = ops.densify(torch.from_numpy(blocks_to_coords(blocks)).cuda(), B, 64, 64, 64)
m.train(x, 2.0, 0.75, 1e-4)
for _ in range(2):
    out = m.train_op(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 3
for _ in range(N):
    out = m.train_op(x)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / N
print(f'train step: {dt * 1e3:.1f} ms for batch {B} -> {B / dt:.1f} blocks/s; loss {out["loss"]:.4f}')
mk = m.trainer.marks
print('host marks (ms):', ', '.join(f'{a}: {(t - mk[0][1]) * 1e3:.1f}' for a, t in mk[1:]))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.train_op(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[e.name.split('(')[0][:70]]
        a[0] += e.time_range.end - e.time_range.start
        a[1] += 1
tot = sum(v for v, _ in agg.values())
print(f'kernel time {tot / 1e3:.1f} ms')
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f'{v / 1e3:9.2f} ms {100 * v / tot:5.1f}%  n={n:3d}  {k}')
print('-- wgrad launches in order')
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and ('wgrad' in e.name and 'finish' not in e.name):
        print(f'{(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms  {e.name.split("(")[0][:60]}')
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    m.train_op(x)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
