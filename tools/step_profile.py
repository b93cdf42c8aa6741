"""Per-kernel device time of the bench's device step in steady state (torch.profiler / CUPTI, kernels run back to back with
warm caches -- unlike ncu, which serialises and flushes).  python tools/step_profile.py [blocks] [precision]"""
import collections
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import ops, synthetic  # noqa: E402
from pcc_geo_cnn_v2_b200.entropy_models import GaussianConditional  # noqa: E402
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords, threshold_f32  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
P.set_precision(sys.argv[2] if len(sys.argv) > 2 else 'bf16x3')
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
m.decompress()
uniq = synthetic.surface_blocks(min(B, 8), size=64, seed=100)
blocks = [uniq[i % len(uniq)] for i in range(B)]
coords = torch.from_numpy(blocks_to_coords(blocks)).cuda()
thr = torch.from_numpy(threshold_f32(m.thresholds, np.full(B, 128))).cuda()


dims = (64, 64, 64)


def device_step():
    # the per-batch device work of compress_blocks(fixed_threshold=True) + decompress_blocks, inputs resident in HBM:
    # densify + stage graphs (latents | synthesis+pack) for encode, (hyper-synthesis+indexes | synthesis+pack) for decode
    lat, st = m.device_encode(coords, B, dims, None)
    m.device_synthesis(lat, st, B, dims, thr)
    st['sym0'].copy_(lat['z_sym'])
    ctx = dict(m._stage('dec1', B, dims, lambda: m._dec1_compute(st['sym0'])))
    ctx['ysym'] = lat['y_sym']
    return m._graph_dev2(ctx, B, dims, thr)

for _ in range(5):
    device_step()
torch.cuda.synchronize()
STEPS = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        device_step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
per_step = len(evs) // STEPS
t_first, t_last = evs[0].time_range.start, evs[-1].time_range.end
print(f'{len(evs)} kernels, {per_step} per step, span {(t_last - t_first) / STEPS:.1f} us per step')
seq = evs[-per_step:]
tot = sum(e.time_range.end - e.time_range.start for e in seq)
print(f'last step: sum of kernel durations {tot:.1f} us, span {seq[-1].time_range.end - seq[0].time_range.start:.1f} us')
for e in seq:
    print(f'{e.time_range.end - e.time_range.start:9.1f} us  {e.name[:90]}')
