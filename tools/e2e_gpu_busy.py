"""How busy is the GPU during the e2e block loops?  CUPTI (torch.profiler) kernel / memcpy intervals of one compress_blocks and one
decompress_blocks call: union of kernel intervals vs the call's span, per-kernel totals, and the largest idle gaps.
    python tools/e2e_gpu_busy.py [batches]"""
import collections
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 24
B = 32
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.codec_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B * NB)]
for _ in range(3):
    dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
    m.decompress_blocks(None, dl[0], (64, 64, 64))
torch.cuda.synchronize()


def analyse(name, fn):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    kern = sorted((a, b, n) for a, b, n in ev if not n.startswith('Memcpy') and not n.startswith('Memset'))
    cop = [(a, b, n) for a, b, n in ev if n.startswith('Memcpy')]
    span = (max(b for _, b, _ in ev) - min(a for a, _, _ in ev)) / 1e3
    busy, cur_a, cur_b, gaps = 0.0, None, None, []
    for a, b, n in kern:
        if cur_a is None:
            cur_a, cur_b = a, b
        elif a <= cur_b:
            cur_b = max(cur_b, b)
        else:
            busy += cur_b - cur_a
            gaps.append((a - cur_b, n))
            cur_a, cur_b = a, b
    busy += cur_b - cur_a
    print(f'{name}: wall {wall * 1e3:.1f} ms, GPU span {span:.1f} ms, kernels busy (union) {busy / 1e3:.1f} ms = {busy / 1e3 / span * 100:.0f} % of span; '
          f'sum of kernel durations {sum(b - a for a, b, _ in kern) / 1e3:.1f} ms; memcpy {sum(b - a for a, b, _ in cop) / 1e3:.1f} ms in {len(cop)} copies')
    gaps.sort(reverse=True)
    tot_gap = sum(g for g, _ in gaps)
    print(f'  idle gaps: {len(gaps)} totalling {tot_gap / 1e3:.1f} ms; largest: ' + ', '.join(f'{g:.0f} us before {n.split("(")[0][-40:]}' for g, n in gaps[:6]))
    hist = collections.Counter()
    for g, n in gaps:
        hist[n.split('(')[0][-50:]] += g
    print('  gap time by the kernel that follows: ' + '; '.join(f'{n}: {v / 1e3:.2f} ms' for n, v in hist.most_common(6)))
    agg = collections.defaultdict(float)
    for a, b, n in kern:
        agg[n.split('(')[0][-50:]] += b - a
    print('  kernel time: ' + '; '.join(f'{n}: {v / 1e3:.1f}' for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    return out


dl = analyse('compress_blocks', lambda: m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True))[0]
analyse('decompress_blocks', lambda: m.decompress_blocks(None, dl[0], (64, 64, 64)))
