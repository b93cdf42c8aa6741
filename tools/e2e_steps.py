"""Per-step wall time of the e2e block loops (drift / warm-up check): python tools/e2e_steps.py [steps] [batches]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = 32
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights((synthetic.trained_like_weights if os.environ.get("E2E_STRESS") else synthetic.codec_like_weights)(m, seed=42))
m.compress((1, 1, 64, 64, 64))
if os.environ.get('E2E_GROUP'):
    m.coder_group_blocks = int(os.environ['E2E_GROUP'])
if os.environ.get('E2E_OVERLAP'):
    m.coder_overlap = os.environ['E2E_OVERLAP'] == '1'
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B * NB)]
import gc  # noqa: E402
for i in range(steps):
    if os.environ.get('E2E_GC_OFF_HALFWAY') and i == steps // 2:
        gc.collect()
        gc.freeze()
        gc.disable()
        print('-- gc frozen + disabled --', flush=True)
    t0 = time.perf_counter()
    dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
    t1 = time.perf_counter()
    m.decompress_blocks(None, dl[0], (64, 64, 64))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'step {i}: encode {(t1 - t0) * 1e3:.1f} ms, decode {(t2 - t1) * 1e3:.1f} ms -> {B * NB / (t2 - t0):.0f} blk/s', flush=True)
