"""e2e throughput vs pipeline depth / coder threads; encode and decode timed separately."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402

B, NB = 32, 8
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.device_coder = False
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B * NB)]
for depth, thr in ((3, 8), (3, 12), (3, 16), (2, 16), (2, 12), (4, 12), (3, 14), (3, 10), (3, 12), (3, 8)):
    m.pipeline_depth, m.coder_threads = depth, thr
    m._executor = None
    for _ in range(2):
        dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
        m.decompress_blocks(None, dl[0], (64, 64, 64))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(4):
        m.decompress_blocks(None, dl[0], (64, 64, 64))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = 4 * B * NB
    print(f'depth={depth} coder_threads={thr}: encode {n / (t1 - t0):.0f} blk/s, decode {n / (t2 - t1):.0f} blk/s, '
          f'enc+dec {n / (t2 - t0):.0f} blk/s', flush=True)
