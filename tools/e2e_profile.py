"""Where does the end-to-end (host in / host out) time go?  cProfile of compress_blocks + decompress_blocks."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B)]


def step():
    data_list, meta, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
    dec, _ = m.decompress_blocks(None, data_list[0], (64, 64, 64))
    return dec


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
print('e2e blocks/s', 3 * B / (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
