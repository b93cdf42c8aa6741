"""Time the GPU threshold optimisation for one batch: python tools/threshold_opt_bench.py [B]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P
from pcc_geo_cnn_v2_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B)]
_, x_hat, _ = m.encode_blocks(blocks, keep_x_hat=True)
for _ in range(2):
    idx, names = m._optimal_thresholds(blocks, x_hat, 64, False, ('d1_mse',), (np.inf,))
torch.cuda.synchronize()
t0 = time.perf_counter()
idx, names = m._optimal_thresholds(blocks, x_hat, 64, False, ('d1_mse',), (np.inf,))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'{B} blocks x {len(m.thresholds)} thresholds: {dt * 1e3:.1f} ms ({B / dt:.0f} blocks/s); chosen', idx[:8, 0].tolist())
t0 = time.perf_counter()
dl, meta, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=False)
torch.cuda.synchronize()
print(f'compress_blocks(fixed_threshold=False): {(time.perf_counter() - t0) * 1e3:.1f} ms for {B} blocks')
