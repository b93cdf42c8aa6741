"""Device range coder: kernel times per launch for growing stream counts, and the e2e block loops with the coder on the GPU
vs in the host workers.  python tools/rc_device_bench.py [steps] [batches]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import ops, synthetic  # noqa: E402
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 8


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t = gaussian_tables(make_scale_table())
dt = ops.device_tables(t)
rng = np.random.default_rng(0)
for spread in (() if os.environ.get('RC_SKIP_MICRO') else (0.05, 1.0)):
    for ns in (32, 256, 1024):
        per = 16384
        idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
        sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * spread).astype(np.int32)
        sd, idd = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
        packed, lengths, offsets, err = ops.range_encode_device(sd, dt, indexes=idd)
        enc = timed(lambda: ops.range_encode_device(sd, dt, indexes=idd))
        dec = timed(lambda: ops.range_decode_device(packed, offsets, ns, per, dt, indexes=idd))
        out, _ = ops.range_decode_device(packed, offsets, ns, per, dt, indexes=idd)
        ok = bool((out == sd).all().item())
        t0 = time.perf_counter()
        ref = ops.range_encode(sym.reshape(-1), np.arange(ns + 1, dtype=np.int64) * per, t, indexes=idx.reshape(-1), threads=os.cpu_count())
        t1 = time.perf_counter()
        ops.range_decode(ref, np.arange(ns + 1, dtype=np.int64) * per, t, indexes=idx.reshape(-1), threads=os.cpu_count())
        t2 = time.perf_counter()
        print(f'spread {spread} streams {ns:5d} x {per}: bytes/stream {int(offsets[-1].item()) // ns:6d}  device encode {enc:7.3f} ms  decode {dec:7.3f} ms '
              f'(roundtrip {"ok" if ok else "BAD"})   host ({os.cpu_count()} threads) encode {(t1 - t0) * 1e3:7.1f} ms  decode {(t2 - t1) * 1e3:7.1f} ms', flush=True)

B = 32
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B * NB)]
ref = None
for dev in (False, True, False, True):
    m.device_coder = dev
    best = 0
    for i in range(steps):
        t0 = time.perf_counter()
        dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
        t1 = time.perf_counter()
        dec, _ = m.decompress_blocks(None, dl[0], (64, 64, 64))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        best = max(best, B * NB / (t2 - t0))
        print(f'device_coder={dev} step {i}: encode {(t1 - t0) * 1e3:.1f} ms, decode {(t2 - t1) * 1e3:.1f} ms -> {B * NB / (t2 - t0):.0f} blk/s', flush=True)
    cur = ([s for s, _ in dl[0]], [p.tobytes() for p in dec])
    if ref is None:
        ref = cur
    print(f'device_coder={dev}: best {best:.0f} blk/s; identical to the first run: {cur == ref}', flush=True)
