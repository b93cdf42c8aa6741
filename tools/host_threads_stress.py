"""Stress test of the C++ host stages fork-join helper teams: concurrent callers, random thread counts, short-lived owner threads, fork."""
import sys, time, os, threading, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table
t = gaussian_tables(make_scale_table())
rng = np.random.default_rng(0)
ns, per = 24, 500
idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * 2).astype(np.int32)
offs = np.arange(ns+1, dtype=np.int64)*per
ref = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=1)
blocks = [rng.integers(0, 64, (int(rng.integers(0, 300)), 3)).astype(np.float32) for _ in range(20)]
cref = blocks_to_coords(blocks, 1)
bits = rng.integers(0, 2**32, (6, 64*64*64//32), dtype=np.uint64).astype(np.uint32)
pref = ops.bits_to_points(bits, (64,64,64), 1)
errors = []
def worker(k, iters):
    r = np.random.default_rng(k)
    for it in range(iters):
        th = int(r.integers(1, 17))
        s = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=th)
        if s != ref: errors.append(('enc', k, it)); return
        d = ops.range_decode(s, offs, t, indexes=idx.reshape(-1), threads=int(r.integers(1, 17)))
        if not np.array_equal(d.reshape(ns, per), sym): errors.append(('dec', k, it)); return
        if not np.array_equal(blocks_to_coords(blocks, int(r.integers(1, 17))), cref): errors.append(('b2c', k, it)); return
        if it % 10 == 0:
            p = ops.bits_to_points(bits, (64,64,64), int(r.integers(1, 9)))
            if not all(np.array_equal(a, b) for a, b in zip(p, pref)): errors.append(('b2p', k, it)); return
t0 = time.time()
ths = [threading.Thread(target=worker, args=(k, 400)) for k in range(5)]
[x.start() for x in ths]; [x.join() for x in ths]
print('stress', time.time()-t0, 's errors', errors)
# short-lived threads: teams are created and destroyed with their owner threads
for rep in range(50):
    x = threading.Thread(target=worker, args=(100+rep, 3)); x.start(); x.join()
print('short-lived threads ok', errors)
# fork: the child must not wait for the parent's helpers
pid = os.fork()
if pid == 0:
    s = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=6)
    os._exit(0 if s == ref else 3)
_, status = os.waitpid(pid, 0)
print('fork child exit', os.WEXITSTATUS(status))
s = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=6)
print('parent after fork ok', s == ref)
