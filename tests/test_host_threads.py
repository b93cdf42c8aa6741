"""The C++ host stages' fork-join helper teams (range_coder.cpp: HelperTeam): concurrent callers with changing thread counts
and short-lived owner threads give the single-threaded results.  (tools/host_threads_stress.py is the long version, incl. fork.)"""
import threading

import numpy as np

from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords


def test_concurrent_callers_and_changing_thread_counts():
    t = gaussian_tables(make_scale_table())
    rng = np.random.default_rng(0)
    ns, per = 24, 300
    idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
    sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * 2).astype(np.int32)
    offs = np.arange(ns + 1, dtype=np.int64) * per
    ref = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=1)
    blocks = [rng.integers(0, 64, (int(rng.integers(0, 300)), 3)).astype(np.float32) for _ in range(20)]
    cref = blocks_to_coords(blocks, 1)
    errors = []

    def worker(k, iters):
        r = np.random.default_rng(k)
        for it in range(iters):
            s = ops.range_encode(sym.reshape(-1), offs, t, indexes=idx.reshape(-1), threads=int(r.integers(1, 17)))
            d = ops.range_decode(s, offs, t, indexes=idx.reshape(-1), threads=int(r.integers(1, 17)))
            c = blocks_to_coords(blocks, int(r.integers(1, 17)))
            if s != ref or not np.array_equal(d.reshape(ns, per), sym) or not np.array_equal(c, cref):
                errors.append((k, it))
                return

    threads = [threading.Thread(target=worker, args=(k, 80)) for k in range(4)]
    for x in threads:
        x.start()
    for x in threads:
        x.join(timeout=120)
    assert not any(x.is_alive() for x in threads), 'a caller is stuck in a fork-join'
    for rep in range(20):   # teams are created and destroyed with their owner threads
        x = threading.Thread(target=worker, args=(100 + rep, 2))
        x.start()
        x.join(timeout=60)
        assert not x.is_alive()
    assert errors == []
