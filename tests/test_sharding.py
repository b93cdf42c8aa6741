"""N>1 path on CPU: block-list sharding + the string gather, world_size 2 over gloo (no GPU, no kernels)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from pcc_geo_cnn_v2_b200 import sharding


def test_shard_range_covers_everything_in_order():
    for n in (0, 1, 7, 8, 4000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_round_trip_incl_empty_strings():
    data = [((b'abc', b''), 128), ((b'', b'\x00\xff' * 300), 0), ((b'x', b'y'), 255)]
    assert sharding.unpack_block_data(sharding.pack_block_data(data)) == data
    assert sharding.unpack_block_data(sharding.pack_block_data([])) == []


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(0)           # same list on every rank
    full = [((rng.bytes(int(rng.integers(0, 700))), rng.bytes(int(rng.integers(0, 60)))), int(rng.integers(0, 256)))
            for _ in range(11)]
    b, e = sharding.shard_range(len(full), rank, world)
    got = sharding.gather_block_data(full[b:e])
    only0 = sharding.gather_block_data(full[b:e], dst=0)   # the container writer alone unpacks
    q.put((rank, got == full and (only0 == full if rank == 0 else only0 is None)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
