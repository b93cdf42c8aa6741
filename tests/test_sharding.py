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


class _StubModel:
    """Host-only stand-in with the two hooks compress_blocks_sharded drives: the per-block part (deterministic in the block's
    content, like the kernels) and the whole-cloud selection."""

    def compress_blocks_local(self, blocks, resolution, with_normals, opt_metrics, max_deltas, fixed_threshold, debug=False):
        n, nm = len(blocks), len(opt_metrics) * len(max_deltas)
        strings = [(bytes([int(b.sum()) % 251]) * (1 + len(b) % 5), b'z' * (len(b) % 3)) for b in blocks]
        thr = np.array([[128 if fixed_threshold else (int(b.sum()) + 7 * m) % 256 for m in range(nm)] for b in blocks], np.int64).reshape(n, nm)
        pts = [[b[: 1 + (len(b) + m) % max(len(b), 1)].astype(np.float32) for b in blocks] for m in range(nm)]
        return {'strings': strings, 'thr_idx': thr, 'opt_metrics': [f'{o}_{d}' for o in opt_metrics for d in max_deltas], 'x_hat_list': pts,
                'debug': [None] * n}

    def _select_best(self, binstr, x_hat_list, level, opt_metrics, points, resolution, with_normals):
        sizes = [sum(len(p) for p in xs) for xs in x_hat_list]
        k = int(np.argmax(sizes))
        return [{'idx': k, 'metrics': {'n': sizes[k]}, 'x_hat_list': x_hat_list[k]}]


def _worker_sharded(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    blocks = [rng.integers(0, 64, size=(int(rng.integers(1, 40)), 3)).astype(np.float32) for _ in range(9)]
    m = _StubModel()
    ok = True
    for fixed, pts in ((True, None), (False, None), (False, np.vstack(blocks)), (True, np.vstack(blocks))):
        kw = dict(resolution=64, level=0, opt_metrics=('d1_mse', 'd2_mse'), max_deltas=(1.0, np.inf), fixed_threshold=fixed)
        data, meta = sharding.compress_blocks_sharded(m, blocks, binstr=None if pts is None else [1], points=pts, **kw)
        if rank != 0:
            ok &= data is None and meta is None
            continue
        loc = m.compress_blocks_local(blocks, 64, False, kw['opt_metrics'], kw['max_deltas'], fixed)
        if pts is None:
            want = [list(zip(loc['strings'], [int(v) for v in loc['thr_idx'][:, 0]]))]
        else:
            sel = m._select_best([1], loc['x_hat_list'], 0, loc['opt_metrics'], pts, 64, False)
            want = [list(zip(loc['strings'], [int(v) for v in loc['thr_idx'][:, x['idx']]])) for x in sel]
            ok &= meta[0]['idx'] == sel[0]['idx'] and meta[0]['metrics'] == sel[0]['metrics']
            ok &= all(np.array_equal(a, b) for a, b in zip(meta[0]['x_hat_list'], sel[0]['x_hat_list']))
        ok &= data == want
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_compress_blocks_sharded_world_size_2_gloo():
    """Per-block work sharded, whole-cloud selection on rank 0: equal to the single-process result for fixed and adaptive
    thresholds, with and without the whole cloud (fixed_threshold / opt_metrics / max_deltas are forwarded)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
