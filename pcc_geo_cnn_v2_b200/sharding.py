"""Multi-GPU sharding of the block list.  Octree blocks are independent (the reference
processes them one by one, src/model_types.py:192-212), so each rank takes a contiguous range of the Morton-ordered
block list -- no data-path collective -- and the only exchange is the final gather of the per-block byte strings
(+ threshold indexes) to rank 0: one all_gather of sizes, one all_gather of padded byte buffers (NCCL on CUDA tensors
over NVLink, or gloo on CPU tensors in the tests)."""
import struct

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of rank's share; concatenating over ranks restores the original order."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_block_data(block_data, out=None):
    """[(strings tuple, threshold_idx)] -> uint8 numpy buffer: u32 n_blocks, u32 n_strings, then per block
    u8 threshold idx + (u32 len | bytes) per string.  `out(nbytes)` may supply the (e.g. pinned) buffer to fill."""
    n = len(block_data)
    ns = len(block_data[0][0]) if n else 0
    total = 8 + sum(1 + sum(4 + len(t) for t in strings) for strings, _ in block_data)
    buf = out(total) if out is not None else np.empty(total, np.uint8)
    struct.pack_into('<II', buf, 0, n, ns)
    pos = 8
    for strings, thr in block_data:
        assert len(strings) == ns
        buf[pos] = int(thr)
        pos += 1
        for t in strings:
            struct.pack_into('<I', buf, pos, len(t))
            pos += 4
            if t:
                buf[pos:pos + len(t)] = np.frombuffer(t, np.uint8)
                pos += len(t)
    return buf


def unpack_block_data(buf):
    mv = memoryview(np.ascontiguousarray(buf, np.uint8))
    n, ns = struct.unpack_from('<II', mv, 0)
    pos, out = 8, []
    for _ in range(n):
        thr = mv[pos]
        pos += 1
        strings = []
        for _ in range(ns):
            ln = struct.unpack_from('<I', mv, pos)[0]
            pos += 4
            strings.append(bytes(mv[pos:pos + ln]))
            pos += ln
        out.append((tuple(strings), int(thr)))
    assert pos == len(mv), 'trailing bytes in a packed shard'
    return out


def gather_block_data(local_block_data, group=None, device=None, dst=None):
    """Every rank passes its shard's [(strings, threshold_idx)]; returns the full, ordered list on every rank, or with `dst`
    only on that rank (None elsewhere: rank 0 is the one that writes the container, the others need not unpack anything).
    Works with any backend: pass device='cuda' under NCCL."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None:
        device = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    cuda = str(device).startswith('cuda')
    if cuda:   # staging through the block loops' recycled pinned pool: one async copy each way
        from .model_types import _pinned
        stage = []

        def pinned(nbytes):
            stage.append(_pinned.get(max(nbytes, 1)))
            return stage[-1][:nbytes].numpy()
        packed = torch.from_numpy(pack_block_data(local_block_data, out=pinned))
    else:
        packed = torch.from_numpy(pack_block_data(local_block_data))
    size = torch.tensor([packed.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(v) for v in torch.cat(sizes).tolist()]
    padded = torch.zeros(max(sizes), dtype=torch.uint8, device=device)
    padded[:packed.numel()].copy_(packed, non_blocking=True)
    allb = torch.empty(world * max(sizes), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allb, padded, group=group) if cuda else dist.all_gather(list(allb.view(world, -1)), padded, group=group)
    if cuda:
        torch.cuda.current_stream().synchronize()
        _pinned.put(stage[0])
    if dst is not None and rank != dst:
        return None
    if cuda:
        hostbuf = _pinned.get(allb.numel())
        host = hostbuf[:allb.numel()]
        host.copy_(allb, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        arr = host.numpy()
    else:
        arr = allb.numpy()
    out = []
    for r, sz in enumerate(sizes):
        out += unpack_block_data(arr[r * max(sizes):r * max(sizes) + sz])
    if cuda:
        _pinned.put(hostbuf)
    return out


def merge_local_results(parts):
    """Rank-ordered per-shard results of CompressionModel.compress_blocks_local -> the same dict for the whole block list."""
    parts = [p for p in parts if p is not None]
    nm = max((p['thr_idx'].shape[1] for p in parts), default=1)
    names = next((p['opt_metrics'] for p in parts if len(p['strings'])), parts[0]['opt_metrics'] if parts else [])
    return {'strings': [s for p in parts for s in p['strings']],
            'thr_idx': np.concatenate([p['thr_idx'].reshape(-1, nm) for p in parts]) if parts else np.zeros((0, nm), np.int64),
            'opt_metrics': list(names),
            'x_hat_list': [[q for p in parts for q in p['x_hat_list'][m]] for m in range(nm)]}


def compress_blocks_sharded(model, blocks, binstr=None, points=None, resolution=0, level=0, with_normals=False,
                            opt_metrics=('d1_mse',), max_deltas=(np.inf,), fixed_threshold=False, dst=0, group=None):
    """CompressionModel.compress_blocks (reference src/model_types.py:184-218) over all ranks: every rank runs the per-block part
    on its contiguous shard of the Morton-ordered block list (networks, entropy coding, per-block threshold search for
    fixed_threshold=False), and rank `dst` alone runs what needs the whole cloud -- select_best_per_opt_metric
    (model_types.py:128-176) -- on the gathered results.  Returns (data_list, metadata) on rank dst, (None, None) elsewhere.
    Exchange: without points/binstr (block-level callers) only the byte strings + threshold indexes travel
    (gather_block_data); with them also the per-metric threshold table and decoded points (all_gather_object)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_range(len(blocks), rank, world)
    loc = model.compress_blocks_local(blocks[b:e], resolution, with_normals, opt_metrics, max_deltas, fixed_threshold)
    if points is None or binstr is None:
        data = gather_block_data(list(zip(loc['strings'], [int(v) for v in loc['thr_idx'][:, 0]])), group=group, dst=dst)
        if data is None:
            return None, None
        return [data], [{'idx': 0, 'metrics': {}, 'x_hat_list': None, 'blocks_depart': None, 'blocks_full': None}]
    parts = [None] * world
    dist.all_gather_object(parts, {k: loc[k] for k in ('strings', 'thr_idx', 'opt_metrics', 'x_hat_list')}, group=group)
    if rank != dst:
        return None, None
    full = merge_local_results(parts)
    threshold_list = [tuple(int(v) for v in full['thr_idx'][:, m]) for m in range(full['thr_idx'].shape[1])]
    metadata = model._select_best(binstr, full['x_hat_list'], level, full['opt_metrics'], points, resolution, with_normals)
    return [list(zip(full['strings'], threshold_list[x['idx']])) for x in metadata], metadata
