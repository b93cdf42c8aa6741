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


def compress_blocks_sharded(model, blocks, **kwargs):
    """model.compress_blocks over this rank's shard + gather: returns the full data_list[0] on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    b, e = shard_range(len(blocks), rank, world)
    local = []
    if e > b:
        data_list, _, _ = model.compress_blocks(None, blocks[b:e], None, None, kwargs.pop('resolution', 0), kwargs.pop('level', 0),
                                                fixed_threshold=True, **kwargs)
        local = data_list[0]
    return gather_block_data(local)
