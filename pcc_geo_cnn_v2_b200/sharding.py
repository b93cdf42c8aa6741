"""Multi-GPU sharding of the block list.  Octree blocks are independent (the reference
processes them one by one, src/model_types.py:192-212), so each rank takes a contiguous range of the Morton-ordered
block list -- no data-path collective -- and the only exchange is the final gather of the per-block byte strings
(+ threshold indexes) to rank 0: one all_gather of sizes, one all_gather of padded byte buffers (NCCL on CUDA tensors
over NVLink, or gloo on CPU tensors in the tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of rank's share; concatenating over ranks restores the original order."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_block_data(block_data):
    """[(strings tuple, threshold_idx)] -> uint8 numpy buffer: u32 n_blocks, u32 n_strings, then per block
    u8 threshold idx + (u32 len | bytes) per string."""
    n = len(block_data)
    ns = len(block_data[0][0]) if n else 0
    parts = [np.array([n, ns], np.uint32).tobytes()]
    for strings, thr in block_data:
        assert len(strings) == ns
        parts.append(np.uint8(thr).tobytes())
        for s in strings:
            parts.append(np.uint32(len(s)).tobytes())
            parts.append(bytes(s))
    return np.frombuffer(b''.join(parts), np.uint8).copy()


def unpack_block_data(buf):
    b = bytes(buf)
    n, ns = np.frombuffer(b[:8], np.uint32)
    pos, out = 8, []
    for _ in range(int(n)):
        thr = b[pos]
        pos += 1
        strings = []
        for _ in range(int(ns)):
            ln = int(np.frombuffer(b[pos:pos + 4], np.uint32)[0])
            pos += 4
            strings.append(b[pos:pos + ln])
            pos += ln
        out.append((tuple(strings), int(thr)))
    assert pos == len(b), 'trailing bytes in a packed shard'
    return out


def gather_block_data(local_block_data, group=None, device=None):
    """Every rank passes its shard's [(strings, threshold_idx)]; returns the full, ordered list on every rank
    (rank 0 is the one that writes the container).  Works with any backend: pass device='cuda' under NCCL."""
    world = dist.get_world_size(group)
    if device is None:
        device = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    buf = torch.from_numpy(pack_block_data(local_block_data)).to(device)
    size = torch.tensor([buf.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    padded = torch.zeros(max(sizes), dtype=torch.uint8, device=device)
    padded[:buf.numel()] = buf
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    out = []
    for b, sz in zip(bufs, sizes):
        out += unpack_block_data(b[:sz].cpu().numpy())
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
