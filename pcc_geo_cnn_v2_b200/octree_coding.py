"""Octree partition / departition of a point cloud into blocks, output-compatible with the reference's
src/utils/octree_coding.py:68-169: same blocks (float64, local coordinates, input order inside a
block), same block order, same occupancy bytes.  Pinned against the reference module's own outputs
(tests/golden/ref_host_fixtures.npz).

The reference orders blocks by a bit-string key of width geo_level - level per coordinate (octree_coding.py:93-97): a true
(z,y,x)-interleaved Morton key whenever geo_level >= 2*level -- every configuration of its experiments -- and a truncated
one otherwise, for which its own departition_octree returns wrong coordinates.  This implementation always uses the
Morton order (identical in the first regime, self-consistent in the second).  The per-point Python loop of the reference
(7.6 s on longdress, octree_coding.py:66) is a C++ counting sort here."""
import numpy as np

from . import _lib as L


def _morton(ids, level):
    """(z,y,x)-interleaved key, most significant bit first: the reference's sort key as an integer"""
    key = np.zeros(len(ids), np.uint64)
    x, y, z = (ids[:, i].astype(np.uint64) for i in range(3))
    for b in range(level - 1, -1, -1):
        for v in (z, y, x):
            key = (key << np.uint64(1)) | ((v >> np.uint64(b)) & np.uint64(1))
    return key


def _binstr(ids, level):
    """Occupancy bytes in the reference's order (octree_coding.py:48-61): a node's byte, then its occupied children's
    subtrees in child order; child index = x_bit | y_bit << 1 | z_bit << 2.  ids: unique block ids in Morton order."""
    out = []

    def rec(sub, lvl):
        if lvl == 0 or len(sub) == 0:
            return
        shift = lvl - 1
        child = ((sub[:, 0] >> shift) & 1) | (((sub[:, 1] >> shift) & 1) << 1) | (((sub[:, 2] >> shift) & 1) << 2)
        byte = 0
        for c in np.unique(child):
            byte |= 1 << int(c)
        out.append(byte)
        if lvl > 1:
            for c in range(8):
                m = child == c
                if m.any():
                    rec(sub[m], lvl - 1)

    rec(ids.astype(np.int64), level)
    return out


def partition_octree(points, bbox_min, bbox_max, level):
    """octree_coding.py:68-113 -> (list of float64 (n_i, cols) blocks in local coordinates, list of occupancy bytes)"""
    points = np.asarray(points)
    if len(points) == 0 or level == 0:
        return [points], None
    np.testing.assert_array_equal(np.asarray(bbox_min), [0, 0, 0])
    geo_level = int(np.ceil(np.log2(np.max(np.asarray(bbox_max)))))
    assert geo_level >= level
    block_size = 2 ** (geo_level - level)
    rows = np.ascontiguousarray(points, np.float64)
    ids = (rows[:, :3] // block_size).astype(np.uint32)
    # the Morton key of every point's block: ascending key order IS the reference's block order, so ranking the occupied
    # keys replaces its row-wise np.unique + string sort (and the per-point Python loop is the C++ counting sort below)
    key = _morton(ids, level)
    if 3 * level <= 24:
        present = np.zeros(1 << (3 * level), np.bool_)
        present[key] = True
        rank_of_key = np.cumsum(present, dtype=np.int64) - 1
        block_idx = np.ascontiguousarray(rank_of_key[key], np.int32)
        ukeys = np.flatnonzero(present).astype(np.uint64)
    else:
        ukeys, inverse = np.unique(key, return_inverse=True)
        block_idx = np.ascontiguousarray(np.asarray(inverse).reshape(-1), np.int32)
    uniq = np.zeros((len(ukeys), 3), np.uint32)      # de-interleave the occupied keys back to (x, y, z) block ids
    for b in range(level):
        for axis, shift in ((0, 0), (1, 1), (2, 2)):
            uniq[:, axis] |= (((ukeys >> np.uint64(3 * b + shift)) & np.uint64(1)) << np.uint64(b)).astype(np.uint32)
    origins = np.ascontiguousarray(uniq.astype(np.float64) * block_size)
    out = np.empty_like(rows)
    offsets = np.zeros(len(uniq) + 1, np.int64)
    L.check(L.lib().pccgeo_group_points_host(L.ptr(rows), L.ptr(block_idx), len(rows), rows.shape[1], len(uniq), L.ptr(origins),
                                             L.ptr(out), L.ptr(offsets)), 'group_points')
    blocks = [out[offsets[i]:offsets[i + 1]] for i in range(len(uniq))]
    return blocks, _binstr(uniq, level)


class DeviceBlocks:
    """Result of partition_octree_gpu(..., device=True): the blocks stay on the GPU as int16 (block, i0, i1, i2) rows grouped by
    block (global block index, Morton order) -- the block loops densify batches straight from it (no host partition, no
    per-batch coordinate packing, no H2D of coordinates).  len() / counts / to_host() give the reference's view."""

    def __init__(self, coords, offsets, rows=None, cols=3):
        self.coords, self.offsets, self._rows, self.cols = coords, np.asarray(offsets, np.int64), rows, cols

    def __len__(self):
        return len(self.offsets) - 1

    @property
    def counts(self):
        return np.diff(self.offsets)

    def to_host(self):
        """list of float64 (n_i, cols) arrays in local coordinates, as partition_octree returns them"""
        if self._rows is not None:
            out = self._rows.cpu().numpy()
        else:
            out = self.coords[:, 1:].cpu().numpy().astype(np.float64)
        return [out[self.offsets[i]:self.offsets[i + 1]] for i in range(len(self))]


def partition_octree_gpu(points, bbox_min, bbox_max, level, device=False):
    """partition_octree (octree_coding.py:68-113) as a stable counting sort on the GPU (csrc/octree.cu): same blocks, block
    order, in-block point order and occupancy bytes.  device=False -> (list of float64 blocks on the host, binstr) exactly like
    partition_octree; device=True -> (DeviceBlocks, binstr) for compress_blocks without the coordinates ever returning to the
    host.  1 <= level <= 6."""
    import torch
    points = np.asarray(points)
    if len(points) == 0 or level == 0:
        return [points], None
    np.testing.assert_array_equal(np.asarray(bbox_min), [0, 0, 0])
    geo_level = int(np.ceil(np.log2(np.max(np.asarray(bbox_max)))))
    assert geo_level >= level and 1 <= level <= 6
    block_size = float(2 ** (geo_level - level))
    L.require_cuda()
    rows = torch.from_numpy(np.ascontiguousarray(points, np.float64)).cuda()
    n, cols = rows.shape
    lib = L.lib()
    ws_bytes = int(lib.pccgeo_octree_ws_bytes(n, level))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    L.check(lib.pccgeo_octree_partition_keys(L.ptr(rows), n, cols, block_size, level, L.ptr(ws), L.stream_ptr()), 'octree_partition_keys')
    meta = ws[ws_bytes - 64:ws_bytes - 56].view(torch.int32).cpu().numpy()     # the one read-back: number of occupied blocks
    if int(meta[1]):
        raise ValueError('partition_octree: coordinates outside [0, block_size * 2^level)')
    nb = int(meta[0])
    ws2 = torch.empty(int(lib.pccgeo_octree_ws2_bytes(n, nb)), dtype=torch.uint8, device='cuda')
    offsets = torch.empty(nb + 1, dtype=torch.int64, device='cuda')
    out_rows = None if device and cols == 3 else torch.empty_like(rows)
    coords = torch.empty((n, 4), dtype=torch.int16, device='cuda') if device else None
    L.check(lib.pccgeo_octree_partition_scatter(L.ptr(rows), n, cols, block_size, level, nb, L.ptr(ws), L.ptr(ws2), L.ptr(out_rows),
                                                L.ptr(coords), L.ptr(offsets), L.stream_ptr()), 'octree_partition_scatter')
    n_keys = 1 << (3 * level)
    kb = (n * 4 + 255) // 256 * 256
    present = ws[kb:kb + 4 * n_keys].view(torch.int32).cpu().numpy()
    ukeys = np.flatnonzero(present).astype(np.uint64)
    uniq = np.zeros((len(ukeys), 3), np.uint32)
    for b in range(level):
        for axis in range(3):
            uniq[:, axis] |= (((ukeys >> np.uint64(3 * b + axis)) & np.uint64(1)) << np.uint64(b)).astype(np.uint32)
    binstr = _binstr(uniq, level)
    offs = offsets.cpu().numpy()
    if device:
        return DeviceBlocks(coords, offs, out_rows, cols), binstr
    host = out_rows.cpu().numpy()
    return [host[offs[i]:offs[i + 1]] for i in range(nb)], binstr


def departition_octree(blocks, binstr_list, bbox_min, bbox_max, level):
    """octree_coding.py:116-169: adds every block's origin back (blocks in the order partition_octree emits them)."""
    bbox_min, bbox_max = np.asarray(bbox_min), np.asarray(bbox_max)
    origins = []
    pos = [0]

    def rec(lo, hi, lvl):
        byte = int(binstr_list[pos[0]])
        pos[0] += 1
        mid = (hi - lo) // 2 + lo
        for c in range(8):
            if not (byte >> c) & 1:
                continue
            clo, chi = lo.copy(), mid.copy()
            for axis in range(3):
                if (c >> axis) & 1:
                    clo[axis], chi[axis] = mid[axis], hi[axis]
            if lvl == level:
                origins.append(clo)
            else:
                rec(clo, chi, lvl + 1)

    rec(bbox_min.astype(np.int64), bbox_max.astype(np.int64), 1)
    assert len(origins) == len(blocks), f'{len(origins)} leaves in the octree, {len(blocks)} blocks'
    out = []
    for b, o in zip(blocks, origins):
        b = np.array(b, copy=True)
        b[:, :3] = b[:, :3] + o
        out.append(b)
    return out
