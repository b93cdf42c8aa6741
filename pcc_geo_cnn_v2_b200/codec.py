"""The bodies of the reference's compress_octree.py / decompress_octree.py as functions (no CLI, no TF session):

    blob, info = compress_point_cloud(model, points, resolution, octree_level, ...)      # compress_octree.py:60-113
    points_hat = decompress_point_cloud(model, blob)                                      # decompress_octree.py:30-60,127-140

Everything between the file formats runs on this package's modules: octree partitioning (GPU counting sort, csrc/octree.cu; or the
C++ host version), the batched
GPU block loops, the per-block threshold search on the GPU, the byte-compatible container, gzip."""
import gzip
import io

import numpy as np

from .model_syntax import load_compressed_file, save_compressed_file
from .octree_coding import departition_octree, partition_octree, partition_octree_gpu


def compress_point_cloud(model, points, resolution, octree_level, opt_metrics=('d1_mse',), max_deltas=(np.inf,),
                         fixed_threshold=False, with_normals=False, partition='gpu'):
    """-> (list of gzip'd container bytes, one per selected opt-metric group; list of metadata dicts with 'metrics' and
    'blocks_full', as compress_octree.py writes them to .enc.metric.json / the decoded PLY)"""
    assert resolution > 0, 'resolution must be positive'
    points = np.asarray(points, np.float64)
    block_size = resolution // (2 ** octree_level)
    if partition == 'gpu' and 1 <= octree_level <= 6 and len(points):
        # octree partition on the GPU (csrc/octree.cu); with a fixed threshold the blocks never return to the host: the block
        # loops densify batches straight from the grouped device rows
        blocks, binstr = partition_octree_gpu(points, [0, 0, 0], [resolution] * 3, octree_level, device=bool(fixed_threshold))
    else:
        blocks, binstr = partition_octree(points, [0, 0, 0], [resolution] * 3, octree_level)
    model.compress((1, 1, block_size, block_size, block_size))
    data_list, data, _ = model.compress_blocks(None, blocks, binstr, points, resolution, octree_level, with_normals=with_normals,
                                               opt_metrics=tuple(opt_metrics), max_deltas=tuple(max_deltas),
                                               fixed_threshold=fixed_threshold)
    # mtime=0: the gzip header otherwise carries the wall-clock second, and two runs on the same input differ in four bytes
    blobs = [gzip.compress(save_compressed_file(binstr, cur, resolution, octree_level), mtime=0) for cur in data_list]
    return blobs, data


def decompress_point_cloud(model, blob):
    """gzip'd container bytes -> float32 (n, 3) points of the whole cloud"""
    resolution, octree_level, binstr, blocks = load_compressed_file(io.BytesIO(gzip.decompress(blob)))
    resolution, octree_level = int(resolution), int(octree_level)
    block_size = resolution // (2 ** octree_level)
    model.decompress()
    dec_blocks, _ = model.decompress_blocks(None, blocks, (block_size,) * 3)
    dec = departition_octree(dec_blocks, [int(b) for b in binstr], [0, 0, 0], [resolution] * 3, octree_level)
    return np.vstack(dec).astype(np.float32) if len(dec) else np.zeros((0, 3), np.float32)
