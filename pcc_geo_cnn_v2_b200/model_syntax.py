"""The compressed-file container, byte-compatible with the reference's src/model_syntax.py:20-58:

    uint16 resolution | uint8 octree_level | uint16 n_blocks | uint8 n_strings | uint16 n_binstr | uint8 binstr[n_binstr]
    per block: uint8 threshold_idx, then per string: uint16 n_bytes, bytes

(little endian; the reference gzips the result, compress_octree.py:112).  Pinned against bytes produced by the reference
module itself (tests/golden/ref_host_fixtures.npz).  The uint16 fields are the reference's limits (65 535 blocks / bytes per
string); values beyond them raise like the reference's asserts do."""
import struct


def _check(x, lo, hi, what):
    if not (lo <= int(x) <= hi):
        raise AssertionError(f'{"Overflow" if int(x) > hi else "Underflow"} {what}={x} (allowed {lo}..{hi})')
    return int(x)


def save_compressed_file(binstr, data_b_list, resolution, octree_level):
    """model_syntax.py:20-35 -> bytes"""
    binstr = [_check(b, 0, 255, 'binstr') for b in binstr]
    out = [struct.pack('<HBHBH', _check(resolution, 0, 65535, 'resolution'), _check(octree_level, 0, 255, 'octree_level'),
                       _check(len(data_b_list), 0, 65535, 'n_blocks'), _check(len(data_b_list[0][0]), 0, 255, 'n_strings'),
                       _check(len(binstr), 0, 65535, 'n_binstr')), bytes(binstr)]
    for strings, best_threshold_idx in data_b_list:
        out.append(struct.pack('<B', _check(best_threshold_idx, 0, 255, 'threshold_idx')))
        for s in strings:
            out.append(struct.pack('<H', _check(len(s), 0, 65535, 'string length')))
            out.append(bytes(s))
    return b''.join(out)


def load_compressed_file(f):
    """model_syntax.py:38-58: file object -> (resolution, level, binstr uint8 array, [(strings, threshold_idx)])"""
    import numpy as np
    resolution, level, n_blocks, n_strings, n_binstr = struct.unpack('<HBHBH', f.read(8))
    binstr = np.frombuffer(f.read(n_binstr), dtype=np.uint8)
    blocks = []
    for _ in range(n_blocks):
        (thr,) = struct.unpack('<B', f.read(1))
        strings = []
        for _ in range(n_strings):
            (nb,) = struct.unpack('<H', f.read(2))
            strings.append(f.read(nb))
        blocks.append((strings, np.uint8(thr)))
    file_end = f.read()
    assert file_end == b'', f'File not read completely file_end {file_end}'
    return np.uint16(resolution), np.uint8(level), binstr, blocks
