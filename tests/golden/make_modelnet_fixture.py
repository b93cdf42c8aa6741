"""Extracts a small sample of the reference's own training blocks (ModelNet40_200_pc512_oct3_4k.zip: 4 000 binary PLY
files of 64^3 octree blocks, integer coordinates in [0, 63]) into tests/golden/modelnet_blocks.npz, so that the GPU tests
and the bench can run on REAL block statistics where /root/reference does not exist.

    python tests/golden/make_modelnet_fixture.py        (needs /root/reference; run in the build container only)

Selection: the smallest and the largest block of the dataset plus 22 blocks evenly spaced over the sorted file list."""
import os
import zipfile

import numpy as np

ZIP = '/root/reference/ModelNet40_200_pc512_oct3_4k.zip'
HERE = os.path.dirname(os.path.abspath(__file__))


def read_ply_xyz(data):
    """Minimal reader for the dataset's layout: binary_little_endian, float x/y/z only."""
    head, _, body = data.partition(b'end_header\n')
    lines = head.decode('ascii').split('\n')
    assert lines[0] == 'ply' and 'binary_little_endian' in lines[1]
    n = int([l for l in lines if l.startswith('element vertex')][0].split()[-1])
    props = [l.split()[-1] for l in lines if l.startswith('property')]
    assert props == ['x', 'y', 'z'], props
    return np.frombuffer(body, '<f4', n * 3).reshape(n, 3)


def main():
    z = zipfile.ZipFile(ZIP)
    names = sorted(n for n in z.namelist() if n.endswith('.ply'))
    sizes = {}
    for n in names:
        head = z.open(n).read(200)
        sizes[n] = int(head.split(b'element vertex ')[1].split(b'\n')[0])
    pick = [min(sizes, key=sizes.get), max(sizes, key=sizes.get)] + names[:: len(names) // 22][:22]
    out = {'names': np.array(pick)}
    for i, n in enumerate(pick):
        pts = read_ply_xyz(z.read(n))
        assert np.all(pts == np.round(pts)) and pts.min() >= 0 and pts.max() <= 63
        assert len(np.unique(pts, axis=0)) == len(pts)
        out[f'block{i}'] = pts.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, 'modelnet_blocks.npz'), **out)
    counts = [len(out[f'block{i}']) for i in range(len(pick))]
    print(len(pick), 'blocks, points min/median/max', min(counts), int(np.median(counts)), max(counts))


if __name__ == '__main__':
    main()
