"""Point-cloud file I/O and shape helpers: the data formats on either side of the hot path (reference src/utils/pc_io.py,
which goes through pyntcloud + pandas; here a small PLY reader / writer on numpy).

Reads ASCII and binary_little_endian PLY vertex elements with scalar properties (x y z, optional nx ny nz / red green blue
...), which covers the reference's datasets (ModelNet40 blocks: binary float x/y/z; MPEG clouds: ASCII or binary with
colours / normals) and what it writes (`pa_to_df`: float32 x y z + uint8 colours)."""
import glob as _glob

import numpy as np

_PLY_TYPES = {'char': 'i1', 'int8': 'i1', 'uchar': 'u1', 'uint8': 'u1', 'short': 'i2', 'int16': 'i2', 'ushort': 'u2', 'uint16': 'u2',
              'int': 'i4', 'int32': 'i4', 'uint': 'u4', 'uint32': 'u4', 'float': 'f4', 'float32': 'f4', 'double': 'f8', 'float64': 'f8'}


def read_ply(path_or_bytes):
    """-> dict property name -> numpy column of the vertex element"""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, 'rb').read()
    head, sep, body = bytes(data).partition(b'end_header')
    if not sep or not head.startswith(b'ply'):
        raise ValueError('not a PLY file')
    body = body[body.index(b'\n') + 1:]
    fmt, n, props, in_vertex = None, 0, [], False
    for line in head.decode('ascii', 'replace').splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == 'format':
            fmt = tok[1]
        elif tok[0] == 'element':
            in_vertex = tok[1] == 'vertex'
            if in_vertex:
                n = int(tok[2])
            elif n and props:
                break   # vertex element is complete; later elements (faces) are ignored
        elif tok[0] == 'property' and in_vertex:
            if tok[1] == 'list':
                raise ValueError('list properties in the vertex element are not supported')
            props.append((tok[2], _PLY_TYPES[tok[1]]))
    if fmt == 'ascii':
        rows = np.loadtxt(body.decode('ascii').splitlines()[:n], dtype=np.float64, ndmin=2)
        return {name: rows[:, i].astype(t) for i, (name, t) in enumerate(props)}
    if fmt not in ('binary_little_endian', 'binary_big_endian'):
        raise ValueError(f'unsupported PLY format {fmt}')
    order = '<' if fmt == 'binary_little_endian' else '>'
    rec = np.frombuffer(body, np.dtype([(name, order + t) for name, t in props]), n)
    return {name: np.ascontiguousarray(rec[name]) for name, _ in props}


def load_pc(path):  # pc_io.py:35-41
    c = read_ply(path)
    return np.stack([c['x'], c['y'], c['z']], axis=1)


def load_normals(path):  # compress_octree.py:56
    c = read_ply(path)
    return np.stack([c['nx'], c['ny'], c['nz']], axis=1)


def load_points(files, batch_size=32):  # pc_io.py:71-78 (sequential: the reader is a frombuffer, not a parser)
    return [load_pc(f) for f in files]


def pa_to_df(points):  # pc_io.py:16-26 -> dict of typed columns (the reference's DataFrame)
    cols = ['x', 'y', 'z', 'red', 'green', 'blue']
    types = (['float32'] * 3) + (['uint8'] * 3)
    points = np.asarray(points)
    assert 3 <= points.shape[1] <= 6
    return {cols[i]: points[:, i].astype(types[i]) for i in range(points.shape[1])}


def write_df(path, df):  # pc_io.py:49-51 (binary_little_endian, like pyntcloud's default writer for .ply)
    names = list(df)
    n = len(df[names[0]])
    inv = {v: k for k, v in (('float', 'f4'), ('uchar', 'u1'), ('double', 'f8'), ('int', 'i4'), ('short', 'i2'), ('ushort', 'u2'), ('uint', 'u4'), ('char', 'i1'))}
    header = ['ply', 'format binary_little_endian 1.0', f'element vertex {n}']
    header += [f'property {inv[np.dtype(df[k].dtype).str[1:]]} {k}' for k in names] + ['end_header']
    rec = np.empty(n, np.dtype([(k, '<' + np.dtype(df[k].dtype).str[1:]) for k in names]))
    for k in names:
        rec[k] = df[k]
    with open(path, 'wb') as f:
        f.write(('\n'.join(header) + '\n').encode('ascii'))
        f.write(rec.tobytes())


def write_pc(path, pc):  # pc_io.py:44-46
    write_df(path, pa_to_df(pc))


def get_shape_data(resolution, data_format):  # pc_io.py:54-64
    assert data_format in ['channels_last', 'channels_first']
    p_max, p_min = np.array([resolution] * 3), np.array([0, 0, 0])
    dense = np.concatenate([p_max, [1]]) if data_format == 'channels_last' else np.concatenate([[1], p_max])
    return p_min, p_max, dense.astype('int64')


def get_files(input_glob):  # pc_io.py:67-68
    return np.array(_glob.glob(input_glob, recursive=True))
