"""Generates tests/golden/*.npz from the ORACLE (oracle/, torch-CPU fp32) -- the reference itself cannot run here
(no TensorFlow 1.15 / tensorflow-compression 1.3; see oracle/__init__.py: parity unpinned).

    python tests/golden/make_golden.py

Each fixture holds: the input blocks (coords), the parameter seed (weights are regenerated with
pcc_geo_cnn_v2_b200.synthetic.trained_like_weights, a deterministic numpy Generator), and the oracle's symbols,
indexes, byte strings and decoded points at the fixed threshold (idx 128)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.model import OracleModel, sparse_to_dense  # noqa: E402
from pcc_geo_cnn_v2_b200 import synthetic  # noqa: E402
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_for(config, seed, **kw):
    m = ModelConfigType[config].build()
    w = synthetic.trained_like_weights(m, seed=seed, **kw)
    o = OracleModel(config)
    eb = w['entropy_bottleneck']
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, eb)
    return o


def make(config, size, n_blocks, seed, name, **kw):
    o = oracle_for(config, seed, **kw)
    blocks = synthetic.surface_blocks(n_blocks, size=size, seed=seed + 1)
    out = {'config': config, 'size': size, 'seed': seed, 'n_blocks': n_blocks, **kw}
    t128 = o.thresholds[128]
    for j, b in enumerate(blocks):
        x = sparse_to_dense(b, (1, 1, size, size, size))
        strings, x_hat, dbg = o.compress(x)
        out[f'block{j}'] = b.astype(np.int16)
        for i, s in enumerate(strings):
            out[f'string{j}_{i}'] = np.frombuffer(s, np.uint8)
        out[f'y_sym{j}'] = dbg['y_symbols'][0].numpy().astype(np.int32)
        if 'z_symbols' in dbg:
            out[f'z_sym{j}'] = dbg['z_symbols'][0].numpy().astype(np.int32)
            out[f'idx{j}'] = dbg['indexes'][0].numpy().astype(np.uint8)
        xh = np.clip(x_hat[0, 0].numpy(), 0, 1)
        out[f'points{j}'] = np.argwhere(xh > t128).astype(np.int16)
        # voxels whose x_hat is within 1e-4 of the threshold: a different summation order may flip them
        out[f'fragile{j}'] = np.argwhere(np.abs(x_hat[0, 0].numpy() - t128) < 1e-4).astype(np.int16)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items() if not k.startswith('block')})


if __name__ == '__main__':
    torch.manual_seed(0)
    make('c3p', 32, 2, 42, 'c3p_32.npz')
    make('c3p', 64, 1, 43, 'c3p_64.npz')
    make('c1', 32, 1, 44, 'c1_32.npz')
    make('c2', 32, 1, 45, 'c2_32.npz')
    make('c3', 32, 1, 46, 'c3_32.npz')
    if 'v1_64' in sys.argv:   # full-size blocks through the k9 / k5 layers of the V1 transforms (gains / output bias: latents over +-8, half of the scale table, non-empty decodes)
        make('c1', 64, 2, 47, 'c1_64.npz', gain=4.0, synthesis_gain=3.0, output_bias=-1.3)
        make('c2', 64, 2, 48, 'c2_64.npz', gain=4.0, synthesis_gain=3.0, output_bias=-1.6)
