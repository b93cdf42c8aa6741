"""Builds the oracle's C restatement (oracle/range_coder_c.c) into oracle/_build/liboracle_rc.so with gcc -- test
infrastructure, independent of the product library (nothing under pcc_geo_cnn_v2_b200/ is compiled or linked here).

    python -m oracle.build

oracle/_build/ is git-ignored but travels to the GPU box with the gpurun snapshot.  The reference itself (TensorFlow 1.15 +
tensorflow-compression 1.3, Python only, no native sources in /root/reference) cannot be compiled: there is no oracle/_ref.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'range_coder_c.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle_rc.so')


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ['gcc', '-O2', '-std=c11', '-fPIC', '-shared', '-Wall', '-o', LIB, SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('oracle build failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
