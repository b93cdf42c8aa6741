"""Timeline of the pipelined block loops: wall-clock spans of the heavy calls per worker thread."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import ops, synthetic, model_types  # noqa: E402

spans = []
T0 = [0.0]


def wrap(mod, name, label=None):
    fn = getattr(mod, name)

    def inner(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        spans.append((threading.get_ident() % 1000, label or name, t0 - T0[0], time.perf_counter() - T0[0]))
        return out
    setattr(mod, name, inner)


for n in ('range_encode', 'range_decode', 'bits_to_points', 'densify', 'threshold_pack'):
    wrap(ops, n)
wrap(model_types, 'blocks_to_coords')
_cpu = torch.Tensor.cpu


def cpu_traced(self, *a, **k):
    t0 = time.perf_counter()
    out = _cpu(self, *a, **k)
    spans.append((threading.get_ident() % 1000, f'D2H.cpu[{self.numel() * self.element_size() >> 10}K]', t0 - T0[0], time.perf_counter() - T0[0]))
    return out


torch.Tensor.cpu = cpu_traced

B, NB = 32, 4
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
uniq = synthetic.surface_blocks(8, size=64, seed=100)
blocks = [uniq[i % 8] for i in range(B * NB)]
for _ in range(2):
    dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
    m.decompress_blocks(None, dl[0], (64, 64, 64))
torch.cuda.synchronize()
spans.clear()
T0[0] = time.perf_counter()
dl, _, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
t_enc = time.perf_counter() - T0[0]
m.decompress_blocks(None, dl[0], (64, 64, 64))
t_all = time.perf_counter() - T0[0]
print(f'encode {t_enc * 1e3:.1f} ms, decode {(t_all - t_enc) * 1e3:.1f} ms for {B * NB} blocks')
for th, name, a, b in sorted(spans, key=lambda s: s[2]):
    print(f'{a * 1e3:8.2f} -> {b * 1e3:8.2f}  ({(b - a) * 1e3:6.2f} ms)  thread {th:3d}  {name}')
