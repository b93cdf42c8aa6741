"""Where does the DRIVER thread of the block loops spend its time?  Wraps the driver-side calls with wall-clock timers."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import synthetic, model_types  # noqa: E402

spans = []
T0 = [0.0]
main = threading.get_ident()


def wrap(obj, name):
    fn = getattr(obj, name)

    def inner(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        spans.append(('D' if threading.get_ident() == main else 'w', name, t0 - T0[0], time.perf_counter() - T0[0]))
        return out
    setattr(obj, name, inner)


B, NB = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = P.ModelConfigType['c3p'].build(batch_size=B)
m.set_weights(synthetic.trained_like_weights(m, seed=42))
m.compress((1, 1, 64, 64, 64))
for n in ('_h2d', '_h2d_staged', '_d2h', 'device_encode', 'device_synthesis', '_graph_dev1', '_graph_dev2', '_encode_host', '_decode_host0',
          '_decode_host1', '_wait', '_copy_in', 'encode_blocks', '_strings_task', '_points_task', '_upload_strings', '_stage',
          '_decompress_blocks_device_coder', '_encode_blocks_device_coder'):
    wrap(m, n)
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402
for n in ('range_encode_device', 'range_decode_device', 'bits_to_points'):
    wrap(ops, n)
wrap(model_types, 'blocks_to_coords')
wrap(model_types._pinned, 'get')
wrap(m, '_worker_stream')
wrap(model_types._pinned, 'put')
print('device_coder', m.device_coder)
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
    print(f'{a * 1e3:8.2f} -> {b * 1e3:8.2f}  ({(b - a) * 1e3:6.2f} ms)  {th}  {name}')
