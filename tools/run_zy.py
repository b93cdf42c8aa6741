"""Run the zy-ring 16->16 kernel in isolation (for ncu / timing):  python tools/run_zy.py [B] [size] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402

B, S, reps = [int(a) for a in (sys.argv[1:] + ['32', '64', '5'][len(sys.argv) - 1:])]
if os.environ.get('PCCGEO_ZY_GROUPS'):
    from pcc_geo_cnn_v2_b200 import _lib
    _lib.check(_lib.lib().pccgeo_set_option(b'zy_groups', int(os.environ['PCCGEO_ZY_GROUPS'])), 'set_option')
rng = np.random.default_rng(0)
x = torch.randn(B, 16, S, S, S, device='cuda').relu_()
w = (rng.normal(size=(27, 16, 16)) / np.sqrt(27 * 16)).astype(np.float32)
bias = torch.zeros(16, device='cuda')
xb = ops.f32_to_blocked(x, 2)
yb = torch.empty_like(xb)
wzy = ops.umma_zy_pack_weights(w, 16, 16, True, 2)
for _ in range(3):
    ops.conv3d_umma_zy(xb, tuple(x.shape), wzy, bias, 16, True, 2, None, yb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv3d_umma_zy(xb, tuple(x.shape), wzy, bias, 16, True, 2, None, yb)
e1.record()
torch.cuda.synchronize()
print(f"zy groups={os.environ.get('PCCGEO_ZY_GROUPS', 'default')} B={B} S={S}: {e0.elapsed_time(e1) / reps:.4f} ms/launch")
