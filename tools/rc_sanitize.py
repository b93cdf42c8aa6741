"""Small device-coder workload for compute-sanitizer (memcheck): both table-index modes, escapes, ragged lengths, corrupt input."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops  # noqa: E402
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table  # noqa: E402

t = gaussian_tables(make_scale_table())
dt = ops.device_tables(t)
rng = np.random.default_rng(0)
for ns, per, spread in ((5, 1000, 1.0), (3, 333, 40.0), (7, 1, 1.0), (2, 4097, 0.05)):
    idx = rng.integers(0, 64, (ns, per)).astype(np.int32)
    sym = np.round(rng.standard_normal((ns, per)) * make_scale_table()[idx] * spread).astype(np.int32)
    sd, idd = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
    packed, lengths, offsets, err = ops.range_encode_device(sd, dt, indexes=idd)
    out, err2 = ops.range_decode_device(packed, offsets, ns, per, dt, indexes=idd)
    assert bool((out == sd).all().item()) and int(err.item()) == 0 and int(err2.item()) == 0
tt = {k: v[:8] for k, v in t.items()}
dtt = ops.device_tables(tt)
sym = rng.integers(-9, 10, (4, 8 * 27)).astype(np.int32)
sd = torch.from_numpy(sym).cuda()
packed, lengths, offsets, err = ops.range_encode_device(sd, dtt, channel_stride=27)
out, _ = ops.range_decode_device(packed, offsets, 4, 8 * 27, dtt, channel_stride=27)
assert bool((out == sd).all().item())
blob = torch.from_numpy(np.frombuffer(rng.bytes(3000) + b'\0', np.uint8).copy()).cuda()
offs = torch.tensor([0, 0, 1500, 3000], dtype=torch.int64, device='cuda')
idd = torch.from_numpy(rng.integers(0, 64, (3, 2000)).astype(np.int32)).cuda()
ops.range_decode_device(blob, offs, 3, 2000, dt, indexes=idd)
torch.cuda.synchronize()
print('rc_sanitize workload ok')
