"""Diagnostic for the tcgen05 conv kernel: runs small cases in both descriptor interpretations and dumps outputs +
fp32 references to gpurun_out/umma_probe.npz for offline analysis.  Run under `timeout`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops, _lib  # noqa: E402

out = {}
rng = np.random.default_rng(0)


def run(tag, cin, cout, shape, n, terms, swap, transposed=False):
    _lib.lib().pccgeo_set_option(b'umma_swap_lbo_sbo', swap)
    x = torch.from_numpy(rng.normal(size=(n, cin) + shape).astype(np.float32)).cuda()
    w = (rng.normal(size=(27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    bias = torch.from_numpy(rng.normal(size=(cout,)).astype(np.float32)).cuda()
    ref = ops.conv3d_f32(x, torch.from_numpy(w).cuda(), bias, cout, 3, 1, transposed, True)
    wp = ops.umma_pack_weights(w, cin, cout, 1, transposed, terms)
    yb, shp = ops.conv3d_umma(ops.f32_to_blocked(x, terms), tuple(x.shape), wp, bias, cout, 1, transposed, True, terms)
    got = ops.blocked_to_f32(yb, shp, terms)
    torch.cuda.synchronize()
    err = float((got - ref).abs().max())
    print(f'{tag}: swap={swap} terms={terms} cin={cin} cout={cout} shape={shape} n={n} max_err={err:.4e} '
          f'ref_max={float(ref.abs().max()):.3f} got_max={float(got.abs().max()):.3f}', flush=True)
    out[f'{tag}_got'] = got.cpu().numpy()
    out[f'{tag}_ref'] = ref.cpu().numpy()
    return err


os.makedirs('gpurun_out', exist_ok=True)
errs = {}
for swap in (0, 1):
    errs[swap] = run(f'a{swap}', 16, 16, (4, 16, 8), 1, 1, swap)
    np.savez_compressed('gpurun_out/umma_probe.npz', **out)
good = min(errs, key=errs.get)
print('best interpretation: swap =', good, errs, flush=True)
run('b', 16, 16, (8, 16, 16), 2, 2, good)
run('c', 32, 32, (6, 32, 8), 1, 2, good)
run('d', 16, 16, (64, 64, 64), 2, 2, good, transposed=True)
np.savez_compressed('gpurun_out/umma_probe.npz', **{k: v for k, v in out.items() if v.size < 2_000_000})
_lib.lib().pccgeo_set_option(b'umma_swap_lbo_sbo', 0)
