import sys, numpy as np, torch
sys.path.insert(0, '.')
from pcc_geo_cnn_v2_b200 import ops
from oracle.model import focal_loss
rng = np.random.default_rng(0)
xt = (rng.random((2,1,32,32,32)) < 0.05).astype(np.float32)
xp = rng.uniform(-0.0, 1.3, size=xt.shape).astype(np.float32); xp[rng.random(xt.shape) < 0.5] = 0.0
for gamma, alpha in ((2, 0.75), (2, 0.9)):
    p = torch.tensor(xp, dtype=torch.float64, requires_grad=True)
    focal_loss(torch.tensor(xt, dtype=torch.float64), p, gamma, alpha).backward()
    want = p.grad.numpy()
    got = ops.focal_loss_bwd(torch.from_numpy(xt).cuda(), torch.from_numpy(xp).cuda(), gamma, alpha, 1.0).cpu().numpy()
    d = np.abs(got - want)
    i = np.unravel_index(d.argmax(), d.shape)
    print('gamma', gamma, 'alpha', alpha, 'max abs err', d.max(), 'max |want|', np.abs(want).max(), 'at xp', xp[i], 'xt', xt[i], 'got', got[i], 'want', want[i])
    print('  sum got', got.astype(np.float64).sum(), 'sum want', want.sum(), 'rel', abs(got.astype(np.float64).sum()-want.sum())/abs(want.sum()))
    rel = d / (np.abs(want) + 1e-12)
    big = rel > 1e-4
    print('  n elements with rel err > 1e-4:', int((big & (np.abs(want) > 1e-6)).sum()), 'of', d.size)
    j = np.argwhere(big & (np.abs(want) > 1e-6))[:5]
    for jj in j:
        jj = tuple(jj); print('    xp', xp[jj], 'xt', xt[jj], 'got', got[jj], 'want', want[jj])
