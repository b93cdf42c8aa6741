"""Data-parallel training check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/train_ddp_check.py
Every rank steps on its slice of a global batch; rank 0 also steps a single-process trainer on the whole batch with the same
noise: losses and updated parameters must agree to fp32 summation-order noise, and the replicas must stay bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P  # noqa: E402
from pcc_geo_cnn_v2_b200 import ops, synthetic  # noqa: E402
from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords  # noqa: E402
from pcc_geo_cnn_v2_b200.training import Trainer  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
size, B = 32, 4 * world
blocks = synthetic.surface_blocks(B, size=size, seed=5)
x = ops.densify(torch.from_numpy(blocks_to_coords(blocks)).cuda(), B, size, size, size)
g = torch.Generator().manual_seed(1)
ny = (torch.rand((B, 64) + (size // 8,) * 3, generator=g) - 0.5).cuda()
nz = (torch.rand((B, 64) + (size // 16,) * 3, generator=g) - 0.5).cuda()


def fresh():
    m = P.ModelConfigType['c3p'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=11, output_bias=-0.3))
    return m


sl = slice(rank * B // world, (rank + 1) * B // world)
m = fresh()
TC = bool(os.environ.get('TRAIN_TC'))   # TRAIN_TC=1: forward, data and weight gradients on the tcgen05 kernels
tr = Trainer(m, gamma=2, alpha=0.75, lmbda=3e-3, tensor_cores=TC)
vals, grads = tr.forward_backward(x[sl].contiguous(), ny[sl].contiguous(), nz[sl].contiguous())
for _ in range(2):
    tr.step(x[sl].contiguous(), ny[sl].contiguous(), nz[sl].contiguous())
flat = torch.cat([p['w'].reshape(-1) for p in tr.params.values()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    m1 = fresh()
    t1 = Trainer(m1, gamma=2, alpha=0.75, lmbda=3e-3, tensor_cores=TC)
    t1.distributed = False
    v1, g1 = t1.forward_backward(x, ny, nz)
    worst = 0.0
    for (la, pa), (lb, pb) in zip(tr.params.items(), t1.params.items()):
        ga, gb = grads[la]['w'], g1[lb]['w']
        worst = max(worst, float((ga - gb).abs().max() / (gb.abs().max() + 1e-30)))
    eb = max(float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)) for a, b in zip(grads['entropy_bottleneck'], g1['entropy_bottleneck']))
    print(f'world {world}: loss {vals["loss"]:.6f} vs single-process {v1["loss"]:.6f}; worst conv-gradient difference {worst:.2e}, '
          f'entropy-bottleneck gradients {eb:.2e}; replicas bit-identical after 2 steps: {same}', flush=True)
    tol = 2e-3 if TC else 1e-4   # bf16x3: the weight gradient of a slice and of the whole batch differ by its operand rounding
    assert abs(vals['loss'] - v1['loss']) < 1e-5 * abs(v1['loss']) and worst < tol and eb < tol and same
    print('DDP CHECK OK', flush=True)
dist.barrier()
dist.destroy_process_group()
