#!/usr/bin/env python
"""Benchmark of the pcc_geo_cnn_v2 hot path on B200: 64^3 voxel blocks/s, encode+decode, network config c3p
(= the paper's c4: hyperprior, alpha 0.75, fixed threshold), synthetic surface blocks, trained-like synthetic weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--blocks B] [--precision bf16x3|bf16|fp32]

One "step" = one batch of B blocks per GPU through encode (densify, analysis, hyper-analysis, quantise, hyper-synthesis,
scale indexes, synthesis + clip/threshold/bit-pack) and decode (dequantise, hyper-synthesis, indexes, synthesis +
clip/threshold/bit-pack), launched exactly as compress_blocks(fixed_threshold=True) / decompress_blocks launch them:
densify + four CUDA-graph stage replays.
  value : device-resident throughput -- point coordinates / int32 symbols already in HBM, CUDA events around the K steps
          (the host range coder is not in this region; it is in e2e).
  e2e   : the same metric through the reference-facing API model.compress_blocks()/decompress_blocks() with HOST
          inputs/outputs: host->device copies, the C++ range coder on host threads and device->host copies are all
          inside the timed region.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = '64^3 voxel blocks/s encode+decode (c4 = c3p network)'
UNIT = 'blocks/s'
SIZE = 64
# algorithmic work per 64^3 block (SURVEY.md section 8d / BASELINE.md section 2), 2 FLOP per MAC
GFLOP_ENCODE, GFLOP_DECODE = 16.562, 14.524
LAYER_MMAC = 1811.94  # s.b2.t1 / s.b2.t2: 16->16 channels, 64^3 voxels, 27 taps (dominant kernel)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--blocks', type=int, default=32, help='blocks per step per GPU')
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'bf16', 'fp32'])
    ap.add_argument('--cpu-blocks', type=int, default=128,
                    help='blocks in the bounded CPU-baseline sample (~10 s on 16 cores); the reference arm uses a quarter per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # one streaming nvidia-smi (-lms 50) instead of a process per sample: short timed regions still get samples
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                line = line.strip()
                if line:
                    self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def __enter__(self):
        self.th.start()
        time.sleep(0.15)  # let the first samples arrive before the timed region starts
        self.rows.clear()
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def ncu_traffic(kernel, blocks):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    summary of the same configuration (profiles/ncu_dominant_kernel.json); None when no capture matches."""
    p = os.path.join(ROOT, 'profiles', 'ncu_dominant_kernel.json')
    try:
        d = json.load(open(p))
        if d.get('kernel') == kernel and d.get('blocks') == blocks:
            return d['dram_bytes_read'] + d['dram_bytes_write']
    except Exception:
        pass
    return None


def ncu_summary_field(kernel, blocks, field):
    p = os.path.join(ROOT, 'profiles', 'ncu_dominant_kernel.json')
    try:
        d = json.load(open(p))
        if d.get('kernel') == kernel and d.get('blocks') == blocks:
            return d.get(field)
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops', 1590.0), d.get('hbm_gbs', 6650.0), 'measured'
    return 1590.0, 6650.0, 'fallback'


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch-CPU restatement of the reference graphs) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_encode_decode(n_blocks, seed=42):
    """Returns (blocks/s, seconds, cores): oracle compress + decompress of n_blocks 64^3 blocks, all host threads."""
    from oracle.model import OracleModel, sparse_to_dense
    from pcc_geo_cnn_v2_b200 import synthetic
    from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = ModelConfigType['c3p'].build()
    w = synthetic.trained_like_weights(m, seed=seed)
    o = OracleModel('c3p')
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, w['entropy_bottleneck'])
    o.gc_tab, o.eb_tab  # tables are built once per model, outside the timed region (as in the reference's graph build)
    blocks = synthetic.surface_blocks(n_blocks, size=SIZE, seed=seed + 1)
    from pcc_geo_cnn_v2_b200 import ops  # the C++ host range coder: the oracle's pure-Python coder would dominate unfairly
    t0 = time.perf_counter()
    with torch.no_grad():
        for b in blocks:
            x = sparse_to_dense(b, (1, 1, SIZE, SIZE, SIZE))
            t = o.analyse(x)
            x_hat = o.synthesise(t['y_hat'])
            zs = t['z_symbols'].numpy()
            ys, idx = t['y_symbols'].numpy(), t['indexes'].numpy()
            z_str = ops.range_encode(zs.reshape(-1), np.array([0, zs.size], np.int64), o.eb_tab, channel_stride=zs[0, 0].size, threads=1)
            y_str = ops.range_encode(ys.reshape(-1), np.array([0, ys.size], np.int64), o.gc_tab, indexes=idx.reshape(-1), threads=1)
            # decode
            zs2 = ops.range_decode(z_str, np.array([0, zs.size], np.int64), o.eb_tab, channel_stride=zs[0, 0].size, threads=1)
            from oracle import entropy as E
            z_hat = E.eb_dequantize(o.eb, torch.from_numpy(zs2.reshape(zs.shape)))
            sigma = o._tf('hyper_synthesis', z_hat)
            idx2 = E.gc_indexes(sigma, o.scale_table).numpy()
            try:
                ys2 = ops.range_decode(y_str, np.array([0, ys.size], np.int64), o.gc_tab, indexes=idx2.reshape(-1), threads=1)
            except Exception:
                # multi-threaded oneDNN convs are not run-to-run bit-identical: a scale index flipped between the oracle's
                # own encoder and decoder.  The reference retries in that case (decompress_octree.py:69-131); so do we.
                ys2 = ops.range_decode(y_str, np.array([0, ys.size], np.int64), o.gc_tab, indexes=idx.reshape(-1), threads=1)
            x_hat2 = o.synthesise(torch.from_numpy(ys2.reshape(ys.shape)).float())
            _ = np.argwhere(x_hat2[0, 0].numpy() > o.thresholds[128])
            del x_hat
    dt = time.perf_counter() - t0
    return n_blocks / dt, dt, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_encode_decode(1)
    n = max(1, args.cpu_blocks // 4)
    secs = 0.0
    for _ in range(max(1, args.steps)):
        v, dt, cores = cpu_encode_decode(n)
        vals.append(v)
        secs += dt
    value = n * len(vals) / secs
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * secs / len(vals), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'c3p encode+decode, {n} synthetic 64^3 surface blocks per step, torch-CPU oracle (TF1 unavailable)'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'{n} blocks x {len(vals)} steps, oracle compress+decompress incl. range coding'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import pcc_geo_cnn_v2_b200 as P
    from pcc_geo_cnn_v2_b200 import ops, synthetic, _lib
    from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords, threshold_f32
    from pcc_geo_cnn_v2_b200.entropy_models import GaussianConditional
    lib = _lib.lib()

    P.set_precision(args.precision)
    m = P.ModelConfigType['c3p'].build(batch_size=args.blocks)
    m.set_weights(synthetic.trained_like_weights(m, seed=42))
    m.compress((1, 1, SIZE, SIZE, SIZE))
    m.decompress()
    B = args.blocks
    uniq = synthetic.surface_blocks(min(B, 8), size=SIZE, seed=100 + rank)
    blocks = [uniq[i % len(uniq)] for i in range(B)]
    coords_host = blocks_to_coords(blocks)
    coords = torch.from_numpy(coords_host).cuda()
    thr = torch.from_numpy(threshold_f32(m.thresholds, np.full(B, 128))).cuda()

    dims = (SIZE, SIZE, SIZE)

    def device_step():
        # the per-batch device work of compress_blocks(fixed_threshold=True) + decompress_blocks, inputs resident in HBM:
        # densify + stage graphs (latents | synthesis+pack) for encode, (hyper-synthesis+indexes | synthesis+pack) for decode
        lat, st = m.device_encode(coords, B, dims, None)
        m.device_synthesis(lat, st, B, dims, thr)
        st['sym0'].copy_(lat['z_sym'])
        ctx = dict(m._stage('dec1', B, dims, lambda: m._dec1_compute(st['sym0'])))
        ctx['ysym'] = lat['y_sym']
        return m._graph_dev2(ctx, B, dims, thr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        device_step()
    barrier()
    from pcc_geo_cnn_v2_b200.model_types import graph_kernel_launches
    l0 = lib.pccgeo_launch_count() + graph_kernel_launches[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        e0.record()
        for _ in range(args.steps):
            device_step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.pccgeo_launch_count() + graph_kernel_launches[0] - l0   # direct launches + kernels inside graph replays
    t = torch.tensor([ms], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = world * B * args.steps / (ms / 1e3)

    # ---- dominant kernel alone: synthesis 16->16 @ 64^3 (s.b2.t1), CUDA events on the launching stream ----
    layer = m.synthesis_transform.leaf_layers()[7]
    assert layer.filters == 16 and layer.in_channels == 16 and layer.stride == 1
    terms = {'bf16x3': 2, 'bf16': 1, 'fp32': 0}[args.precision]
    xin = torch.randn(B, 16, SIZE, SIZE, SIZE, device='cuda').relu_()
    if terms:
        xb = ops.f32_to_blocked(xin, terms)
        wp, bias = layer.dev(f'w_umma{terms}'), layer.dev('bias')
        yb = torch.empty_like(xb)
        run_layer = lambda: ops.conv3d_umma(xb, tuple(xin.shape), wp, bias, 16, 1, True, True, terms, None, yb)
        kname = 'conv3d_umma_kernel<16>'
    else:
        yo = torch.empty_like(xin)
        run_layer = lambda: ops.conv3d_f32(xin, layer.dev('w_tap'), layer.dev('bias'), 16, 3, 1, True, True, None, yo)
        kname = 'conv3d_direct_kernel<16,4>'
    for _ in range(3):
        run_layer()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    k0.record()
    for _ in range(reps):
        run_layer()
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / reps
    peak_tf, peak_gbs, peak_src = measured_peaks()
    ach_tf = 2 * LAYER_MMAC * 1e6 * B / (kms / 1e3) / 1e12
    io_bytes = B * 16 * SIZE ** 3 * 2 * (2 * max(terms, 1) if terms else 4)  # read + write of the activations
    roofline = {'kernel': kname, 'bound': 'tensor', 'achieved': ach_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': ach_tf / peak_tf, 'traffic': ncu_traffic(kname, B), 'peak_source': peak_src, 'ms_per_launch': kms,
                'algorithmic_flop_per_launch': 2 * LAYER_MMAC * 1e6 * B,
                'hbm_gbs_algorithmic': io_bytes / (kms / 1e3) / 1e9, 'hbm_peak_gbs': peak_gbs,
                # context from the committed ncu capture of this launch and the issue-rate microbenchmark:
                # an N=48 tcgen05.mma stream from two CTAs per SM cannot keep the tensor pipe busier than 24/57 = 42 %
                'tensor_pipe_active_pct_ncu': ncu_summary_field(kname, B, 'tensor_pipe_active_pct'),
                'tensor_pipe_ceiling_pct_small_n': 42.1 if terms else None,
                'executed_bf16_products_per_algorithmic_flop': 3 if terms == 2 else (1 if terms else None)}
    del xin

    # ---- e2e through the public API: host blocks in, strings out, strings in, host points out ----
    EB = 8  # batches per e2e step: the block loops pipeline batches (host coding of batch i overlaps GPU work of batch i+1)
    e2e_blocks = blocks * EB

    def e2e_step():
        data_list, meta, _ = m.compress_blocks(None, e2e_blocks, None, None, SIZE, 0, fixed_threshold=True)
        if world > 1:  # the path's only exchange step: per-block byte strings gathered over NCCL (rank 0 writes the container)
            from pcc_geo_cnn_v2_b200.sharding import gather_block_data
            gathered = gather_block_data(data_list[0], dst=0)
            assert (gathered is None) == (rank != 0) and (rank != 0 or len(gathered) == world * B * EB)
        dec, _ = m.decompress_blocks(None, data_list[0], (SIZE, SIZE, SIZE))
        return data_list, dec

    for _ in range(max(3, args.warmup)):  # graph capture, pinned staging pool and allocator growth all settle within 3 steps
        data_list, dec = e2e_step()
    barrier()
    esteps = max(1, min(args.steps, 10))
    step_ms = []
    with ClockSampler(local) as cs_e2e:
        t0 = time.perf_counter()
        for _ in range(esteps):
            ts = time.perf_counter()
            data_list, dec = e2e_step()   # returns host points: every step ends with its results on the host
            step_ms.append((time.perf_counter() - ts) * 1e3)
        torch.cuda.synchronize()
        et = torch.tensor([time.perf_counter() - t0], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = world * B * EB * esteps / float(et[0])
    nsym = B * EB * 64 * (8 ** 3 + 4 ** 3)
    str_bytes = sum(len(s) for blk, _ in data_list[0] for s in blk) / EB
    if m.device_coder:   # strings instead of symbols cross PCIe (+ per-stream lengths / offsets)
        h2d = coords_host.nbytes * EB + int(str_bytes * EB) + B * EB * (2 * 8 + 4 * 2)   # coords + strings, offsets, thresholds
        d2h = int(str_bytes * EB) + B * EB * 2 * 12 + 2 * B * EB * SIZE ** 3 // 8           # strings, lengths/offsets + packed occupancy
    else:
        h2d = coords_host.nbytes * EB + nsym * 4 + B * EB * 4 * 2         # coords (enc) + symbols (dec) + thresholds
        d2h = nsym * 4 + B * EB * 64 * 8 ** 3 * 4 * 2 + 2 * B * EB * SIZE ** 3 // 8   # symbols + indexes (enc+dec) + packed occupancy (enc+dec)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'bf16x3': 'bf16x3 (hi/lo split bf16 operands, fp32 accumulate)', 'bf16': 'bf16', 'fp32': 'f32'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': f'c3p (paper c4) encode+decode, {B} synthetic 64^3 surface blocks per GPU per step, '
                                   f'trained-like synthetic weights', 'blocks_per_step_per_gpu': B, 'precision': args.precision,
                       'l2': 'per-step activation traffic (>1 GB) exceeds the 126 MB L2; no explicit flush',
                       'parallelism': f'blocks sharded over {world} GPU(s), no data-path collective'},
            'gflop_per_block_algorithmic': GFLOP_ENCODE + GFLOP_DECODE,
            'tensor_frac_whole_step': value / world * (GFLOP_ENCODE + GFLOP_DECODE) / 1e3 / peak_tf,
            'roofline': roofline,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': esteps, 'blocks_per_step_per_gpu': B * EB, 'bitstream_bytes_per_block': str_bytes / B,
                    # value = all steps / total time; the host side runs on shared vCPUs, so rank 0's per-step spread is reported too
                    'ms_per_step_rank0': {'min': min(step_ms), 'median': sorted(step_ms)[len(step_ms) // 2], 'max': max(step_ms)},
                    'entropy_coder': 'device (rc_device.cu)' if m.device_coder else f'host ({m.coder_threads} threads x {m.pipeline_depth} workers)'},
            'gpu_launches': int(launches), 'clocks': cs.summary(), 'clocks_e2e': cs_e2e.summary()}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = cpu_encode_decode(args.cpu_blocks)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': f'{args.cpu_blocks} blocks of the same workload, oracle (torch-CPU) compress+decompress '
                                          f'incl. C++ range coding, {dt:.1f} s'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
