#!/usr/bin/env python
"""Benchmark of the pcc_geo_cnn_v2 hot path on B200: 64^3 voxel blocks/s, encode+decode, network config c3p
(= the paper's c4: hyperprior, alpha 0.75, fixed threshold), synthetic surface blocks, synthetic weights at a codec-like
(rate-realistic) operating point.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--blocks B] [--batches S]
                    [--precision bf16x3|bf16|fp32] [--weights codec|stress]

One "step" = one job of S batches (default 24) of B = 32 blocks per GPU through encode (densify, analysis, hyper-analysis,
quantise, hyper-synthesis, scale indexes, synthesis + clip/threshold/bit-pack, range ENcoding of both latents) and decode
(range DEcoding of the hyper-latent, hyper-synthesis, indexes, range decoding of the latent, synthesis +
clip/threshold/bit-pack), i.e. everything compress_blocks(fixed_threshold=True) + decompress_blocks do for 768 blocks; the
kernels run at batch 32 (BASELINE.json configs[1]).  20 steps = 2 s of GPU work: a sustained, not a burst, number.
  value : device-resident -- point coordinates in HBM at the start, packed occupancy bits in HBM at the end, entropy coding by
          the device range coder (strings stay in HBM); CUDA events around the K steps.  `transforms_only` is the same loop
          without the entropy coder (round 1's definition of `value`).
  e2e   : the same metric through the reference-facing API model.compress_blocks()/decompress_blocks() with HOST
          inputs/outputs: host->device copies, the range coder (host threads or device, whichever the model selects for this
          host) and device->host copies are all inside the timed region.
`--weights codec` (default): synthetic.codec_like_weights -- ~1 KB of bitstream per block, most latents 0, no escape codes,
like a trained codec; `--weights stress`: round 1's synthetic.trained_like_weights (30 KB per block, escape-heavy), also
reported as e2e_escape_heavy.  Extra driver-visible numbers on the same line: `train_step` (tr_train.py path, batch 32) and
`workloads` (24 ModelNet40 blocks; 128^3 blocks).
`--impl reference`: the CPU arm -- the oracle (torch-CPU restatement of the reference's TF1 graphs + the oracle's own C range
coder), all host threads, batch 1 like the reference; nothing of libpccgeo is loaded.
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
GFLOP_TRAIN = 3 * GFLOP_ENCODE   # forward + data gradients + weight gradients


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--blocks', type=int, default=32, help='blocks per batch (kernel launch) per GPU')
    ap.add_argument('--batches', type=int, default=24, help='batches per step per GPU')
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'bf16', 'fp32'])
    ap.add_argument('--weights', default='codec', choices=['codec', 'stress'])
    ap.add_argument('--cpu-blocks', type=int, default=128, help='blocks in the bounded CPU-baseline sample (~10 s on 16 cores)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--lanes', type=int, default=2, help='independent chains of batches in flight on separate streams in the device-resident step')
    ap.add_argument('--no-extras', action='store_true', help='skip train_step / workloads / the second coder (profiling runs)')
    return ap.parse_args()


def workload_name(args):
    kind = 'codec-like synthetic weights (~1 KB/block)' if args.weights == 'codec' else 'escape-heavy synthetic weights (~30 KB/block)'
    return (f'c3p (paper c4) encode+decode incl. entropy coding, synthetic 64^3 surface blocks, {args.blocks} blocks per batch x '
            f'{args.batches} batches per step per GPU, {kind}')


def make_weights(model, args, seed=42):
    from pcc_geo_cnn_v2_b200 import synthetic
    return synthetic.codec_like_weights(model, seed=seed) if args.weights == 'codec' else synthetic.trained_like_weights(model, seed=seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
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
        def num(r, i):
            try:
                return float(r[i])
            except Exception:
                return None
        sm = [v for v in (num(r, 0) for r in self.rows) if v is not None]
        mx = [v for v in (num(r, 1) for r in self.rows) if v is not None]
        pw = [v for v in (num(r, 6) for r in self.rows if len(r) > 6) if v is not None]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_mhz_min': min(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'reasons': reasons, 'samples': len(self.rows)}


def ncu_summary(kernel, blocks):
    """The committed `ncu --set full` summary of the dominant kernel in this configuration (None when no capture matches)."""
    p = os.path.join(ROOT, 'profiles', 'ncu_dominant_kernel.json')
    try:
        d = json.load(open(p))
        if d.get('kernel') == kernel and d.get('blocks') == blocks:
            return d
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), d.get('hbm_gbs', 6650.0), 'measured'
    return 1590.0, 1400.0, 6650.0, 'fallback'


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch-CPU restatement of the reference graphs + its own C range coder) on the host cores.
# Nothing of the product library is loaded here (asserted at the end of run_reference).
# ---------------------------------------------------------------------------------------------------------
def cpu_encode_decode(args, n_blocks, seed=42):
    """Returns (blocks/s, seconds, cores): oracle compress + decompress of n_blocks 64^3 blocks, batch 1 like the reference's
    block loops (model_types.py:192-198), all host threads inside every op."""
    from oracle import entropy as E
    from oracle import range_coder as RC
    from oracle.model import OracleModel, sparse_to_dense
    from pcc_geo_cnn_v2_b200 import synthetic
    from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = ModelConfigType['c3p'].build()          # host object only: supplies the layer shapes for the weight generator
    w = make_weights(m, args, seed)
    o = OracleModel('c3p')
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, w['entropy_bottleneck'])
    gt, et = o.gc_tab, o.eb_tab  # tables are built once per model, outside the timed region (as in the reference's graph build)
    blocks = synthetic.surface_blocks(n_blocks, size=SIZE, seed=seed + 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        for b in blocks:
            x = sparse_to_dense(b, (1, 1, SIZE, SIZE, SIZE))
            t = o.analyse(x)
            x_hat = o.synthesise(t['y_hat'])
            zs = t['z_symbols'][0].numpy()
            ys, idx = t['y_symbols'][0].numpy(), t['indexes'][0].numpy()
            zidx = o._channel_indexes(zs.shape)
            z_str = RC.encode_c(zs, zidx, et['cdf'], et['cdf_length'], et['offset'])
            y_str = RC.encode_c(ys, idx, gt['cdf'], gt['cdf_length'], gt['offset'])
            # decode
            zs2 = RC.decode_c(z_str, zidx, et['cdf'], et['cdf_length'], et['offset'])
            z_hat = E.eb_dequantize(o.eb, torch.from_numpy(zs2[None]))
            sigma = o._tf('hyper_synthesis', z_hat)
            idx2 = E.gc_indexes(sigma, o.scale_table)[0].numpy()
            try:
                ys2 = RC.decode_c(y_str, idx2, gt['cdf'], gt['cdf_length'], gt['offset'])
            except ValueError:
                # multi-threaded oneDNN convs are not run-to-run bit-identical: a scale index flipped between the oracle's
                # own encoder and decoder.  The reference retries in that case (decompress_octree.py:69-131); so do we.
                ys2 = RC.decode_c(y_str, idx, gt['cdf'], gt['cdf_length'], gt['offset'])
            x_hat2 = o.synthesise(torch.from_numpy(ys2[None]).float())
            _ = np.argwhere(x_hat2[0, 0].numpy() > o.thresholds[128])
            del x_hat
    dt = time.perf_counter() - t0
    return n_blocks / dt, dt, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    cpu_encode_decode(args, 1)   # warm-up: oneDNN primitive caches, table construction
    # each step codes a bounded sample of the step's blocks, sized so that K steps end within a few minutes (~11 blocks/s on 16 cores)
    n = int(max(4, min(args.blocks, round(150.0 * 11.0 / max(1, args.steps)))))
    secs, cores = 0.0, 1
    for _ in range(max(1, args.steps)):
        v, dt, cores = cpu_encode_decode(args, n)
        secs += dt
    steps = max(1, args.steps)
    value = n * steps / secs
    import pcc_geo_cnn_v2_b200._lib as L
    assert L._lib is None, 'the CPU arm must not load the product library'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * secs / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'weights': args.weights},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'{n} blocks of the workload per step x {steps} steps, batch 1; torch-CPU oracle of the TF1 graphs + '
                                       f'the oracle\'s own C range coder (TF1 / tfc cannot be installed); product library not loaded'},
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
    from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords, threshold_f32, graph_kernel_launches
    lib = _lib.lib()

    P.set_precision(args.precision)
    B, S = args.blocks, args.batches
    NB = B * S
    dims = (SIZE, SIZE, SIZE)

    def build_model(weights_kind):
        m = P.ModelConfigType['c3p'].build(batch_size=B)
        a = argparse.Namespace(**{**vars(args), 'weights': weights_kind})
        m.set_weights(make_weights(m, a))
        m.compress((1, 1, SIZE, SIZE, SIZE))
        m.decompress()
        return m

    m = build_model(args.weights)
    uniq = synthetic.surface_blocks(min(B, 8), size=SIZE, seed=100 + rank)
    blocks = [uniq[i % len(uniq)] for i in range(B)]
    coords_host = blocks_to_coords(blocks)
    coords = torch.from_numpy(coords_host).cuda()
    thr = torch.from_numpy(threshold_f32(m.thresholds, np.full(B, 128))).cuda()
    lats = m._coder_latents(dims)                      # [y (indexed by the hyperprior's scales), z (per-channel tables)]
    dtabs = [ops.device_tables(l['tables']) for l in lats]
    LANES = max(1, min(args.lanes, S))
    errs = [torch.zeros(1, dtype=torch.int32, device='cuda') for _ in range(LANES)]
    lane_streams = [torch.cuda.Stream() for _ in range(LANES)]

    def device_step(code=True, check=None):
        """One step = S batches of B blocks.  With --lanes L > 1 the batches are split into L independent chains (own static buffers
        and stage graphs in the model, own stream): the small-volume layers and the latency-bound entropy-coder kernels of one chain
        then share the GPU with the large convolutions of the other instead of leaving most SMs idle."""
        if LANES == 1:
            return lane_step(S, errs[0], code, check)
        cur = torch.cuda.current_stream()
        out, per = [], [S // LANES + (1 if i < S % LANES else 0) for i in range(LANES)]
        for lane in range(LANES):
            lane_streams[lane].wait_stream(cur)
            with torch.cuda.stream(lane_streams[lane]):
                m.lane = lane
                out.append(lane_step(per[lane], errs[lane], code, check))
        m.lane = 0
        for st_ in lane_streams:
            cur.wait_stream(st_)
        return None if out[0] is None else sum(o * n for o, n in zip(out, per)) / S

    def lane_step(S, err, code=True, check=None):
        """All device work of compress_blocks(fixed_threshold=True) + decompress_blocks for S batches of B blocks, inputs and
        outputs resident in HBM.  code=False leaves the entropy coder out (the decoder is fed the encoder's symbols)."""
        NB = B * S
        big = {k: torch.empty((NB,) + tuple(l['shape']), dtype=torch.int32, device='cuda')
               for l in lats for k in (l['sym'], l['idx']) if k is not None}
        enc_bits = []
        for s in range(S):
            lat, st = m.device_encode(coords, B, dims, None)
            for k, t in big.items():
                t[s * B:(s + 1) * B].copy_(lat[k].view(t[s * B:(s + 1) * B].shape))
            bits = m.device_synthesis(lat, st, B, dims, thr)
            if check is not None:
                enc_bits.append(bits.clone())
        if code:
            coded = []
            for l, dt in zip(lats, dtabs):
                per = int(np.prod(l['shape']))
                packed, lengths, offsets, e = ops.range_encode_device(big[l['sym']], dt, indexes=big[l['idx']] if l['idx'] else None,
                                                                      channel_stride=per // l['shape'][0])
                coded.append((packed, offsets, lengths, e))
            perz = int(np.prod(lats[1]['shape']))
            zsym, _ = ops.range_decode_device(coded[1][0], coded[1][1], NB, perz, dtabs[1], channel_stride=perz // lats[1]['shape'][0], err=err)
        else:
            zsym = big['z_sym']
        idx_big = torch.empty_like(big['indexes'])
        for s in range(S):
            st = m._static(B, dims)
            st['sym0'].copy_(zsym[s * B:(s + 1) * B].view(st['sym0'].shape))
            ctx = m._stage('dec1', B, dims, lambda: m._dec1_compute(st['sym0']))
            idx_big[s * B:(s + 1) * B].copy_(ctx['indexes'].view(idx_big[s * B:(s + 1) * B].shape))
        if code:
            pery = int(np.prod(lats[0]['shape']))
            ysym, _ = ops.range_decode_device(coded[0][0], coded[0][1], NB, pery, dtabs[0], indexes=idx_big, err=err)
        else:
            ysym = big['y_sym']
        for s in range(S):
            st = m._static(B, dims)
            st['sym1'].copy_(ysym[s * B:(s + 1) * B].view(st['sym1'].shape))
            st['thr'].copy_(thr)
            bits = m._stage('dec2', B, dims, lambda: m._dec2_compute(None, st))
            if check is not None:
                check.append(bool(torch.equal(bits, enc_bits[s])))
        if check is not None and code:
            check.append(bool(torch.equal(ysym.view(-1), big['y_sym'].view(-1))) and bool(torch.equal(idx_big, big['indexes'])))
            check.append(int(err.item()) == 0 and all(int(c[3].item()) == 0 for c in coded) and all(int(c[2].min().item()) >= 0 for c in coded))
            return float(sum(int(c[1][-1].item()) for c in coded)) / NB
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    W = max(3, args.warmup)
    for _ in range(W):
        device_step()
    checks = []
    dev_bytes_per_block = device_step(check=checks)     # decoder output == encoder output, symbols round-trip through the strings
    assert all(checks), 'device-resident step: decode does not reproduce the encoder'
    barrier()
    l0 = lib.pccgeo_launch_count() + graph_kernel_launches[0]
    with ClockSampler(local) as cs:
        ms = timed(device_step, args.steps)
    launches = lib.pccgeo_launch_count() + graph_kernel_launches[0] - l0   # direct launches + kernels inside graph replays
    value = world * NB * args.steps / (ms / 1e3)
    for _ in range(2):
        device_step(code=False)
    barrier()
    tsteps = max(2, args.steps // 4)
    ms_t = timed(lambda: device_step(code=False), tsteps)
    value_t = world * NB * tsteps / (ms_t / 1e3)

    # ---- dominant kernel alone: synthesis 16->16 @ 64^3 (s.b2.t1), CUDA events on the launching stream ----
    layer = m.synthesis_transform.leaf_layers()[7]
    assert layer.filters == 16 and layer.in_channels == 16 and layer.stride == 1
    terms = {'bf16x3': 2, 'bf16': 1, 'fp32': 0}[args.precision]
    xin = torch.randn(B, 16, SIZE, SIZE, SIZE, device='cuda').relu_()
    if terms:
        # same dispatch as the transforms (whichever tcgen05 kernel serves 16->16 layers), blocked bf16 in and out
        xb = ops.f32_to_blocked(xin, terms)
        run_layer, kname = P.model_transforms.layer_runner(layer, xb, tuple(xin.shape), terms)
    else:
        yo = torch.empty_like(xin)
        run_layer = lambda: ops.conv3d_f32(xin, layer.dev('w_tap'), layer.dev('bias'), 16, 3, 1, True, True, None, yo)
        kname = 'conv3d_direct_kernel<16,4>'
    for _ in range(3):
        run_layer()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    k0.record()
    for _ in range(reps):
        run_layer()
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / reps
    peak_tf, peak_tf_sus, peak_gbs, peak_src = measured_peaks()
    ach_tf = 2 * LAYER_MMAC * 1e6 * B / (kms / 1e3) / 1e12
    io_bytes = B * 16 * SIZE ** 3 * 2 * (2 * max(terms, 1) if terms else 4)  # read + write of the activations
    ncu = ncu_summary(kname, B) or {}
    roofline = {'kernel': kname, 'bound': 'tensor', 'achieved': ach_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': ach_tf / peak_tf, 'traffic': (ncu['dram_bytes_read'] + ncu['dram_bytes_write']) if ncu else None,
                'traffic_source': 'profiles/ncu_dominant_kernel.json (ncu --set full of this launch)' if ncu else None,
                'peak_source': peak_src + ' (cuBLAS bf16 burst: the kernel is timed alone)', 'ms_per_launch': kms,
                'algorithmic_flop_per_launch': 2 * LAYER_MMAC * 1e6 * B, 'algorithmic_bytes_per_launch': io_bytes,
                'hbm_gbs_algorithmic': io_bytes / (kms / 1e3) / 1e9, 'hbm_peak_gbs': peak_gbs,
                'tensor_pipe_active_pct_ncu': ncu.get('tensor_pipe_active_pct'),
                'executed_bf16_products_per_algorithmic_flop': 3 if terms == 2 else (1 if terms else None)}
    del xin

    # ---- e2e through the public API: host blocks in, strings out, strings in, host points out ----
    e2e_blocks = blocks * S

    def e2e_run(model, blks, steps_cap=10):
        def step():
            data_list, meta, _ = model.compress_blocks(None, blks, None, None, SIZE, 0, fixed_threshold=True)
            if world > 1:  # the path's only exchange step: per-block byte strings gathered over NCCL (rank 0 writes the container)
                from pcc_geo_cnn_v2_b200.sharding import gather_block_data
                gathered = gather_block_data(data_list[0], dst=0)
                assert (gathered is None) == (rank != 0) and (rank != 0 or len(gathered) == world * len(blks))
            dec, _ = model.decompress_blocks(None, data_list[0], dims)
            return data_list, meta, dec
        for _ in range(W):  # graph capture, pinned staging pool and allocator growth all settle within 3 steps
            data_list, meta, dec = step()
        same = all(np.array_equal(a, b) for a, b in zip(meta[0]['x_hat_list'], dec))   # every rank: decoder == encoder
        barrier()
        esteps = max(1, min(args.steps, steps_cap))
        step_ms = []
        with ClockSampler(local) as cs_e:
            t0 = time.perf_counter()
            for _ in range(esteps):
                ts = time.perf_counter()
                data_list, meta, dec = step()   # returns host points: every step ends with its results on the host
                step_ms.append((time.perf_counter() - ts) * 1e3)
            torch.cuda.synchronize()
            et = torch.tensor([time.perf_counter() - t0], device='cuda', dtype=torch.float64)
        ok = torch.tensor([1 if same else 0], device='cuda')
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        n = len(blks)
        str_bytes = sum(len(s) for blk, _ in data_list[0] for s in blk)
        npts = sum(len(b) for b in blks)
        nsym = n * 64 * (8 ** 3 + 4 ** 3)
        if model.device_coder:   # strings instead of symbols cross PCIe (+ per-stream lengths / offsets)
            h2d = npts * 8 + str_bytes + n * (2 * 8 + 4 * 2)
            d2h = str_bytes + n * 2 * 12 + 2 * n * SIZE ** 3 // 8
        else:
            sb, ib = model.symbol_bytes, model.index_bytes
            h2d = npts * 8 + nsym * sb + n * 4 * 2
            d2h = nsym * sb + n * 64 * 8 ** 3 * ib * 2 + 2 * n * SIZE ** 3 // 8
        return {'value': world * n * esteps / float(et[0]), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'steps': esteps, 'blocks_per_step_per_gpu': n, 'bitstream_bytes_per_block': str_bytes / n,
                'bits_per_input_point': 8.0 * str_bytes / max(1, npts),
                'ms_per_step_rank0': {'min': min(step_ms), 'median': sorted(step_ms)[len(step_ms) // 2], 'max': max(step_ms)},
                'decoder_equals_encoder': bool(int(ok[0])),
                'entropy_coder': 'device (rc_device.cu)' if model.device_coder else f'host ({model.coder_threads} threads x {model.pipeline_depth} workers)',
                'clocks': cs_e.summary()}

    e2e = e2e_run(m, e2e_blocks)
    extras = {}
    if not args.no_extras:
        # the other entropy coder on the same workload, and the other weight set with the default coder
        m.device_coder = not m.device_coder
        extras['e2e_other_coder'] = e2e_run(m, e2e_blocks, steps_cap=5)
        m.device_coder = not m.device_coder
        other = 'stress' if args.weights == 'codec' else 'codec'
        m2 = build_model(other)
        extras['e2e_realistic' if other == 'codec' else 'e2e_escape_heavy'] = e2e_run(m2, e2e_blocks, steps_cap=5)
        del m2
        if args.weights == 'codec':
            extras['e2e_realistic'] = dict(e2e, note='same run as `e2e`: the default workload is the rate-realistic one')
        torch.cuda.empty_cache()
        if world == 1:
            extras['workloads'] = extra_workloads(P, m, args, e2e_run)
            extras['train_step'] = train_step(P, args, peak_tf_sus)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': W,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'bf16x3': 'bf16x3 (hi/lo split bf16 operands, fp32 accumulate)', 'bf16': 'bf16', 'fp32': 'f32'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': workload_name(args), 'weights': args.weights, 'blocks_per_batch': B, 'batches_per_step': S,
                       'blocks_per_step_per_gpu': NB, 'precision': args.precision,
                       'lanes': f'{LANES} independent chain(s) of batches in flight on separate CUDA streams',
                       'value_includes': 'densify, all four transforms, quantisation, scale indexes, device range encoder + decoder '
                                         '(strings stay in HBM), clip/threshold/bit-pack',
                       'l2': 'per-batch activation traffic (>1 GB) exceeds the 126 MB L2; no explicit flush',
                       'parallelism': f'blocks sharded over {world} GPU(s), no data-path collective'},
            'transforms_only': {'value': value_t, 'unit': UNIT, 'ms_per_batch': ms_t / tsteps / S,
                                'note': 'same loop without the entropy coder (the decoder is fed the encoder\'s symbols): round 1\'s `value`'},
            'device_bitstream_bytes_per_block': dev_bytes_per_block,
            'roundtrip_checked': True,
            'gflop_per_block_algorithmic': GFLOP_ENCODE + GFLOP_DECODE,
            'tensor_frac_whole_step': value_t / world * (GFLOP_ENCODE + GFLOP_DECODE) / 1e3 / peak_tf_sus,
            'tensor_frac_whole_step_peak': f'{peak_tf_sus} TFLOP/s ({peak_src}, cuBLAS bf16 sustained: the step runs for seconds)',
            'roofline': roofline, 'e2e': e2e, **extras,
            'gpu_launches': int(launches), 'clocks': cs.summary()}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = cpu_encode_decode(args, args.cpu_blocks)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': f'{args.cpu_blocks} blocks of the same workload, batch 1, oracle (torch-CPU restatement of the '
                                          f'TF1 graphs) + the oracle\'s own C range coder, {dt:.1f} s'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_workloads(P, m, args, e2e_run):
    """BASELINE.json configs 3/4 stand-ins that fit one GPU: the reference's own ModelNet40 blocks (tests/golden/modelnet_blocks.npz,
    24 blocks of ModelNet40_200_pc512_oct3_4k) and 128^3 blocks (1024-resolution clouds at octree level 3)."""
    from pcc_geo_cnn_v2_b200 import synthetic
    out = {}
    try:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'modelnet_blocks.npz'))
        real = [g[f'block{i}'].astype(np.float32) for i in range(len(g['names']))]
        reps = 32   # 768 blocks per call, like the headline step (the reference's set has 4 k blocks; a call's fill / drain is amortised over them)
        r = e2e_run(m, real * reps, steps_cap=5)
        out['modelnet24'] = {'value': r['value'], 'unit': UNIT, 'blocks_per_step': len(real) * reps, 'unique_blocks': len(real),
                             'points_per_block_mean': float(np.mean([len(b) for b in real])), 'bits_per_input_point': r['bits_per_input_point'],
                             'bitstream_bytes_per_block': r['bitstream_bytes_per_block'], 'decoder_equals_encoder': r['decoder_equals_encoder'],
                             'entropy_coder': r['entropy_coder'], 'what': 'e2e compress_blocks + decompress_blocks, host in / host out'}
    except Exception as e:  # the fixture is part of the repo; report rather than hide
        out['modelnet24'] = {'error': repr(e)}
    size = 128
    m128 = P.ModelConfigType['c3p'].build(batch_size=4)
    m128.set_weights(m.get_weights())
    m128.compress((1, 1, size, size, size))
    m128.decompress()
    blks = synthetic.surface_blocks(4, size=size, seed=7) * 24   # 96 blocks = 768 64^3-equivalents per call
    W = 2

    def step():
        dl, meta, _ = m128.compress_blocks(None, blks, None, None, 1024, 3, fixed_threshold=True)
        dec, _ = m128.decompress_blocks(None, dl[0], (size,) * 3)
        return meta, dec
    for _ in range(W):
        meta, dec = step()
    same = all(np.array_equal(a, b) for a, b in zip(meta[0]['x_hat_list'], dec))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out['cube128'] = {'value': len(blks) * n / dt, 'unit': '128^3 blocks/s', 'equivalent_64_cube_blocks_per_s': 8 * len(blks) * n / dt,
                      'blocks_per_step': len(blks), 'batch': 4, 'decoder_equals_encoder': bool(same),
                      'what': 'e2e compress_blocks + decompress_blocks on 128^3 blocks (8x the voxels of a 64^3 block)'}
    del m128
    torch.cuda.empty_cache()
    return out


def train_step(P, args, peak_tf):
    """BASELINE.json configs[4]: tr_train.py path, c3p, focal loss + entropy-model backward, both Adam steps, batch 32, lambda 1e-4."""
    from pcc_geo_cnn_v2_b200 import synthetic
    from pcc_geo_cnn_v2_b200.model_types import blocks_to_coords
    from pcc_geo_cnn_v2_b200 import ops
    B = 32
    out = {}
    for mode in ('tensor_cores', 'fp32'):
        m = P.ModelConfigType['c3p'].build(batch_size=B)
        m.set_weights(synthetic.trained_like_weights(m, seed=42))
        m.train_tensor_cores = mode == 'tensor_cores'
        blocks = synthetic.surface_blocks(8, size=SIZE, seed=3) * (B // 8)
        x = ops.densify(torch.from_numpy(blocks_to_coords(blocks)).cuda(), B, SIZE, SIZE, SIZE)
        m.train(x, 2, 0.75, 1e-4)
        for _ in range(2):
            v = m.train_op(x)
        torch.cuda.synchronize()
        n = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            v = m.train_op(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[mode] = {'ms_per_step': ms, 'tflops_algorithmic': GFLOP_TRAIN * B / ms, 'frac_of_sustained_bf16_peak': GFLOP_TRAIN * B / ms / peak_tf,
                     'loss': float(v['loss'])}
        del m, x
        torch.cuda.empty_cache()
    best = min(out, key=lambda k: out[k]['ms_per_step'])
    wg = None
    try:   # committed ncu summary of the tcgen05 weight-gradient kernel (one launch, 16 -> 16 @ 64^3, batch 32)
        with open(os.path.join(ROOT, 'profiles', 'r02_ncu_wgrad16.json')) as f:
            d = json.load(f)
        wg = {'kernel': 'wgrad_umma_kernel<16,3,3,2>', 'tensor_pipe_active_pct_ncu': d.get('tensor_pipe_active_pct'),
              'dram_bytes_per_launch_ncu': (d.get('dram_bytes_read') or 0) + (d.get('dram_bytes_write') or 0),
              'source': 'profiles/r02_ncu_wgrad16.json (static, not from this run)'}
    except Exception:
        pass
    return {'ms_per_step': out[best]['ms_per_step'], 'mode': best, 'batch': B, 'config': 'c3p, 64^3, gamma 2, alpha 0.75, lambda 1e-4',
            'gflop_per_step_algorithmic': GFLOP_TRAIN * B, 'modes': out, 'weight_gradient_kernel': wg,
            'what': 'forward + backward (data and weight gradients, all convs on tcgen05 in tensor_cores mode) + 2 Adam steps, CUDA events around 3 steps'}


if __name__ == '__main__':
    main()
