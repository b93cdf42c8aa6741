"""Host mirror of the reference's model graphs (src/model_types.py:179-416): CompressionModelV1 (factorized
prior) and CompressionModelV2 (scale hyperprior) with the same builder methods and block loops

    m.compress(x_shape); m.compress_blocks(sess, blocks, binstr, points, resolution, level, ...)
    m.decompress();      m.decompress_blocks(sess, blocks, x_shape, debug=False)
    m.train(x, gamma, alpha, lmbda)  -> m.train_loss / m.train_fl / m.train_mbpov (forward values)

There is no TF session: `sess` is accepted and ignored.  Unlike the reference (one sess.run per block,
model_types.py:192-198), blocks are processed in batches of `self.batch_size` on the GPU: densify ->
analysis -> [hyper] -> quantise -> synthesis -> clip/threshold/bit-pack all run as libpccgeo CUDA kernels; only
int32 symbols / indexes and packed occupancy bits cross PCIe; the range coder runs on host threads, one
independent stream per block and latent.
"""
import logging
from enum import Enum

import numpy as np
import torch

from . import ops
from .entropy_models import EntropyBottleneck, GaussianConditional, make_scale_table
from .focal_loss import focal_loss
from .model_transforms import TransformType

logger = logging.getLogger(__name__)


def sparse_to_dense(block, x_shape, data_format='channels_first'):
    """src/model_types.py:108-114 on the GPU: (n,3) integer coords -> fp32 occupancy of shape x_shape."""
    assert data_format == 'channels_first'
    b = np.asarray(block)[:, :3].astype(np.int16)
    coords = np.concatenate([np.zeros((len(b), 1), np.int16), b], axis=1)
    return ops.densify(torch.from_numpy(np.ascontiguousarray(coords)).cuda(), 1, *[int(s) for s in x_shape[2:]])


def blocks_to_coords(blocks):
    """list of (n_i, >=3) arrays -> one int16 (sum n_i, 4) array of (block, z, y, x) rows."""
    if not len(blocks):
        return np.zeros((0, 4), np.int16)
    lens = [len(b) for b in blocks]
    out = np.empty((sum(lens), 4), np.int16)
    out[:, 0] = np.repeat(np.arange(len(blocks), dtype=np.int16), lens)
    pos = 0
    for b, n in zip(blocks, lens):  # one slice assignment per block (float -> int16 cast inside numpy)
        out[pos:pos + n, 1:] = np.asarray(b)[:, :3]
        pos += n
    return out


def bits_to_points(bits_host, shape):
    """packed occupancy words (uint32 little-endian bit order) -> float32 (m,3) argwhere rows, C order."""
    occ = np.unpackbits(bits_host.view(np.uint8), bitorder='little').reshape(shape)
    return np.argwhere(occ).astype(np.float32)


def threshold_f32(thresholds, idx):
    """largest float32 <= the float64 threshold, so that (fp32 x > t32) == (x > t64) (model_types.py:209,233)."""
    t64 = np.asarray(thresholds, np.float64)[idx]
    t32 = t64.astype(np.float32)
    t32 = np.where(t32.astype(np.float64) > t64, np.nextafter(t32, np.float32(-np.inf)), t32)
    return t32.astype(np.float32)


class CompressionModel:
    def __init__(self, n_thresholds=2 ** 8, data_format='channels_first', batch_size=32):
        self.thresholds = np.linspace(0, 1.0, n_thresholds)  # model_types.py:181
        self.data_format = data_format
        self.batch_size = batch_size
        import os
        # host threads per range-coder / point-extraction call: a quarter of the cores, because `pipeline_depth` batches
        # are coded concurrently (measured on the 16-core B200 host: depth 4 x 4 threads is the sweet spot)
        self.coder_threads = max(1, (os.cpu_count() or 4) // 4)
        self.pipeline_depth = 4  # batches in flight (worker threads / CUDA streams) in the block loops
        self.x = self.x_hat = self.strings = self.debug_tensors = None
        self.x_shape = None

    # -- weights -------------------------------------------------------------------------------------
    def transforms(self):
        raise NotImplementedError

    def get_weights(self):
        """{'analysis': [...], 'synthesis': [...], ['hyper_*': ...], 'entropy_bottleneck': {...}} (Keras layouts)."""
        w = {k: t.get_weights() for k, t in self.transforms().items()}
        w['entropy_bottleneck'] = self.entropy_bottleneck.get_weights()
        return w

    def set_weights(self, w):
        for k, t in self.transforms().items():
            t.set_weights(w[k])
        self.entropy_bottleneck.set_weights(w['entropy_bottleneck'])

    # -- batched device passes (implemented by V1 / V2) ----------------------------------------------
    def _encode_device(self, x):
        raise NotImplementedError

    def _encode_host(self, dev):
        raise NotImplementedError

    def _decode_batch(self, strings_list, x_shape):
        raise NotImplementedError

    def _synthesize(self, y_hat, thresholds, want_x_hat):
        """x_hat = synthesis(y_hat); with thresholds also the packed occupancy (fused into the last layer when possible)."""
        if thresholds is None:
            return self.synthesis_transform(y_hat), None
        x_hat, bits, _ = self.synthesis_transform.packed(y_hat, thresholds, want_f32=want_x_hat)
        return x_hat, bits

    # -- batch pipeline ------------------------------------------------------------------------------
    def _map_batches(self, fn, batches):
        """Run fn(batch) for every batch, results in order.  With more than one batch, each batch's whole chain
        (H2D -> kernels -> D2H -> C++ range coding / point extraction) runs in a worker thread on its own CUDA stream,
        so the host stages of batch i overlap the GPU stages of batch i+1.  ctypes / torch release the GIL in the heavy
        calls; kernels are per-sample deterministic, so results do not depend on the schedule."""
        if len(batches) <= 1 or self.pipeline_depth <= 1:
            self._warmed = True
            return [fn(b) for b in batches]
        from concurrent.futures import ThreadPoolExecutor
        first = []
        if not getattr(self, '_warmed', False):  # fill the per-layer device caches single-threaded, once per model
            first = [fn(batches[0])]
            batches = batches[1:]
            self._warmed = True
        dev = torch.cuda.current_device()
        stream = torch.cuda.current_stream()

        def run(b):
            # all workers enqueue on ONE stream: the GPU runs the batches FIFO, so batch i's results (async D2H into
            # pinned buffers + an event) arrive while batch i+1 is still computing and the workers never fall in lockstep
            torch.cuda.set_device(dev)
            with torch.cuda.stream(stream):
                return fn(b)

        with ThreadPoolExecutor(max_workers=self.pipeline_depth) as pool:
            rest = list(pool.map(run, batches))
        return first + rest

    @staticmethod
    def _d2h(*tensors):
        """Enqueue async device->pinned-host copies on the current stream; returns (host tensors, event)."""
        outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        for o, t in zip(outs, tensors):
            o.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return outs, ev

    @staticmethod
    def _wait(pending):
        outs, ev = pending
        ev.synchronize()
        return [o.numpy() for o in outs]

    def _chunks(self, items):
        return [items[i:i + self.batch_size] for i in range(0, len(items), self.batch_size)]

    def _h2d(self, arr):
        """numpy -> CUDA through pinned memory (async on the current stream)."""
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.pin_memory().cuda(non_blocking=True) if t.numel() else t.cuda()

    # -- public block loops --------------------------------------------------------------------------
    def encode_blocks(self, blocks, x_shape=None, thr_idx=None, keep_x_hat=True):
        """Batched analysis + entropy coding + synthesis.
        Returns (strings per block, x_hat fp32 CUDA (n,1,D,H,W) or None, points per block or None).
        thr_idx (n,) fixes the per-block threshold so that clip/threshold/bit-pack and the point extraction ride in
        the same pipelined pass (compress_blocks with fixed_threshold)."""
        dims = [int(s) for s in (x_shape if x_shape is not None else self.x_shape)][-3:]
        spans = [(i, min(i + self.batch_size, len(blocks))) for i in range(0, len(blocks), self.batch_size)]

        def run(span):
            chunk = blocks[span[0]:span[1]]
            x = ops.densify(self._h2d(blocks_to_coords(chunk)), len(chunk), *dims)
            pend = {}
            # the symbol D2H is enqueued BEFORE synthesis is launched: range coding overlaps the synthesis kernels
            t = self._h2d(threshold_f32(self.thresholds, thr_idx[span[0]:span[1]])) if thr_idx is not None else None
            dev = self._encode_device(x, lambda d: pend.__setitem__('sym', self._d2h(*self._latent_tensors(d))),
                                      thresholds=t, want_x_hat=keep_x_hat)
            pts = None
            if thr_idx is not None:
                pend['bits'] = self._d2h(dev['bits'])
            strings = self._encode_host(dev, self._wait(pend['sym']))
            if thr_idx is not None:
                pts = ops.bits_to_points(self._wait(pend['bits'])[0], dims, self.coder_threads)
            return strings, (dev['x_hat'] if keep_x_hat else None), pts

        res = self._map_batches(run, spans)
        strings = [s for r in res for s in r[0]]
        x_hat = None
        if keep_x_hat:
            xs = [r[1] for r in res]
            torch.cuda.current_stream().synchronize()
            x_hat = torch.cat(xs) if len(xs) > 1 else xs[0]
        pts = [q for r in res for q in r[2]] if thr_idx is not None else None
        return strings, x_hat, pts

    def compress_blocks(self, sess, blocks, binstr, points, resolution, level, with_normals=False,
                        opt_metrics=('d1_mse',), max_deltas=(np.inf,), fixed_threshold=False, debug=False):
        """src/model_types.py:184-218.  Returns (data_list, metadata, debug_t_list)."""
        assert self.x_shape is not None, 'call compress(x_shape) first'
        n = len(blocks)
        if fixed_threshold:
            opt_metrics_ret = list(opt_metrics)
            thr_idx = np.full((n, len(opt_metrics_ret)), len(self.thresholds) // 2, np.int64)  # model_opt.py:27-31
            strings_list, _, pts = self.encode_blocks(blocks, thr_idx=thr_idx[:, 0], keep_x_hat=False)
            x_hat_list = [pts] * thr_idx.shape[1]  # every opt_metric gets the same fixed threshold
        else:
            strings_list, x_hat, _ = self.encode_blocks(blocks)
            thr_idx, opt_metrics_ret = self._optimal_thresholds(blocks, x_hat, resolution, with_normals, opt_metrics, max_deltas)
            x_hat_list = []
            for m in range(thr_idx.shape[1]):
                t = self._h2d(threshold_f32(self.thresholds, thr_idx[:, m]))
                bits, _ = ops.threshold_pack(x_hat, t)
                x_hat_list.append(ops.bits_to_points(bits.cpu().numpy(), tuple(x_hat.shape[2:]), self.coder_threads))
        threshold_list = [tuple(int(v) for v in thr_idx[:, m]) for m in range(thr_idx.shape[1])]
        metadata = self._select_best(binstr, x_hat_list, level, opt_metrics_ret, points, resolution, with_normals)
        data_list = [list(zip(strings_list, threshold_list[x['idx']])) for x in metadata]
        debug_t_list = [None] * n
        return data_list, metadata, debug_t_list

    def _optimal_thresholds(self, blocks, x_hat, resolution, with_normals, opt_metrics, max_deltas):
        """Per-block threshold search (reference src/model_opt.py:21-77) is host-side kd-tree work outside the hot
        path ("next" #1): reuse the reference's own module when it is importable."""
        try:
            from model_opt import compute_optimal_thresholds  # the reference's src/ on sys.path
        except ImportError as e:
            raise NotImplementedError('adaptive thresholds need the reference host module model_opt.py on sys.path '
                                      '(out of the hot path); pass fixed_threshold=True otherwise') from e
        xh = torch.clamp(x_hat[:, 0], 0.0, 1.0).cpu().numpy()
        idx, ret = [], None
        for j, block in enumerate(blocks):
            normals = block[:, block.shape[1] - 3:] if with_normals else None
            ret, best = compute_optimal_thresholds(block, xh[j], self.thresholds, resolution, normals=normals,
                                                   opt_metrics=opt_metrics, max_deltas=max_deltas, fixed_threshold=False)
            idx.append(best)
        return np.asarray(idx, np.int64), list(ret)

    def _select_best(self, binstr, x_hat_list, level, opt_metrics, points, resolution, with_normals):
        """select_best_per_opt_metric (model_types.py:128-176) needs the reference's octree + metric host modules;
        without them the first opt_metric is selected and no metrics are reported."""
        try:
            from model_types import select_best_per_opt_metric  # noqa: the reference's own host code
            return select_best_per_opt_metric(binstr, x_hat_list, level, opt_metrics, points, resolution, with_normals)
        except Exception:
            return [{'idx': 0, 'metrics': {}, 'x_hat_list': x_hat_list[0], 'blocks_depart': None, 'blocks_full': None}]

    def decompress_blocks(self, sess, blocks, x_shape, debug=False):
        """src/model_types.py:220-238: blocks = [(strings, threshold_idx)] -> ([float32 (m,3)], debug list)."""
        dims = tuple(int(s) for s in x_shape)[-3:]

        def run(chunk):
            strings = [c[0] for c in chunk]
            idx = np.asarray([int(c[1]) for c in chunk], np.int64)
            x_hat, dbg = self._decode_batch(strings, dims, thresholds=self._h2d(threshold_f32(self.thresholds, idx)),
                                            want_x_hat=debug)
            pts = ops.bits_to_points(self._wait(self._d2h(dbg['bits']))[0], dims, self.coder_threads)
            return pts, [dbg if debug else None] * len(chunk)

        res = self._map_batches(run, self._chunks(list(blocks)))
        return [p for r in res for p in r[0]], [d for r in res for d in r[1]]

    # -- training graph (forward values; see DESIGN.md for the backward status) ----------------------
    def _finish_train(self, x, x_tilde, log_sums, gamma, alpha, lmbda):
        n_occ = x.sum(dtype=torch.float64)
        denom = -np.log(2) * n_occ
        mb = [s[0] / denom for s in log_sums]
        self.train_mbpov = sum(mb[1:], mb[0])
        self.train_fl = focal_loss(x, x_tilde, gamma=gamma, alpha=alpha)
        self.train_loss = lmbda * self.train_fl + self.train_mbpov
        self.num_occupied_voxels = n_occ
        self.x_tilde = x_tilde
        self.merged_summary = {'loss': self.train_loss, 'fl': self.train_fl, 'mbpov/total': self.train_mbpov,
                               'num_occupied_voxels': n_occ}
        self.step = getattr(self, 'step', 0)
        # sess.run(m.train_op) of the reference == m.train_op(x): forward + backward + both Adam steps + table refresh
        from .training import Trainer
        if getattr(self, 'trainer', None) is None:
            self.trainer = Trainer(self, gamma, alpha, lmbda)
        self.train_op = self.trainer.step
        return mb


class CompressionModelV1(CompressionModel):
    def __init__(self, num_filters=32, analysis_transform_type=TransformType.AnalysisTransformV1,
                 synthesis_transform_type=TransformType.SynthesisTransformV1, *args, **kwargs):
        self.num_filters = num_filters
        self.analysis_transform_class = analysis_transform_type.value
        self.synthesis_transform_class = synthesis_transform_type.value
        super().__init__(*args, **kwargs)
        self.analysis_transform = self.analysis_transform_class(num_filters, data_format=self.data_format)
        self.synthesis_transform = self.synthesis_transform_class(num_filters, data_format=self.data_format)
        self.entropy_bottleneck = EntropyBottleneck(data_format=self.data_format)
        self.entropy_bottleneck.build(num_filters)

    def transforms(self):
        return {'analysis': self.analysis_transform, 'synthesis': self.synthesis_transform}

    def train(self, x, gamma, alpha, lmbda, noise_y=None):  # model_types.py:250-281
        y = self.analysis_transform(x)
        y_tilde, _ = self.entropy_bottleneck(y, training=True, noise=noise_y)
        x_tilde = self.synthesis_transform(y_tilde)
        self._finish_train(x, x_tilde, [self.entropy_bottleneck.log_likelihood_sum(y_tilde)], gamma, alpha, lmbda)
        self.y, self.y_tilde = y, y_tilde

    def compress(self, x_shape):  # model_types.py:283-295
        self.x_shape = tuple(int(s) for s in x_shape)

    def decompress(self):  # model_types.py:297-309
        pass

    def _encode_device(self, x, after_latents=None, thresholds=None, want_x_hat=True):
        y = self.analysis_transform(x)
        y_sym, y_hat = self.entropy_bottleneck.quantize(y)
        dev = {'y_sym': y_sym, 'y_hat': y_hat}
        if after_latents is not None:
            after_latents(dev)
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.x, self.x_hat = x, x_hat
        self.debug_tensors = {'y_hat': y_hat, 'x_hat': x_hat}
        dev['x_hat'], dev['bits'] = x_hat, bits
        return dev

    @staticmethod
    def _latent_tensors(dev):
        return (dev['y_sym'],)

    def _encode_host(self, dev, host=None):
        y_sym = host[0] if host is not None else dev['y_sym'].cpu().numpy()
        ys = self.entropy_bottleneck.encode_symbols(y_sym, self.coder_threads)
        return [(s,) for s in ys]

    def _decode_batch(self, strings_list, dims, thresholds=None, want_x_hat=True):
        f = self.num_filters
        shp = (f,) + tuple(d // 8 for d in dims)  # model_types.py:305
        sym = self.entropy_bottleneck.decode_symbols([s[0] for s in strings_list], shp, self.coder_threads)
        y_hat = ops.eb_dequantize(self._h2d(sym), self.entropy_bottleneck.device_params())
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.x_hat = x_hat
        return x_hat, {'y_hat': y_hat, 'x_hat': x_hat, 'bits': bits}


class CompressionModelV2(CompressionModel):
    def __init__(self, num_filters=32, analysis_transform_type=TransformType.AnalysisTransformV1,
                 synthesis_transform_type=TransformType.SynthesisTransformV1,
                 hyper_analysis_transform_type=TransformType.HyperAnalysisTransform,
                 hyper_synthesis_transform_type=TransformType.HyperSynthesisTransform,
                 scales_min=0.11, scales_max=256, scales_levels=64, *args, **kwargs):
        self.num_filters = num_filters
        self.analysis_transform_class = analysis_transform_type.value
        self.synthesis_transform_class = synthesis_transform_type.value
        self.hyper_analysis_transform_class = hyper_analysis_transform_type.value
        self.hyper_synthesis_transform_class = hyper_synthesis_transform_type.value
        self.scale_table = make_scale_table(scales_min, scales_max, scales_levels)  # model_types.py:324
        super().__init__(*args, **kwargs)
        df = self.data_format
        self.analysis_transform = self.analysis_transform_class(num_filters, data_format=df)
        self.synthesis_transform = self.synthesis_transform_class(num_filters, data_format=df)
        self.hyper_analysis_transform = self.hyper_analysis_transform_class(num_filters, data_format=df)
        self.hyper_synthesis_transform = self.hyper_synthesis_transform_class(num_filters, data_format=df)
        self.entropy_bottleneck = EntropyBottleneck(data_format=df)
        self.entropy_bottleneck.build(num_filters)

    def transforms(self):
        return {'analysis': self.analysis_transform, 'synthesis': self.synthesis_transform,
                'hyper_analysis': self.hyper_analysis_transform, 'hyper_synthesis': self.hyper_synthesis_transform}

    def train(self, x, gamma, alpha, lmbda, noise_y=None, noise_z=None):  # model_types.py:327-369
        y = self.analysis_transform(x)
        z = self.hyper_analysis_transform(y)
        z_tilde, _ = self.entropy_bottleneck(z, training=True, noise=noise_z)
        sigma_tilde = self.hyper_synthesis_transform(z_tilde)
        cb = GaussianConditional(sigma_tilde, self.scale_table, data_format=self.data_format)
        y_tilde, _ = cb(y, training=True, noise=noise_y)
        x_tilde = self.synthesis_transform(y_tilde)
        mb = self._finish_train(x, x_tilde, [cb.log_likelihood_sum(y_tilde), self.entropy_bottleneck.log_likelihood_sum(z_tilde)],
                                gamma, alpha, lmbda)
        self.train_mbpov_y, self.train_mbpov_z = mb
        self.y, self.z, self.y_tilde, self.z_tilde, self.sigma_tilde = y, z, y_tilde, z_tilde, sigma_tilde

    def compress(self, x_shape):  # model_types.py:371-391
        self.x_shape = tuple(int(s) for s in x_shape)

    def decompress(self):  # model_types.py:393-411
        pass

    def _encode_device(self, x, after_latents=None, thresholds=None, want_x_hat=True):
        y = self.analysis_transform(x)
        z = self.hyper_analysis_transform(y)
        z_sym, z_hat = self.entropy_bottleneck.quantize(z)
        sigma_hat = self.hyper_synthesis_transform(z_hat)
        cb = GaussianConditional(sigma_hat, self.scale_table, data_format=self.data_format)
        y_sym, y_hat, idx = cb.quantize(y)
        dev = {'y': y, 'z': z, 'z_sym': z_sym, 'z_hat': z_hat, 'sigma_hat': sigma_hat, 'y_sym': y_sym, 'y_hat': y_hat,
               'indexes': idx, 'cb': cb}
        if after_latents is not None:
            after_latents(dev)
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.x, self.x_hat = x, x_hat
        self.debug_tensors = {'z_hat': z_hat, 'sigma_hat': sigma_hat, 'indexes': idx, 'y_hat': y_hat, 'x_hat': x_hat}
        dev['x_hat'], dev['bits'] = x_hat, bits
        return dev

    @staticmethod
    def _latent_tensors(dev):
        return (dev['z_sym'], dev['y_sym'], dev['indexes'])

    def _encode_host(self, dev, host=None):
        z_sym, y_sym, idx = host if host is not None else [t.cpu().numpy() for t in self._latent_tensors(dev)]
        zs = self.entropy_bottleneck.encode_symbols(z_sym, self.coder_threads)
        ys = dev['cb'].encode_symbols(y_sym, idx, self.coder_threads)
        return list(zip(ys, zs))  # (y_string, z_string): model_types.py:389

    def _decode_batch(self, strings_list, dims, thresholds=None, want_x_hat=True):
        f = self.num_filters
        zshp = (f,) + tuple(d // 16 for d in dims)  # model_types.py:403
        zsym = self.entropy_bottleneck.decode_symbols([s[1] for s in strings_list], zshp, self.coder_threads)
        z_hat = ops.eb_dequantize(self._h2d(zsym), self.entropy_bottleneck.device_params())
        sigma_hat = self.hyper_synthesis_transform(z_hat)
        cb = GaussianConditional(sigma_hat, self.scale_table, data_format=self.data_format)
        idx = cb.indexes()
        ysym = cb.decode_symbols([s[0] for s in strings_list], self._wait(self._d2h(idx))[0], self.coder_threads)
        y_hat = ops.i32_to_f32(self._h2d(ysym))
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.x_hat = x_hat
        return x_hat, {'z_hat': z_hat, 'sigma_hat': sigma_hat, 'indexes': idx, 'y_hat': y_hat, 'x_hat': x_hat, 'bits': bits}


class ModelType(Enum):
    v1 = CompressionModelV1
    v2 = CompressionModelV2
