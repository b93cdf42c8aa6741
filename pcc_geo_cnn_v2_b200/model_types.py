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
import os
from enum import Enum

import numpy as np
import torch

from . import ops
from .entropy_models import EntropyBottleneck, GaussianConditional, make_scale_table, gaussian_tables
from ._lib import PccGeoError
from .focal_loss import focal_loss
from . import model_transforms as _mt
from .model_transforms import TransformType

logger = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------------------
# host glue of the training script (reference src/model_types.py:23-62,65-105,121-125), same names and argument meaning
# ---------------------------------------------------------------------------------------------------------
class SparseBlock:
    """Stand-in for tf.sparse.SparseTensor(indices, ones, dense_shape) as pc_to_tf builds it."""

    def __init__(self, indices, dense_shape):
        self.indices, self.dense_shape = np.asarray(indices, np.int64), tuple(int(s) for s in dense_shape)


def pc_to_tf(points, dense_tensor_shape, data_format):  # model_types.py:23-32
    x = np.asarray(points, np.int64)
    assert data_format in ['channels_last', 'channels_first']
    pad = [[0, 0], [0, 1]] if data_format == 'channels_last' else [[0, 0], [1, 0]]  # the channel index (always 0)
    return SparseBlock(np.pad(x, pad), dense_tensor_shape)


def process_x(x, dense_tensor_shape):  # model_types.py:35-39: sparse -> dense fp32, on the GPU (densify kernel)
    shape = tuple(int(s) for s in dense_tensor_shape)
    assert shape == x.dense_shape
    last = shape[-1] == 1 and shape[0] != 1
    sp = x.indices[:, :3] if last else x.indices[:, 1:]
    dims = shape[:3] if last else shape[1:]
    coords = np.concatenate([np.zeros((len(sp), 1), np.int16), sp.astype(np.int16)], axis=1)
    dense = ops.densify(torch.from_numpy(np.ascontiguousarray(coords)).cuda() if len(sp) else None, 1, *dims)
    return dense.reshape(shape)


def quantize_tensor(x):  # model_types.py:42-46 (TensorBoard summaries only)
    return torch.round(torch.clamp(x, 0, 1)).to(torch.uint8)


def add_channels(shape, channels, data_format):  # model_types.py:121-125
    shape = [int(s) for s in shape]
    return [int(channels)] + shape if data_format == 'channels_first' else shape + [int(channels)]


def get_normals_if(x, with_normals):  # model_types.py:117-118
    return x[:, x.shape[1] - 3:x.shape[1]] if with_normals else None


def input_fn(points, batch_size, dense_tensor_shape, data_format, repeat=True, shuffle=True, prefetch_size=1):
    """model_types.py:49-62 as a Python generator of (B, ...) fp32 CUDA batches: shuffle over the whole set every epoch
    (numpy's global RNG, seeded by the caller like tr_train.py:20), repeat, then batch -- so with repeat=True batches span
    epoch boundaries and are always full, and only without repeat the last batch is partial (tf.data's batch()) -- and `prefetch_size` batches packed (C++ coordinate packer) ahead on a host thread."""
    import queue
    import threading
    points = list(points)
    shape = tuple(int(s) for s in dense_tensor_shape)
    last = data_format == 'channels_last'
    dims = shape[:3] if last else shape[1:]

    def batches():
        # shuffle(len) -> repeat -> batch: the element stream runs across epoch boundaries, so with repeat=True every batch is
        # full (a batch may hold the tail of one epoch and the head of the next); without repeat the last batch is partial
        pending = []
        while True:
            order = np.random.permutation(len(points)) if shuffle else np.arange(len(points))
            for j in order:
                pending.append(np.asarray(points[j], np.float32))
                if len(pending) == batch_size:
                    yield len(pending), blocks_to_coords(pending)
                    pending = []
            if not repeat:
                if pending:
                    yield len(pending), blocks_to_coords(pending)
                return
            if not len(points):
                return

    q = queue.Queue(maxsize=max(1, int(prefetch_size)))

    def producer():
        for item in batches():
            q.put(item)
        q.put(None)

    threading.Thread(target=producer, daemon=True).start()
    while True:
        item = q.get()
        if item is None:
            return
        n, coords = item
        x = ops.densify(torch.from_numpy(coords).cuda() if len(coords) else None, n, *dims)
        yield x.reshape((n,) + shape)


def binary_classification_summaries(x_quant, x_tilde_quant):  # model_types.py:91-105
    xq, xt = x_quant.to(torch.float32), x_tilde_quant.to(torch.float32)
    tp = torch.count_nonzero(xt * xq).double()
    tn = torch.count_nonzero((xt - 1) * (xq - 1)).double()
    fp = torch.count_nonzero(xt * (xq - 1)).double()
    fn = torch.count_nonzero((xt - 1) * xq).double()
    precision, recall = tp / (tp + fp), tp / (tp + fn)
    return {'bc/precision': precision, 'bc/recall': recall, 'bc/accuracy': (tp + tn) / (tp + tn + fp + fn),
            'bc/specificity': tn / (tn + fp), 'bc/f1_score': (2 * precision * recall) / (precision + recall)}


def v1_summaries(train_loss, mbpov_y, mbpov_total, train_fl, log_y_likelihoods, num_occupied_voxels, x, x_tilde,
                 x_tilde_quant, y, y_likelihoods, y_tilde):  # model_types.py:65-79: name -> scalar / histogram source tensor
    return {'loss': train_loss, 'mbpov/y': mbpov_y, 'mbpov/total': mbpov_total, 'fl': train_fl,
            'num_occupied_voxels': num_occupied_voxels, 'y': y, 'y_tilde': y_tilde, 'x': x, 'x_tilde': x_tilde,
            'x_tilde_quant': x_tilde_quant, 'y_likelihoods': y_likelihoods, 'log_y_likelihoods': log_y_likelihoods}


def v2_summaries(log_z_likelihoods, sigma_tilde, train_mbpov_z, z, z_likelihoods, z_tilde):  # model_types.py:82-88
    return {'z': z, 'z_tilde': z_tilde, 'mbpov/z': train_mbpov_z, 'sigma_tilde': sigma_tilde,
            'z_likelihoods': z_likelihoods, 'log_z_likelihoods': log_z_likelihoods}


def select_best_per_opt_metric(binstr, x_hat_list, level, opt_metrics, points, resolution, with_normals, opt_groups=('d1', 'd2')):
    """src/model_types.py:128-176: per metric group, re-assemble every candidate reconstruction of the whole cloud
    (departition_octree), score it against the original points and keep the best by `<group>_psnr`.  Host code (kd-trees)
    like the reference)."""
    from scipy.spatial import cKDTree
    from .octree_coding import departition_octree
    from .pc_metric import compute_metrics
    assert len(opt_metrics) == len(x_hat_list), (f'lengths of opt_metrics {len(opt_metrics)} and x_hat_list'
                                                 f' {len(x_hat_list)} should be equal')
    om_groups = [[(x, y, i) for i, (x, y) in enumerate(zip(opt_metrics, x_hat_list)) if x.startswith(group)] for group in opt_groups]
    points = np.asarray(points)
    bbox_min, bbox_max = [0, 0, 0], [resolution] * 3
    t1 = cKDTree(points[:, :3])
    metadata = []
    for group, om_group in zip(opt_groups, om_groups):
        if len(om_group) == 0:
            continue
        metric_key = f'{group}_psnr'
        om_names, cur_x_hat_list, indexes = zip(*om_group)
        cur_blocks_depart = [departition_octree(x, list(binstr), bbox_min, bbox_max, level) for x in cur_x_hat_list]
        cur_blocks_full = [np.vstack(x) for x in cur_blocks_depart]
        normals = points[:, points.shape[1] - 3:] if with_normals else None   # get_normals_if, model_types.py:117-118
        cur_metrics_full = [compute_metrics(points[:, :3], x, resolution - 1, p1_n=normals, t1=t1) for x in cur_blocks_full]
        local_best_idx = int(np.argmax([x[metric_key] for x in cur_metrics_full]))
        metadata.append({'idx': indexes[local_best_idx], 'metrics': cur_metrics_full[local_best_idx],
                         'x_hat_list': cur_x_hat_list[local_best_idx], 'blocks_depart': cur_blocks_depart[local_best_idx],
                         'blocks_full': cur_blocks_full[local_best_idx]})
    return metadata


def sparse_to_dense(block, x_shape, data_format='channels_first'):
    """src/model_types.py:108-114 on the GPU: (n,3) integer coords -> fp32 occupancy of shape x_shape."""
    assert data_format == 'channels_first'
    b = np.asarray(block)[:, :3].astype(np.int16)
    coords = np.concatenate([np.zeros((len(b), 1), np.int16), b], axis=1)
    return ops.densify(torch.from_numpy(np.ascontiguousarray(coords)).cuda(), 1, *[int(s) for s in x_shape[2:]])


def blocks_to_coords(blocks, threads=1, staged=False):
    """list of (n_i, >=3) arrays -> one int16 (sum n_i, 4) array of (block, z, y, x) rows (C++ host helper).
    staged=True writes the rows straight into a pinned staging buffer and returns (pinned int16 torch view, pool buffer):
    the block loops' driver thread then only enqueues the H2D copy instead of copying megabytes itself."""
    n = len(blocks)
    if not n:
        if staged:
            buf = _pinned.get(1)
            return buf[:0].view(torch.int16).view(0, 4), buf
        return np.zeros((0, 4), np.int16)
    arrs = [np.asarray(b) for b in blocks]
    f64 = any(a.dtype == np.float64 for a in arrs)
    want = np.float64 if f64 else np.float32
    for i, a in enumerate(arrs):
        if a.ndim != 2 or a.shape[1] < 3:
            raise ValueError(f'block {i}: expected (n, >=3) coordinates, got {a.shape}')
        if a.dtype != want or (len(a) and a.strides[1] != a.itemsize):
            arrs[i] = np.ascontiguousarray(a, want)
    counts = np.array([len(a) for a in arrs], np.int64)
    pitch = np.array([a.strides[0] if len(a) else 0 for a in arrs], np.int64)
    ptrs = np.array([a.__array_interface__['data'][0] if len(a) else 0 for a in arrs], np.uint64)
    rows = int(counts.sum())
    if staged:
        buf = _pinned.get(max(rows * 8, 1))
        host = buf[:rows * 8].view(torch.int16).view(rows, 4)
        out = host.numpy()
    else:
        out = np.empty((rows, 4), np.int16)
    from . import _lib as L
    L.check(L.lib().pccgeo_blocks_to_coords_host(L.ptr(ptrs), L.ptr(counts), L.ptr(pitch), n, int(f64), L.ptr(out), int(threads)),
            'blocks_to_coords')
    return (host, buf) if staged else out


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


class GraphHandle:
    """Stand-in for the TF placeholders / tensors the reference's builder methods leave on the model (m.x, m.x_hat, m.strings,
    m.strings_t, m.x_shape_t; src/model_types.py:290-295,378-391,399-411).  There is no graph to feed: the block loops take the
    data directly.  The handles carry what the scripts read from them -- the name and the static shape (compress_blocks uses
    self.x.shape, model_types.py:194)."""

    def __init__(self, name, shape=None, dtype='float32'):
        self.name, self.shape, self.dtype = name, (None if shape is None else tuple(int(v) for v in shape)), dtype

    def __repr__(self):
        return f'<GraphHandle {self.name} shape={self.shape} dtype={self.dtype}>'


def split_debug(batch_tensors, strings, static=None):
    """{key: CUDA tensor with a leading batch axis} -> one dict of numpy arrays per block, each with batch axis 1 -- the
    per-block `debug_tensors` a sess.run of the reference's batch-1 graph returns (model_types.py:198-199,229).  `strings`
    (per block tuple of bytes) become the 'decompress/strings' entry of the Gaussian conditional's dbg_dec
    (patch_gaussian_conditional.py:43-45); `static` entries (tables) are shared between the blocks."""
    host = {k: v.detach().cpu().numpy() for k, v in batch_tensors.items() if v is not None}
    n = len(strings)
    out = []
    for j in range(n):
        d = {k: v[j:j + 1] for k, v in host.items()}
        if static:
            d.update(static)
            d['decompress/strings'] = np.array([strings[j][0]], dtype=object)
        out.append(d)
    return out


class _DeviceSpan:
    """One batch of a DeviceBlocks input, with the future-like surface of the host packing tasks."""

    def __init__(self, coords, block0):
        self.coords, self.block0 = coords, block0

    def result(self):
        return self.coords


class _PinnedPool:
    """Recycled pinned host staging buffers.  cudaHostAlloc costs milliseconds and stalls the driver thread of the block
    loops (torch's caching host allocator falls back to it whenever every cached block still has a copy in flight), so
    the pipelines take their staging memory from here: power-of-two byte buffers, handed back explicitly by the consumer
    (D2H) or after the event behind an H2D copy has fired."""

    def __init__(self):
        import threading
        self.free, self.pending, self.lock = {}, [], threading.Lock()

    def get(self, nbytes):
        cap = 256
        while cap < nbytes:
            cap *= 2
        with self.lock:
            if self.pending:
                still = []
                for buf, ev in self.pending:
                    if ev.query():
                        self.free.setdefault(buf.numel(), []).append(buf)
                    else:
                        still.append((buf, ev))
                self.pending = still
            lst = self.free.get(cap)
            if lst:
                return lst.pop()
        return torch.empty(cap, dtype=torch.uint8, pin_memory=True)

    def put(self, buf):
        with self.lock:
            self.free.setdefault(buf.numel(), []).append(buf)

    def put_after(self, buf, event):
        with self.lock:
            self.pending.append((buf, event))


_pinned = _PinnedPool()

graph_kernel_launches = [0]   # libpccgeo kernels launched through graph replays (bench.py adds them to gpu_launches)


class _StageGraph:
    """fn() captured once into a CUDA graph (after two eager warm-up runs that fill the per-layer device caches); replay()
    re-runs the whole kernel sequence with ONE launch call -- the Python driver of the block loops then costs microseconds
    per batch instead of ~60 ctypes launches, and the kernels run back to back without launch gaps.  fn must read its
    inputs from tensors that outlive the graph (static buffers or outputs of other stage graphs)."""

    def __init__(self, fn):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()
            fn()
        cur.wait_stream(side)
        side.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib as L
        l0 = L.lib().pccgeo_launch_count()
        # thread_local: host workers of the pipeline may call cudaEventSynchronize / pinned allocations meanwhile
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.out = fn()
        self.n_kernels = int(L.lib().pccgeo_launch_count() - l0)   # libpccgeo kernels one replay launches

    def replay(self):
        self.graph.replay()
        graph_kernel_launches[0] += self.n_kernels
        return self.out


class CompressionModel:
    def __init__(self, n_thresholds=2 ** 8, data_format='channels_first', batch_size=32):
        self.thresholds = np.linspace(0, 1.0, n_thresholds)  # model_types.py:181
        self.data_format = data_format
        self.batch_size = batch_size
        import os
        # host threads: this rank's share of the cores (one process per GPU under torchrun); `pipeline_depth` host workers
        # run the C++ stages of different batches concurrently, each call with `coder_threads` threads (sweep on the
        # 16-core B200 host: 3 workers x 8 threads)
        try:
            usable = len(os.sched_getaffinity(0))   # respects taskset / cgroup cpusets (os.cpu_count() does not)
        except AttributeError:
            usable = os.cpu_count() or 4
        cores = max(1, usable // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1'))))
        self.coder_threads = max(1, cores // 2)
        self.pipeline_depth = 3
        self.x = self.x_hat = self.strings = self.debug_tensors = None
        self.x_shape = None
        self.use_graphs = True   # capture the per-batch kernel sequences of the block loops into CUDA graphs
        # entropy-code on the GPU (csrc/rc_device.cu: one warp per stream, all streams of up to `coder_group_blocks` blocks in
        # one launch) instead of in the host workers; byte-identical strings.  Used by the CUDA-graph block loops.
        # Default: on when this rank has at most 8 host cores (several GPUs per host).  Measured on the rate-realistic
        # workload at the end of round 2 (blocks/s end to end, host / device coder, `taskset` on one box): 16 cores 7.6 k / 7.2 k,
        # 12 cores 7.4 k / 7.0 k, 8 cores 7.0 k / 7.1 k; 8 ranks x 4 cores 33.9 k / 52.0 k in total -- the device coder does not
        # depend on the host at all, the host coder is free as long as cores are idle.
        # PCCGEO_DEVICE_CODER=0/1 overrides.  A group's coding costs a fixed few milliseconds (the length
        # of one stream's serial chain), so groups are large; `coder_overlap` moves it to a side stream under the next
        # group's transforms (off: its one-warp CTAs displace the persistent conv CTAs and cost more than they hide).
        env = os.environ.get('PCCGEO_DEVICE_CODER', '')
        self.device_coder = env == '1' if env in ('0', '1') else cores <= 8
        self.coder_group_blocks = 1024
        self.coder_overlap = False
        self.symbol_bytes = self.index_bytes = 4   # width of the symbols / scale indexes that cross PCIe with the host coder
        self._graphs, self._statics, self._graph_epoch = {}, {}, -1
        self._pipe_evs = {}   # events of the decode pipeline's lanes (reset by every decompress_blocks call)

    # -- weights -------------------------------------------------------------------------------------
    def transforms(self):
        raise NotImplementedError

    def _sync_trainer(self):
        """After train_op steps the trained conv weights live in the trainer's device buffers; bring the layer objects up to
        date before anything else reads them (validation forward, get_weights, the codec loops)."""
        tr = getattr(self, 'trainer', None)
        if tr is not None and tr.dirty:
            tr.sync_to_model()

    def get_weights(self):
        """{'analysis': [...], 'synthesis': [...], ['hyper_*': ...], 'entropy_bottleneck': {...}} (Keras layouts)."""
        self._sync_trainer()
        w = {k: t.get_weights() for k, t in self.transforms().items()}
        w['entropy_bottleneck'] = self.entropy_bottleneck.get_weights()
        return w

    def set_weights(self, w):
        for k, t in self.transforms().items():
            t.set_weights(w[k])
        self.entropy_bottleneck.set_weights(w['entropy_bottleneck'])
        self.trainer = None   # optimiser state belongs to the replaced variables

    def load_weights(self, source):
        """What `saver.restore(sess, checkpoint)` does in the reference's scripts (compress_octree.py:82-92): `source` is an
        .npz path or a dict keyed by the TF variable names (weights_io.py documents the name map)."""
        from . import weights_io
        return weights_io.load_weights(self, source)

    def save_weights(self, path):
        from . import weights_io
        weights_io.save_weights(self, path)

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

    # -- CUDA-graph stages ---------------------------------------------------------------------------
    def _stage(self, name, n, dims, fn):
        """Replay (capturing on first use) the graph of stage `name` for batches of n blocks of size dims."""
        if self._graph_epoch != _mt.params_epoch[0]:  # parameters changed: packed weights were re-uploaded
            self._graphs.clear()
            self._graph_epoch = _mt.params_epoch[0]
        key = (name, n, tuple(dims), _mt.get_precision(), torch.cuda.current_device(), getattr(self, 'lane', 0))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = _StageGraph(fn)
        return g.replay()

    def _static(self, n, dims):
        """Static device buffers the stage graphs of (n, dims) read their per-batch inputs from."""
        # `lane`: independent sets of static buffers / stage graphs, so that two chains of batches can be in flight on two streams
        key = (n, tuple(dims), torch.cuda.current_device(), getattr(self, 'lane', 0))
        st = self._statics.get(key)
        if st is None:
            st = self._statics[key] = {'x': torch.zeros((n, 1) + tuple(dims), device='cuda'),
                                       'thr': torch.zeros(n, device='cuda'), **self._static_latents(n, dims)}
            # the zero fills run on the stream that is current NOW: whoever writes these buffers from another stream first waits here
            st['ready'] = torch.cuda.Event()
            st['ready'].record()
        return st

    def device_encode(self, coords_dev, n, dims, thr, block0=None):
        """densify -> analysis [-> hyper] -> quantise (latents) -> synthesis + threshold + pack, as stage graphs.
        coords_dev: CUDA int16 (npts,4); thr: CUDA fp32 (n,) or a host array.  -> (latent dict, bits).  `on_latents`
        work (the symbol D2H) goes between the two graphs: see encode_blocks."""
        st = self._static(n, dims)
        ops.densify(coords_dev, n, *dims, out=st['x'], block0=block0)
        lat = self._stage('latents', n, dims, lambda: self._latents(st['x']))
        return lat, st

    def device_synthesis(self, lat, st, n, dims, thr):
        st['thr'].copy_(thr) if torch.is_tensor(thr) else self._copy_in(st['thr'], thr)
        return self._stage('synthesis', n, dims, lambda: self._synthesize(lat['y_hat'], st['thr'], False)[1])

    # -- batch pipeline ------------------------------------------------------------------------------
    # The block loops are software pipelines over batches with ONE driver (the calling thread) that owns the CUDA stream
    # and enqueues every batch's kernels + async pinned copies back to back, and a pool of host workers that run the
    # C++ stages (coordinate packing, range coding, point extraction; ctypes releases the GIL) as soon as the event
    # behind their input fires.  The GPU never waits for the host between batches, host coding of batch i overlaps the
    # kernels of batch i+1..., and results do not depend on the schedule (kernels are per-sample deterministic).
    def _pool(self):
        if getattr(self, '_executor', None) is None:
            from concurrent.futures import ThreadPoolExecutor
            # pool threads do not inherit the caller's CUDA device: without the initializer their streams / events / copies
            # would be created on device 0 whatever GPU this rank drives
            self._executor = ThreadPoolExecutor(max_workers=max(1, self.pipeline_depth), thread_name_prefix='pccgeo-host',
                                                initializer=torch.cuda.set_device, initargs=(torch.cuda.current_device(),))
        return self._executor

    @staticmethod
    def _stage_host(arr):
        """numpy array -> (pinned torch view with the same contents, pool buffer)."""
        a = np.ascontiguousarray(arr)
        buf = _pinned.get(max(a.nbytes, 1))
        t = torch.from_numpy(a)
        host = buf[:a.nbytes].view(t.dtype).view(t.shape)
        # plain single-threaded memcpy: torch's CPU copy_ would enter its OpenMP pool, which stalls for milliseconds while the
        # host workers keep every core busy with range coding
        np.copyto(host.numpy(), a)
        return host, buf

    @staticmethod
    def _d2h(*tensors):
        """Enqueue async device->pinned-host copies on the current stream; returns (host tensors, event, pool buffers)."""
        outs, bufs = [], []
        for t in tensors:
            nb = t.numel() * t.element_size()
            buf = _pinned.get(max(nb, 1))
            o = buf[:nb].view(t.dtype).view(t.shape)
            o.copy_(t, non_blocking=True)
            outs.append(o)
            bufs.append(buf)
        ev = torch.cuda.Event()
        ev.record()
        return outs, ev, bufs

    @staticmethod
    def _wait(pending):
        pending[1].synchronize()
        return [o.numpy() for o in pending[0]]

    @staticmethod
    def _release(pending):
        """Hand the staging buffers of a finished _d2h back (the numpy views from _wait must not be used afterwards)."""
        for buf in pending[2]:
            _pinned.put(buf)

    def _chunks(self, items):
        return [items[i:i + self.batch_size] for i in range(0, len(items), self.batch_size)]

    def _copy_in(self, dst, arr):
        """numpy -> existing CUDA tensor through recycled pinned memory (async on the current stream)."""
        host, buf = self._stage_host(arr)
        dst.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _pinned.put_after(buf, ev)
        return dst

    @staticmethod
    def _h2d_staged(staged):
        """(pinned host tensor, pool buffer) from a host worker -> new CUDA tensor (async on the current stream)."""
        if torch.is_tensor(staged):   # already on the device (octree partition on the GPU)
            return staged
        host, buf = staged
        dst = torch.empty(host.shape, dtype=host.dtype, device='cuda')
        if host.numel():
            dst.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _pinned.put_after(buf, ev)
        return dst

    def _h2d(self, arr):
        """numpy -> new CUDA tensor through recycled pinned memory (async on the current stream)."""
        a = np.asarray(arr)
        dst = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, a.dtype)).dtype, device='cuda')
        return self._copy_in(dst, a) if a.size else dst

    # -- public block loops --------------------------------------------------------------------------
    def encode_blocks(self, blocks, x_shape=None, thr_idx=None, keep_x_hat=True, debug_out=None):
        """Batched analysis + entropy coding + synthesis.
        Returns (strings per block, x_hat fp32 CUDA (n,1,D,H,W) or None, points per block or None).
        thr_idx (n,) fixes the per-block threshold so that clip/threshold/bit-pack (fused into the last synthesis layer) and
        the point extraction ride in the same pipelined pass (compress_blocks with fixed_threshold)."""
        self._sync_trainer()
        dims = [int(s) for s in (x_shape if x_shape is not None else self.x_shape)][-3:]
        spans = [(i, min(i + self.batch_size, len(blocks))) for i in range(0, len(blocks), self.batch_size)]
        pool = self._pool()
        from .octree_coding import DeviceBlocks
        if isinstance(blocks, DeviceBlocks):
            # blocks partitioned on the GPU (octree_coding.partition_octree_gpu): batches are slices of the grouped device rows
            coords_f = [_DeviceSpan(blocks.coords[int(blocks.offsets[a]):int(blocks.offsets[b])], a) for a, b in spans]
        else:
            coords_f = [pool.submit(blocks_to_coords, blocks[a:b], self.coder_threads, True) for a, b in spans]

        def post(dev, pend):
            strings = self._encode_host(dev, self._wait(pend['sym']))
            self._release(pend['sym'])
            pts = None
            if 'bits' in pend:
                pts = ops.bits_to_points(self._wait(pend['bits'])[0], dims, self.coder_threads)
                self._release(pend['bits'])
            return strings, pts

        post_f, xs, devs, lane_evs = [], [], [], {}
        graphs = self.use_graphs and thr_idx is not None and not keep_x_hat and debug_out is None
        if graphs and self.device_coder:
            return self._encode_blocks_device_coder(spans, coords_f, dims, thr_idx)
        for (a, b), cf in zip(spans, coords_f):
            if len(post_f) >= self.pipeline_depth + 4:  # bound the driver's run-ahead (staging memory in flight)
                post_f[len(post_f) - self.pipeline_depth - 4].result()
            if graphs:
                # Two compute streams and two lanes (sets of static buffers / stage graphs): batch i+1's analysis + hyper transforms --
                # small volumes that leave most SMs idle -- run on the front stream beside batch i's synthesis on the main stream.
                # Copies run on their own streams (copy engines): on a compute stream a batch's 7 MB of coordinates and 8 MB of
                # symbols would sit between two stage graphs, 0.4 ms of idle SMs per batch (tools/e2e_gpu_busy.py).  Every hand-over
                # is an event; lane_evs[(lane, n)] = what the next user of that lane's buffers waits for.
                main = torch.cuda.current_stream()
                hs, ds = self._copy_streams()
                front = self._front_stream()
                if not lane_evs:
                    front.wait_stream(main)       # whatever produced the inputs (e.g. the octree partition kernels)
                lane, n = len(post_f) % 2, b - a
                staged = cf.result()
                with torch.cuda.stream(hs):
                    coords = self._h2d_staged(staged)
                    ev_in = torch.cuda.Event()
                    ev_in.record()
                if torch.is_tensor(coords):
                    coords.record_stream(front)
                self.lane = lane
                try:
                    with torch.cuda.stream(front):
                        front.wait_event(ev_in)
                        for ev in lane_evs.get((lane, n), ()):
                            front.wait_event(ev)
                        lat, st = self.device_encode(coords, n, dims, None, block0=getattr(cf, 'block0', None))
                        ev_lat = torch.cuda.Event()
                        ev_lat.record()
                    ds.wait_event(ev_lat)
                    with torch.cuda.stream(ds):
                        pend = {'sym': self._d2h(*self._latent_tensors(lat))}
                    main.wait_event(ev_lat)
                    bits = self.device_synthesis(lat, st, n, dims, threshold_f32(self.thresholds, thr_idx[a:b]))
                finally:
                    self.lane = 0
                ev_bits = torch.cuda.Event()
                ev_bits.record()
                ds.wait_event(ev_bits)
                with torch.cuda.stream(ds):
                    pend['bits'] = self._d2h(bits)
                lane_evs[(lane, n)] = [ev_bits, pend['sym'][1], pend['bits'][1]]
                xs.append(None)
                post_f.append(pool.submit(post, lat, pend))
                continue
            x = ops.densify(self._h2d_staged(cf.result()), b - a, *dims, block0=getattr(cf, 'block0', None))
            pend = {}
            t = self._h2d(threshold_f32(self.thresholds, thr_idx[a:b])) if thr_idx is not None else None
            # the symbol D2H is enqueued BEFORE synthesis is launched: range coding overlaps the synthesis kernels
            dev = self._encode_device(x, lambda d: pend.__setitem__('sym', self._d2h(*self._latent_tensors(d))),
                                      thresholds=t, want_x_hat=keep_x_hat)
            if thr_idx is not None:
                pend['bits'] = self._d2h(dev['bits'])
            xs.append(dev.pop('x_hat'))
            if debug_out is not None:
                devs.append(dict(dev, x_hat=xs[-1]))
            post_f.append(pool.submit(post, dev, pend))
        res = [f.result() for f in post_f]
        strings = [s for r in res for s in r[0]]
        if debug_out is not None:   # per-block dicts of the tensors decompress_octree.py:94-119 compares with the decoder's
            for dev, r in zip(devs, res):
                debug_out.extend(split_debug(self._debug_batch(dev), r[0], self._debug_static()))
        x_hat = None
        if keep_x_hat and xs:
            x_hat = torch.cat(xs) if len(xs) > 1 else xs[0]
        pts = [q for r in res for q in r[1]] if thr_idx is not None else None
        return strings, x_hat, pts

    def compress_blocks(self, sess, blocks, binstr, points, resolution, level, with_normals=False,
                        opt_metrics=('d1_mse',), max_deltas=(np.inf,), fixed_threshold=False, debug=False):
        """src/model_types.py:184-218.  Returns (data_list, metadata, debug_t_list)."""
        loc = self.compress_blocks_local(blocks, resolution, with_normals, opt_metrics, max_deltas, fixed_threshold, debug)
        threshold_list = [tuple(int(v) for v in loc['thr_idx'][:, m]) for m in range(loc['thr_idx'].shape[1])]
        metadata = self._select_best(binstr, loc['x_hat_list'], level, loc['opt_metrics'], points, resolution, with_normals)
        data_list = [list(zip(loc['strings'], threshold_list[x['idx']])) for x in metadata]
        return data_list, metadata, loc['debug']

    def compress_blocks_local(self, blocks, resolution, with_normals=False, opt_metrics=('d1_mse',), max_deltas=(np.inf,),
                              fixed_threshold=False, debug=False):
        """The per-block part of compress_blocks (model_types.py:192-212), without the whole-cloud selection that follows it:
        -> {'strings': per block tuple of bytes, 'thr_idx': (n, n_metrics) threshold indexes, 'opt_metrics': names,
            'x_hat_list': per metric a list of per-block float32 (m,3) points, 'debug': per-block debug dicts or None}.
        A rank of a multi-GPU job runs this on its shard (sharding.compress_blocks_sharded)."""
        assert self.x_shape is not None, 'call compress(x_shape) first'
        n = len(blocks)
        debug_t_list = [] if debug else [None] * n
        if fixed_threshold and not debug:
            opt_metrics_ret = list(opt_metrics)
            thr_idx = np.full((n, len(opt_metrics_ret)), len(self.thresholds) // 2, np.int64)  # model_opt.py:27-31
            strings_list, _, pts = self.encode_blocks(blocks, thr_idx=thr_idx[:, 0], keep_x_hat=False)
            x_hat_list = [pts] * thr_idx.shape[1]  # every opt_metric gets the same fixed threshold
        else:
            # debug=True keeps every block's intermediate tensors (eager kernels, fp32 x_hat materialised): same bytes and points
            strings_list, x_hat, _ = self.encode_blocks(blocks, **({'debug_out': debug_t_list} if debug else {}))
            if fixed_threshold or not n:
                opt_metrics_ret = list(opt_metrics)
                thr_idx = np.full((n, len(opt_metrics_ret)), len(self.thresholds) // 2, np.int64)
            else:
                host_blocks = blocks.to_host() if hasattr(blocks, 'to_host') else blocks
                thr_idx, opt_metrics_ret = self._optimal_thresholds(host_blocks, x_hat, resolution, with_normals, opt_metrics, max_deltas)
            x_hat_list = []
            for m in range(thr_idx.shape[1] if n else 0):
                t = self._h2d(threshold_f32(self.thresholds, thr_idx[:, m]))
                bits, _ = ops.threshold_pack(x_hat, t)
                x_hat_list.append(ops.bits_to_points(bits.cpu().numpy(), tuple(x_hat.shape[2:]), self.coder_threads))
            if not n:
                x_hat_list = [[] for _ in range(thr_idx.shape[1])]
        return {'strings': strings_list, 'thr_idx': thr_idx, 'opt_metrics': opt_metrics_ret, 'x_hat_list': x_hat_list,
                'debug': debug_t_list}

    def _optimal_thresholds(self, blocks, x_hat, resolution, with_normals, opt_metrics, max_deltas):
        """Per-block threshold search (reference src/model_opt.py:21-77).  D1 metrics: every threshold's nearest-neighbour
        sums come from the GPU (csrc/threshold_opt.cu, exact integer distance transforms) and the reference's selection
        rules are applied to them.  D2 metrics (point-to-plane, normals) run the same search on the host with kd-trees."""
        from . import model_opt as MO
        if not any(str(m).startswith('d2') for m in opt_metrics):
            t32 = threshold_f32(self.thresholds, np.arange(len(self.thresholds)))
            idx, ret = [], None
            for a in range(0, len(blocks), self.batch_size):
                chunk = blocks[a:a + self.batch_size]
                coords = blocks_to_coords(chunk, self.coder_threads)
                offsets = np.concatenate([[0], np.cumsum([len(b) for b in chunk])]).astype(np.int64)
                ret, best = MO.compute_optimal_thresholds_batch(chunk, x_hat[a:a + len(chunk)], t32, self._h2d(coords), offsets,
                                                                opt_metrics=opt_metrics, max_deltas=max_deltas)
                idx.append(best)
            return np.concatenate(idx), list(ret)
        # D2 (point-to-plane) metrics: host kd-tree search with the reference's normal transfer, block by block
        xh = torch.clamp(x_hat[:, 0], 0.0, 1.0).cpu().numpy()
        idx, ret = [], None
        for j, block in enumerate(blocks):
            block = np.asarray(block)
            normals = block[:, block.shape[1] - 3:] if with_normals else None
            ret, best = MO.compute_optimal_thresholds(block, xh[j], self.thresholds, resolution, normals=normals,
                                                      opt_metrics=opt_metrics, max_deltas=max_deltas, fixed_threshold=False)
            idx.append(best)
        return np.asarray(idx, np.int64), list(ret)

    def _select_best(self, binstr, x_hat_list, level, opt_metrics, points, resolution, with_normals):
        """select_best_per_opt_metric (model_types.py:128-176) when the caller passes the whole cloud and its octree;
        without them (block-level callers, benchmarks) the first opt_metric is selected and no metrics are reported."""
        if points is None or binstr is None:
            return [{'idx': 0, 'metrics': {}, 'x_hat_list': x_hat_list[0], 'blocks_depart': None, 'blocks_full': None}]
        return select_best_per_opt_metric(binstr, x_hat_list, level, opt_metrics, points, resolution, with_normals)

    def decompress_blocks(self, sess, blocks, x_shape, debug=False):
        """src/model_types.py:220-238: blocks = [(strings, threshold_idx)] -> ([float32 (m,3)], debug list)."""
        self._sync_trainer()
        dims = tuple(int(s) for s in x_shape)[-3:]
        chunks = self._chunks(list(blocks))
        pool = self._pool()
        strings = [[c[0] for c in chunk] for chunk in chunks]
        # stage 0 (host): first latent's strings -> symbols; stage 1 (GPU): what the second latent's decoding needs
        # (hyper-synthesis -> scale indexes); stage 2 (host): second latent; stage 3 (GPU): synthesis + threshold + pack;
        # stage 4 (host): packed bits -> points
        # The driver walks the batches with a lag: stage 1 of batch i is enqueued before it waits for stage 2 of batch
        # i-lag, and the first-latent decodes are submitted only two batches ahead, so that the host work the GPU is waiting
        # for (batch 0's second latent) is at the front of the workers' queue instead of behind every batch's stage 0.
        graphs = self.use_graphs and not debug
        if graphs and self.device_coder:
            return self._decompress_blocks_device_coder(chunks, dims)
        nb, lag, ahead = len(chunks), 2, 2
        f0 = {i: pool.submit(self._decode_host0, strings[i], dims) for i in range(min(ahead, nb))}
        ctxs, f2, f4, dbgs = {}, {}, [], []
        # graphs: the batches alternate between two lanes (two sets of static buffers / stage graphs), so that a batch's symbols can
        # be copied in (and its scale indexes / occupancy bits copied out) on the copy streams while the other lane's graph runs
        self._pipe_evs = {}
        for i in range(nb + lag):
            if i < nb:
                ctxs[i] = self._graph_dev1(f0.pop(i).result(), len(strings[i]), dims, lane=i % 2) if graphs else self._decode_dev1(f0.pop(i).result(), dims)
                f2[i] = pool.submit(self._decode_host1, ctxs[i], strings[i])
                if i + ahead < nb:
                    f0[i + ahead] = pool.submit(self._decode_host0, strings[i + ahead], dims)
            j = i - lag
            if j >= 0:
                f2.pop(j).result()
                chunk, ctx = chunks[j], ctxs.pop(j)
                idx = np.asarray([int(c[1]) for c in chunk], np.int64)
                if graphs:
                    dbg = {'bits': self._graph_dev2(ctx, len(chunk), dims, threshold_f32(self.thresholds, idx), lane=j % 2)}
                    pend = self._d2h_side(dbg['bits'], ('dec2', j % 2, len(chunk)))
                else:
                    x_hat, dbg = self._decode_dev2(ctx, self._h2d(threshold_f32(self.thresholds, idx)), debug)
                    pend = self._d2h(dbg['bits'])
                f4.append(pool.submit(self._points_task, pend, dims))
                dbgs.append(split_debug(self._debug_batch(dbg), [c[0] for c in chunk], self._debug_static()) if debug
                            else [None] * len(chunk))
        pts = [f.result() for f in f4]
        return [p for r in pts for p in r], [d for r in dbgs for d in r]

    # -- block loops with the entropy coder on the GPU ----------------------------------------------------
    # A stream is serial, so the device coder's parallelism is the number of streams: the loops below collect the symbols
    # of a whole group of batches (<= coder_group_blocks blocks) in HBM and code all of the group's streams with one launch
    # per latent -- a few milliseconds per group, however many blocks it holds -- instead of handing every batch's symbols
    # to the host workers.  _coder_latents() (V1/V2) describes the latents in the order of the strings tuple.
    def _coder_latents(self, dims):
        """[{'sym': latent-dict key, 'idx': key or None, 'shape': per-block shape, 'tables': host tables, 'st': static key}]"""
        raise NotImplementedError

    def _groups(self, spans):
        per = max(1, self.coder_group_blocks // self.batch_size)
        return [spans[i:i + per] for i in range(0, len(spans), per)]

    def _copy_streams(self, who='enc'):
        """(host-to-device, device-to-host) side streams of this model on the current device; the compute stream itself when
        PCCGEO_SIDE_COPIES=0 (every copy then sits between the stage graphs again: for A/B measurements and bisecting)"""
        if os.environ.get('PCCGEO_SIDE_COPIES', '1') in ('0', 'dec' if who == 'enc' else 'enc'):
            return torch.cuda.current_stream(), torch.cuda.current_stream()
        key = ('copy_streams', torch.cuda.current_device())
        if key not in self.__dict__:
            self.__dict__[key] = (torch.cuda.Stream(), torch.cuda.Stream())
        return self.__dict__[key]

    def _front_stream(self, who='enc'):
        """second compute stream of the block loops (the main stream itself when PCCGEO_SIDE_COPIES switches the side streams off)"""
        if os.environ.get('PCCGEO_SIDE_COPIES', '1') in ('0', 'dec' if who == 'enc' else 'enc'):
            return torch.cuda.current_stream()
        key = ('front_stream', torch.cuda.current_device())
        if key not in self.__dict__:
            self.__dict__[key] = torch.cuda.Stream()
        return self.__dict__[key]

    def _coder_stream(self):
        """The stream of the device coder: the current one, or with coder_overlap a side stream (one per model and device)."""
        if not self.coder_overlap:
            return torch.cuda.current_stream()
        key = ('coder_stream', torch.cuda.current_device())
        if key not in self.__dict__:
            self.__dict__[key] = torch.cuda.Stream()
        return self.__dict__[key]

    def _encode_blocks_device_coder(self, spans, coords_f, dims, thr_idx):
        pool = self._pool()
        lats = self._coder_latents(dims)
        cf_of = dict(zip(spans, coords_f))
        main, side = torch.cuda.current_stream(), self._coder_stream()
        front = self._front_stream()
        front.wait_stream(main)
        hs, ds = self._copy_streams()
        str_f, pts_f, lane_evs, nbatch = [], [], {}, 0
        for group in self._groups(spans):
            g0, g1 = group[0][0], group[-1][1]
            big = {}
            for l in lats:
                for k in (l['sym'], l['idx']):
                    if k is not None and k not in big:
                        big[k] = torch.empty((g1 - g0,) + tuple(l['shape']), dtype=torch.int32, device='cuda')
                        big[k].record_stream(side)
                        big[k].record_stream(front)
            for a, b in group:
                cf = cf_of[(a, b)]
                # as in encode_blocks: coordinates in / bits out on the copy streams, analysis + hyper transforms of a batch on the
                # front stream beside the previous batch's synthesis, two lanes of static buffers / stage graphs
                lane, n = nbatch % 2, b - a
                nbatch += 1
                staged = cf.result()
                with torch.cuda.stream(hs):
                    coords = self._h2d_staged(staged)
                    ev_in = torch.cuda.Event()
                    ev_in.record()
                if torch.is_tensor(coords):
                    coords.record_stream(front)
                self.lane = lane
                try:
                    with torch.cuda.stream(front):
                        front.wait_event(ev_in)
                        for ev in lane_evs.get((lane, n), ()):
                            front.wait_event(ev)
                        lat, st = self.device_encode(coords, n, dims, None, block0=getattr(cf, 'block0', None))
                        for k, t in big.items():
                            t[a - g0:b - g0].copy_(lat[k].view(t[a - g0:b - g0].shape))
                        ev_lat = torch.cuda.Event()
                        ev_lat.record()
                    main.wait_event(ev_lat)
                    bits = self.device_synthesis(lat, st, n, dims, threshold_f32(self.thresholds, thr_idx[a:b]))
                finally:
                    self.lane = 0
                ev_bits = torch.cuda.Event()
                ev_bits.record()
                ds.wait_event(ev_bits)
                with torch.cuda.stream(ds):
                    pend = self._d2h(bits)
                lane_evs[(lane, n)] = [ev_bits, pend[1]]
                pts_f.append(pool.submit(self._points_task, pend, dims))
            # the group's streams are coded on the side stream, under the next group's transforms
            side.wait_stream(main)
            coded = []
            with torch.cuda.stream(side):
                for l in lats:
                    per = int(np.prod(l['shape']))
                    idx = big[l['idx']] if l['idx'] is not None else None
                    packed, lengths, offsets, err = ops.range_encode_device(big[l['sym']], ops.device_tables(l['tables']), indexes=idx,
                                                                            channel_stride=per // l['shape'][0])
                    coded.append((packed, self._d2h(lengths, offsets, err)))
            str_f.append(pool.submit(self._strings_task, coded, big, lats, g1 - g0))
        strings = [t for f in str_f for t in f.result()]
        pts = [q for f in pts_f for q in f.result()]
        return strings, None, pts

    def _strings_task(self, coded, big, lats, n):
        """Host worker: byte ranges of a coded group -> per-block tuples of bytes.  The (rare) stream that outgrew the device
        buffer sends its latent through the host coder (same bytes)."""
        per_latent = []
        for (packed, pend), l in zip(coded, lats):
            side = self._worker_stream(packed.device)
            lengths, offsets, err = self._wait(pend)
            lengths, offsets, bad = lengths.copy(), offsets.copy(), int(err[0])
            self._release(pend)
            if bad & 1:
                raise PccGeoError('range_encode_device: table index out of range')
            total = int(offsets[-1])
            # bad & 2: an escape value beyond the device coder's 32-bit word (|value| ~ 2^31) -> host coder, same bytes
            if (bad & 2) or (lengths < 0).any() or total > packed.numel():
                sym = big[l['sym']].cpu().numpy()
                offs = np.arange(n + 1, dtype=np.int64) * int(np.prod(l['shape']))
                if l['idx'] is not None:
                    per_latent.append(ops.range_encode(sym.reshape(-1), offs, l['tables'], indexes=big[l['idx']].cpu().numpy().reshape(-1),
                                                       threads=self.coder_threads))
                else:
                    per_latent.append(ops.range_encode(sym.reshape(-1), offs, l['tables'],
                                                       channel_stride=int(np.prod(l['shape'][1:])), threads=self.coder_threads))
                continue
            buf = _pinned.get(max(total, 1))
            with torch.cuda.device(packed.device), torch.cuda.stream(side):
                # the exact byte count is only known now: second, tight copy from this worker, on a stream of the tensor's
                # own device (a stream of another device would leave the copy on that device's default stream)
                buf[:total].copy_(packed[:total], non_blocking=True)
                done = torch.cuda.Event()
                done.record(side)
            done.synchronize()
            view = memoryview(buf.numpy())
            offs = offsets.tolist()
            per_latent.append([bytes(view[offs[i]:offs[i + 1]]) for i in range(n)])
            _pinned.put(buf)
        return list(zip(*per_latent))

    def _worker_stream(self, device):
        """A side stream of this worker thread ON `device` (one per thread and device)."""
        import threading
        tl = self.__dict__.setdefault('_tls', threading.local())
        streams = tl.__dict__.setdefault('streams', {})
        key = torch.device(device).index
        if key not in streams:
            with torch.cuda.device(device):
                streams[key] = torch.cuda.Stream(device=device)
        return streams[key]

    def _upload_strings(self, strings):
        """[bytes] -> (uint8 CUDA blob, int64 CUDA offsets (n+1)); the strings are gathered straight into pinned memory."""
        offs = np.zeros(len(strings) + 1, np.int64)
        offs[1:] = np.cumsum([len(t) for t in strings])
        total = int(offs[-1])
        buf = _pinned.get(total + 1)   # + one pad byte: never an empty device buffer
        view = buf[:total + 1].numpy()
        for t, o in zip(strings, offs):
            if t:
                view[o:o + len(t)] = np.frombuffer(t, np.uint8)
        view[total] = 0
        blob = torch.empty(total + 1, dtype=torch.uint8, device='cuda')
        blob.copy_(buf[:total + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _pinned.put_after(buf, ev)
        return blob, self._h2d(offs)

    def _decompress_blocks_device_coder(self, chunks, dims):
        pool = self._pool()
        lats = self._coder_latents(dims)
        main, side = torch.cuda.current_stream(), self._coder_stream()
        spans, a = [], 0
        for c in chunks:
            spans.append((a, a + len(c)))
            a += len(c)
        flat = [c for chunk in chunks for c in chunk]
        f4, errs = [], []

        def decode_group(group):
            """Side stream: strings -> symbols of every latent, in decoding order (the last latent of the strings tuple -- the
            hyperprior, whose tables are fixed -- first, then what its values parameterise); main stream, in between: the
            group's hyper-synthesis batches.  -> (symbols of the first latent, event after which they are complete)"""
            g0, g1 = group[0][0], group[-1][1]
            n = g1 - g0
            sym, idx_big = {}, None
            for li in reversed(range(len(lats))):
                l = lats[li]
                per = int(np.prod(l['shape']))
                with torch.cuda.stream(side):
                    blob, offs = self._upload_strings([flat[g0 + i][0][li] for i in range(n)])
                if l['idx'] is not None:   # decoded hyper-latent -> scale indexes of the whole group, batch by batch
                    main.wait_stream(side)
                    idx_big = torch.empty((n,) + tuple(l['shape']), dtype=torch.int32, device='cuda')
                    idx_big.record_stream(side)
                    for a, b in group:
                        st = self._static(b - a, dims)
                        st['sym0'].copy_(sym[lats[li + 1]['sym']][a - g0:b - g0].view(st['sym0'].shape))
                        ctx = self._stage('dec1', b - a, dims, lambda: self._dec1_compute(st['sym0']))
                        idx_big[a - g0:b - g0].copy_(ctx['indexes'].view(idx_big[a - g0:b - g0].shape))
                    side.wait_stream(main)
                with torch.cuda.stream(side):
                    err = torch.zeros(1, dtype=torch.int32, device='cuda')
                    errs.append(err)
                    out, _ = ops.range_decode_device(blob, offs, n, per, ops.device_tables(l['tables']),
                                                     indexes=idx_big if l['idx'] is not None else None,
                                                     channel_stride=per // l['shape'][0], err=err)
                    out.record_stream(main)
                    sym[l['sym']] = out
            ev = torch.cuda.Event()
            ev.record(side)
            return sym[lats[0]['sym']], ev

        bits_ev = [None]
        _, ds = self._copy_streams('dec')

        def synthesize_group(group, ysym, ev):
            g0 = group[0][0]
            main.wait_event(ev)
            for a, b in group:
                st = self._static(b - a, dims)
                st['sym1'].copy_(ysym[a - g0:b - g0].view(st['sym1'].shape))
                thr = threshold_f32(self.thresholds, np.asarray([int(c[1]) for c in flat[a:b]], np.int64))
                self._copy_in(st['thr'], thr)
                if bits_ev[0] is not None:
                    main.wait_event(bits_ev[0])   # the previous batch's bits (static output of this graph) are copied on the D2H stream
                bits = self._stage('dec2', b - a, dims, lambda: self._dec2_compute(None, st))
                done = torch.cuda.Event()
                done.record()
                ds.wait_event(done)
                with torch.cuda.stream(ds):
                    pend = self._d2h(bits)
                bits_ev[0] = pend[1]
                f4.append(pool.submit(self._points_task, pend, dims))

        # software pipeline with a lag of one group: the side stream decodes group g+1 under the synthesis of group g
        prev = None
        for group in self._groups(spans):
            cur = (group,) + decode_group(group)
            if prev is not None:
                synthesize_group(*prev)
            prev = cur
        if prev is not None:
            synthesize_group(*prev)
        main.wait_stream(side)
        bad = torch.stack(errs).sum().item() if errs else 0   # one sync, after everything is enqueued
        pts = [f.result() for f in f4]
        if bad:
            raise PccGeoError('range_decode_device: corrupt stream')
        return [p for r in pts for p in r], [None] * len(flat)

    def _debug_batch(self, dev):
        """The reference's debug_tensors of one batch (model_types.py:295,309): {key: CUDA tensor with a batch axis}."""
        return {'y_hat': dev['y_hat'], 'x_hat': dev['x_hat']}

    def _debug_static(self):
        return None

    def _points_task(self, pend, dims):
        pts = ops.bits_to_points(self._wait(pend)[0], dims, self.coder_threads)
        self._release(pend)
        return pts

    # Stage graphs of the decode pipeline with their copies on the copy streams.  self._pipe_evs[(stage, lane, batch size)] -- the key of
    # a set of static buffers -- holds what the next user of those buffers has to wait for: 'graph' = the last replay (it reads the static inputs), 'out' = the
    # device-to-host copy of its static outputs.
    def _lane_in(self, key, fill, st):
        """Run `fill()` (host-to-device copies into this lane's static inputs) on the H2D stream, after the lane's previous graph;
        make the compute stream wait for it and for the lane's previous output copy."""
        main = torch.cuda.current_stream()
        hs, _ = self._copy_streams('dec')
        prev = self._pipe_evs.get(key, {})
        with torch.cuda.stream(hs):
            hs.wait_event(st['ready'])     # first use of the lane: its static buffers were just allocated and zero-filled
            if 'graph' in prev:
                hs.wait_event(prev['graph'])
            fill()
            ev = torch.cuda.Event()
            ev.record()
        main.wait_event(ev)
        if 'out' in prev:
            main.wait_event(prev['out'])

    def _lane_done(self, key):
        ev = torch.cuda.Event()
        ev.record()
        self._pipe_evs.setdefault(key, {})['graph'] = ev

    def _d2h_side(self, tensor, key):
        """_d2h of a stage graph's static output on the D2H stream, after the graph recorded by _lane_done(key)."""
        _, ds = self._copy_streams('dec')
        ds.wait_event(self._pipe_evs[key]['graph'])
        with torch.cuda.stream(ds):
            pend = self._d2h(tensor)
        self._pipe_evs[key]['out'] = pend[1]
        return pend

    def _graph_dev1(self, sym0_host, n, dims, lane=0):
        """decode stage 1 as a graph: first latent's symbols (host) -> static buffer -> _dec1_compute."""
        self.lane = lane
        try:
            # on the front stream: the hyper-synthesis of a batch (2^3 .. 8^3 volumes) runs beside another batch's synthesis
            with torch.cuda.stream(self._front_stream('dec')):
                st = self._static(n, dims)
                self._lane_in(('dec1', lane, n), lambda: self._copy_in(st['sym0'], sym0_host), st)
                ctx = dict(self._stage('dec1', n, dims, lambda: self._dec1_compute(st['sym0'])))
                self._lane_done(('dec1', lane, n))
            if 'indexes' in ctx:
                ctx['idx_pending'] = self._d2h_side(ctx['indexes'], ('dec1', lane, n))
            return ctx
        finally:
            self.lane = 0

    def _graph_dev2(self, ctx, n, dims, thr, lane=0):
        """decode stage 3 as a graph: [second latent's symbols ->] synthesis + threshold + pack -> bits."""
        self.lane = lane
        try:
            st = self._static(n, dims)

            def fill():
                if 'ysym_staged' in ctx:
                    host, buf = ctx.pop('ysym_staged')
                    st['sym1'].copy_(host.view(st['sym1'].shape), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    _pinned.put_after(buf, ev)
                elif 'ysym' in ctx:
                    st['sym1'].copy_(ctx['ysym']) if torch.is_tensor(ctx['ysym']) else self._copy_in(st['sym1'], ctx['ysym'])
                st['thr'].copy_(thr) if torch.is_tensor(thr) else self._copy_in(st['thr'], thr)
            self._lane_in(('dec2', lane, n), fill, st)
            bits = self._stage('dec2', n, dims, lambda: self._dec2_compute(ctx, st))
            self._lane_done(('dec2', lane, n))
            return bits
        finally:
            self.lane = 0

    def _decode_batch(self, strings_list, dims, thresholds=None, want_x_hat=True):
        """One batch through the four decode stages, sequentially."""
        ctx = self._decode_dev1(self._decode_host0(strings_list, dims), dims)
        self._decode_host1(ctx, strings_list)
        return self._decode_dev2(ctx, thresholds, want_x_hat)

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
        # tf.summary.merge_all() of the reference's graph (model_types.py:271-274,357-362): scalars here, the bc/* scores
        # of the quantised reconstruction included; histograms are the tensors kept on the model (y, y_tilde, ...)
        self.merged_summary = {'loss': self.train_loss, 'fl': self.train_fl, 'mbpov/total': self.train_mbpov,
                               'num_occupied_voxels': n_occ,
                               **binary_classification_summaries(quantize_tensor(x), quantize_tensor(x_tilde))}
        self.step = getattr(self, 'step', 0)
        # sess.run(m.train_op) of the reference == m.train_op(x): forward + backward + both Adam steps + table refresh
        from .training import Trainer
        if getattr(self, 'trainer', None) is None:
            self.trainer = Trainer(self, gamma, alpha, lmbda, tensor_cores=getattr(self, 'train_tensor_cores', False))
        self.trainer.gamma, self.trainer.alpha, self.trainer.lmbda = gamma, alpha, lmbda
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
        self._sync_trainer()
        y = self.analysis_transform(x)
        y_tilde, _ = self.entropy_bottleneck(y, training=True, noise=noise_y)
        x_tilde = self.synthesis_transform(y_tilde)
        self._finish_train(x, x_tilde, [self.entropy_bottleneck.log_likelihood_sum(y_tilde)], gamma, alpha, lmbda)
        self.y, self.y_tilde = y, y_tilde

    def compress(self, x_shape):  # model_types.py:283-295
        self.x_shape = tuple(int(s) for s in x_shape)
        self.x, self.x_hat = GraphHandle('x', self.x_shape), GraphHandle('x_hat', self.x_shape)
        self.strings = (GraphHandle('y_string', (self.x_shape[0],), 'string'),)
        self.debug_tensors = {k: GraphHandle(k) for k in ('y_hat', 'x_hat')}

    def decompress(self):  # model_types.py:297-309
        self.strings_t = [GraphHandle('y_string', None, 'string')]
        self.x_shape_t = GraphHandle('x_shape', (3,), 'int32')
        self.x_hat = GraphHandle('x_hat')
        self.debug_tensors = {k: GraphHandle(k) for k in ('y_hat', 'x_hat')}

    def _latents(self, x):
        y = self.analysis_transform(x)
        y_sym, y_hat = self.entropy_bottleneck.quantize(y)
        return {'y_sym': y_sym, 'y_hat': y_hat}

    def _static_latents(self, n, dims):
        return {'sym1': torch.zeros((n, self.num_filters) + tuple(d // 8 for d in dims), device='cuda', dtype=torch.int32)}

    # one latent: nothing on the GPU between the two host stages, so the decode graph is a single stage (dequantise +
    # synthesis + threshold + pack) fed from the static symbol buffer right before its replay -- a 'dec1' graph would
    # leave its output in a static tensor that the next batch's replay overwrites before this batch's synthesis runs
    def _graph_dev1(self, sym0_host, n, dims, lane=0):
        return {'ysym': sym0_host}

    def _dec2_compute(self, ctx, st):
        y_hat = ops.eb_dequantize(st['sym1'], self.entropy_bottleneck.device_params())
        return self._synthesize(y_hat, st['thr'], False)[1]

    def _encode_device(self, x, after_latents=None, thresholds=None, want_x_hat=True):
        dev = self._latents(x)
        y_hat = dev['y_hat']
        if after_latents is not None:
            after_latents(dev)
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.last_x, self.last_x_hat = x, x_hat          # the most recent batch's tensors (CUDA), for interactive use
        dev['x_hat'], dev['bits'] = x_hat, bits
        return dev

    @staticmethod
    def _latent_tensors(dev):
        return (dev['y_sym'],)

    def _coder_latents(self, dims):
        return [{'sym': 'y_sym', 'idx': None, 'shape': (self.num_filters,) + tuple(d // 8 for d in dims),
                 'tables': self.entropy_bottleneck.tables}]

    def _encode_host(self, dev, host=None):
        y_sym = host[0] if host is not None else dev['y_sym'].cpu().numpy()
        ys = self.entropy_bottleneck.encode_symbols(y_sym, self.coder_threads)
        return [(s,) for s in ys]

    def _decode_host0(self, strings_list, dims):
        shp = (self.num_filters,) + tuple(d // 8 for d in dims)  # model_types.py:305
        return self.entropy_bottleneck.decode_symbols([s[0] for s in strings_list], shp, self.coder_threads)

    def _decode_dev1(self, sym, dims):
        return {'y_hat': ops.eb_dequantize(self._h2d(sym), self.entropy_bottleneck.device_params())}

    def _decode_host1(self, ctx, strings_list):
        return None

    def _decode_dev2(self, ctx, thresholds=None, want_x_hat=True):
        y_hat = ctx['y_hat']
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.last_x_hat = x_hat
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
        self._sync_trainer()
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
        self.x, self.x_hat = GraphHandle('x', self.x_shape), GraphHandle('x_hat', self.x_shape)
        self.strings = tuple(GraphHandle(k, (self.x_shape[0],), 'string') for k in ('y_string', 'z_string'))
        self.debug_tensors = {k: GraphHandle(k) for k in ('z_hat', 'sigma_hat', 'decompress/indexes', 'decompress/symbols', 'y_hat', 'x_hat')}

    def decompress(self):  # model_types.py:393-411
        self.strings_t = [GraphHandle(k, None, 'string') for k in ('y_string', 'z_string')]
        self.x_shape_t = GraphHandle('x_shape', (3,), 'int32')
        self.x_hat = GraphHandle('x_hat')
        self.debug_tensors = {k: GraphHandle(k) for k in ('z_hat', 'sigma_hat', 'decompress/indexes', 'decompress/symbols', 'y_hat', 'x_hat')}

    def _latents(self, x):
        y = self.analysis_transform(x)
        z = self.hyper_analysis_transform(y)
        z_sym, z_hat = self.entropy_bottleneck.quantize(z)
        sigma_hat = self.hyper_synthesis_transform(z_hat)
        cb = GaussianConditional(sigma_hat, self.scale_table, data_format=self.data_format)
        y_sym, y_hat, idx = cb.quantize(y)
        return {'y': y, 'z': z, 'z_sym': z_sym, 'z_hat': z_hat, 'sigma_hat': sigma_hat, 'y_sym': y_sym, 'y_hat': y_hat,
                'indexes': idx, 'cb': cb}

    def _static_latents(self, n, dims):
        f = self.num_filters
        return {'sym0': torch.zeros((n, f) + tuple(d // 16 for d in dims), device='cuda', dtype=torch.int32),
                'sym1': torch.zeros((n, f) + tuple(d // 8 for d in dims), device='cuda', dtype=torch.int32)}

    def _dec1_compute(self, zsym):
        z_hat = ops.eb_dequantize(zsym, self.entropy_bottleneck.device_params())
        sigma_hat = self.hyper_synthesis_transform(z_hat)
        cb = GaussianConditional(sigma_hat, self.scale_table, data_format=self.data_format)
        return {'z_hat': z_hat, 'sigma_hat': sigma_hat, 'cb': cb, 'indexes': cb.indexes()}

    def _dec2_compute(self, ctx, st):
        return self._synthesize(ops.i32_to_f32(st['sym1']), st['thr'], False)[1]

    def _encode_device(self, x, after_latents=None, thresholds=None, want_x_hat=True):
        dev = self._latents(x)
        z_hat, sigma_hat, idx, y_hat = dev['z_hat'], dev['sigma_hat'], dev['indexes'], dev['y_hat']
        if after_latents is not None:
            after_latents(dev)
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.last_x, self.last_x_hat = x, x_hat
        dev['x_hat'], dev['bits'] = x_hat, bits
        return dev

    def _debug_batch(self, dev):
        """model_types.py:390-391,410-411: z_hat, sigma_hat, the Gaussian conditional's dbg_dec, y_hat, x_hat."""
        return {'z_hat': dev['z_hat'], 'sigma_hat': dev['sigma_hat'], 'decompress/build/_scale': dev['sigma_hat'],
                'decompress/build/_indexes': dev['indexes'], 'decompress/indexes': dev['indexes'],
                'decompress/symbols': dev['y_hat'].to(torch.int32), 'decompress/outputs': dev['y_hat'],
                'y_hat': dev['y_hat'], 'x_hat': dev['x_hat']}

    def _debug_static(self):
        t = gaussian_tables(self.scale_table)
        return {'decompress/build/scale_table': self.scale_table, 'decompress/build/_quantized_cdf': t['cdf'],
                'decompress/quantized_cdf': t['cdf'], 'decompress/build/_cdf_length': t['cdf_length'],
                'decompress/build/_offset': t['offset'], 'decompress/build/fill': np.int32(len(self.scale_table) - 1)}

    @staticmethod
    def _latent_tensors(dev):
        return (dev['z_sym'], dev['y_sym'], dev['indexes'])

    def _coder_latents(self, dims):
        f = self.num_filters
        return [{'sym': 'y_sym', 'idx': 'indexes', 'shape': (f,) + tuple(d // 8 for d in dims),
                 'tables': gaussian_tables(self.scale_table)},
                {'sym': 'z_sym', 'idx': None, 'shape': (f,) + tuple(d // 16 for d in dims), 'tables': self.entropy_bottleneck.tables}]

    def _encode_host(self, dev, host=None):
        z_sym, y_sym, idx = host if host is not None else [t.cpu().numpy() for t in self._latent_tensors(dev)]
        zs = self.entropy_bottleneck.encode_symbols(z_sym, self.coder_threads)
        ys = dev['cb'].encode_symbols(y_sym, idx, self.coder_threads)
        return list(zip(ys, zs))  # (y_string, z_string): model_types.py:389

    def _decode_host0(self, strings_list, dims):
        zshp = (self.num_filters,) + tuple(d // 16 for d in dims)  # model_types.py:403
        return self.entropy_bottleneck.decode_symbols([s[1] for s in strings_list], zshp, self.coder_threads)

    def _decode_dev1(self, zsym, dims):
        ctx = self._dec1_compute(self._h2d(zsym))
        ctx['idx_pending'] = self._d2h(ctx['indexes'])
        return ctx

    def _decode_host1(self, ctx, strings_list):
        pend = ctx.pop('idx_pending')
        ctx['ysym'] = ctx['cb'].decode_symbols([s[0] for s in strings_list], self._wait(pend)[0], self.coder_threads)
        self._release(pend)
        # stage the symbols into pinned memory HERE, on the worker: the 4 MB copy (0.6 ms) otherwise sits on the driver thread
        # between two graph launches of every batch
        ctx['ysym_staged'] = self._stage_host(ctx['ysym'])

    def _decode_dev2(self, ctx, thresholds=None, want_x_hat=True):
        ysym = self._h2d_staged(ctx.pop('ysym_staged')) if 'ysym_staged' in ctx else self._h2d(ctx['ysym'])
        y_hat = ops.i32_to_f32(ysym)
        x_hat, bits = self._synthesize(y_hat, thresholds, want_x_hat)
        self.last_x_hat = x_hat
        return x_hat, {'z_hat': ctx['z_hat'], 'sigma_hat': ctx['sigma_hat'], 'indexes': ctx['indexes'], 'y_hat': y_hat,
                       'x_hat': x_hat, 'bits': bits}


class ModelType(Enum):
    v1 = CompressionModelV1
    v2 = CompressionModelV2
