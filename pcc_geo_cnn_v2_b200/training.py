"""The tr_train.py hot loop (reference src/model_types.py:250-281 / 327-369, src/tr_train.py:91-134) on the GPU:

    trainer = Trainer(model, gamma, alpha, lmbda)      # == model.train(x, gamma, alpha, lmbda) building the graph
    out = trainer.step(x)                              # == sess.run(train_op): forward, backward, both Adam steps,
                                                       #    entropy-bottleneck table refresh

Forward and backward run as libpccgeo kernels with saved activations (no autograd).  The convolutions of the forward pass
and the data gradients (the adjoint layer: conv <-> transposed conv with the same kernel array) go through the same
kernel dispatch as the codec -- the tcgen05 kernels in the active precision mode (bf16x3 by default, fp32-class) -- with
the packed weight images rebuilt from the fp32 master weights every step; the weight gradients of the 3x3x3 layers with 16 / 32 / 64
channels run on the tcgen05 weight-gradient kernel (csrc/conv3d_wgrad_umma.cu; stride-2 layers by phase decomposition), bias
gradients, ReLU masks, focal-loss and likelihood backward are the fp32 kernels of csrc/train.cu.  That is
`Trainer(..., tensor_cores=True)` (`model.train_tensor_cores = True`): 155 -> 29 ms per batch-32 c3p step, conv-kernel gradients
within 1e-2 of float64 autograd (the focal loss amplifies the 1e-5 forward differences).  The default keeps every conv on the fp32
CUDA-core kernel (gradients within 2e-3).  Both optimisers follow TF1's AdamOptimizer
(lr_t = lr*sqrt(1-b2^t)/(1-b1^t), theta -= lr_t*m/(sqrt(v)+eps)): Adam(1e-4) on every trainable of the main loss, Adam(1e-3)
on the entropy bottleneck's quantiles (auxiliary loss).  Whole-batch sums everywhere (FL is sum-reduced and mbpov divides
by the batch's occupied-voxel count), so data-parallel training all-reduces SUMS, not averages: with torch.distributed
initialised (one process per GPU, NCCL) every rank runs its slice of the global batch, the occupied-voxel count is
all-reduced BEFORE the backward pass (it scales every likelihood gradient), and all gradients travel in ONE flat
all-reduce; every rank then applies the same Adam step, so the replicas stay bit-identical without a broadcast.
"""
import math

import numpy as np
import torch

from . import ops
from .entropy_models import GaussianConditional
from . import model_transforms as MT
from .model_transforms import trace


def _softplus(a):
    return np.logaddexp(0.0, a)


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


class _HostAdam:
    """TF1-form Adam on small host arrays (entropy-bottleneck variables)."""

    def __init__(self, arrays, lr):
        self.lr, self.t = lr, 0
        self.m = [np.zeros_like(a, np.float64) for a in arrays]
        self.v = [np.zeros_like(a, np.float64) for a in arrays]

    def step(self, arrays, grads, b1=0.9, b2=0.999, eps=1e-8):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        for a, g, m, v in zip(arrays, grads, self.m, self.v):
            g = np.asarray(g, np.float64)
            m *= b1
            m += (1 - b1) * g
            v *= b2
            v += (1 - b2) * g * g
            a -= (lr_t * m / (np.sqrt(v) + eps)).astype(a.dtype)


class Trainer:
    def __init__(self, model, gamma=2, alpha=0.9, lmbda=1e-4, lr=1e-4, aux_lr=1e-3, tensor_cores=False):
        self.model, self.gamma, self.alpha, self.lmbda, self.lr = model, gamma, alpha, lmbda, lr
        self.tensor_cores = tensor_cores
        self.twins = {}    # layer -> adjoint layer (data gradient), sharing the kernel array
        self.distributed = None   # None: use torch.distributed when it is initialised with more than one rank
        self.v2 = hasattr(model, 'hyper_analysis_transform')
        self.transforms = model.transforms()
        self.traces = {k: trace(t, fuse_residual=False) for k, t in self.transforms.items()}
        self.params = {}   # layer -> dict(w, b, mw, vw, mb, vb) device fp32, tap-major weights
        self.t = 0
        self.dirty = False  # the device master weights are ahead of the layer objects (set by step, cleared by sync_to_model)
        eb = model.entropy_bottleneck
        self._eb_arrays = lambda: eb.matrices + eb.biases + eb.factors
        self.eb_adam = None
        self.aux_adam = None
        self.aux_lr = aux_lr
        self.marks = []     # (label, host seconds) of the last step: where the host thread was when (tools/train_profile.py)

    # -- parameters ------------------------------------------------------------------------------------
    def _p(self, layer, in_channels):
        if layer not in self.params:
            layer.build(in_channels)
            w = layer.dev('w_tap').clone()
            b = layer.dev('bias')
            p = {'w': w, 'b': None if b is None else b.clone(), 'mw': torch.zeros_like(w), 'vw': torch.zeros_like(w)}
            if b is not None:
                p['mb'], p['vb'] = torch.zeros_like(b), torch.zeros_like(b)
            self.params[layer] = p
        return self.params[layer]

    def _params_to_host(self):
        """{layer: (w numpy, b numpy | None)} with ONE device-to-host copy (a .cpu() per tensor is a stream sync per tensor)."""
        parts = []
        for p in self.params.values():
            parts.append(p['w'].reshape(-1))
            if p['b'] is not None:
                parts.append(p['b'].reshape(-1))
        dev = torch.cat(parts)
        pin = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)     # pageable D2H runs at ~3 GB/s: 2.3 ms for c3p's 6.4 MB
        pin.copy_(dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        flat = pin.numpy()
        out, pos = {}, 0
        for layer, p in self.params.items():
            nw = p['w'].numel()
            w = flat[pos:pos + nw].reshape(tuple(p['w'].shape))
            pos += nw
            b = None
            if p['b'] is not None:
                b = flat[pos:pos + p['b'].numel()].copy()
                pos += p['b'].numel()
            out[layer] = (w, b)
        return out

    def sync_to_model(self):
        """Copy the trained device parameters back into the layers (Keras layouts) so that the model's own forward
        (validation `train()`), `get_weights()` and the codec path see them.  The model calls this lazily whenever it is used
        after a step (CompressionModel._sync_trainer), the way the reference's variables are simply shared by every graph."""
        self.dirty = False
        host = self._params_to_host() if self.params else {}
        for layer, p in self.params.items():
            w = host[layer][0].reshape(layer.k, layer.k, layer.k, layer.in_channels, layer.filters)
            if layer.transposed:
                w = w.transpose(0, 1, 2, 4, 3)
            layer.set_weights(np.ascontiguousarray(w), host[layer][1])
        self.model.entropy_bottleneck._invalidate()

    # -- tensor-core path: layers carry the current master weights ----------------------------------------
    def _refresh_layers(self):
        """Write the fp32 master weights into the layer objects (and their adjoint twins) so that the kernel dispatch packs
        this step's weights; one D2H per layer, packing happens lazily per kernel format."""
        host = self._params_to_host()
        for layer, p in self.params.items():
            w = host[layer][0].reshape(layer.k, layer.k, layer.k, layer.in_channels, layer.filters)
            kern = np.ascontiguousarray(w.transpose(0, 1, 2, 4, 3)) if layer.transposed else w
            layer.set_weights(kern, host[layer][1])
            twin = self.twins.get(layer)
            if twin is None:
                cls = MT.Conv3D if layer.transposed else MT.Conv3DTranspose
                twin = self.twins[layer] = cls(layer.in_channels, (layer.k,) * 3, strides=(layer.stride,) * 3, padding='same',
                                               data_format='channels_first', use_bias=False, activation=None)
            # conv (k,k,k,Cin,Cout) <-> transposed conv (k,k,k,out=Cin,in=Cout): the adjoint uses the same array
            twin.set_weights(kern, None)

    def _forward(self, name, x):
        steps, out_id = self.traces[name]
        if self.tensor_cores:
            for s in steps:
                if s[0] not in ('conv', 'add'):
                    raise NotImplementedError("residual_mode='concat' is not used by any reference config; training supports 'add'")
            keep, kv = {}, {}
            MT.run_steps([tuple(s) for s in steps], out_id, x, keep=keep, keep_vals=kv)
            keep[0] = x
            return keep[out_id], (steps, out_id, keep, kv)
        vals = {0: x}
        for s in steps:
            if s[0] == 'conv':
                _, layer, src, dst, _ = s
                p = self._p(layer, vals[src].shape[1])
                vals[dst] = ops.conv3d_f32(vals[src], p['w'], p['b'], layer.filters, layer.k, layer.stride, layer.transposed, layer.relu)
            elif s[0] == 'add':
                vals[s[3]] = ops.axpby(vals[s[1]], vals[s[2]], 1.0, 1.0)
            else:
                raise NotImplementedError("residual_mode='concat' is not used by any reference config; training supports 'add'")
        return vals[out_id], (steps, out_id, vals, {})

    def _wgrad(self, layer, x_in, g_val, x_val=None):
        """Weight gradient, tap-major (k^3, Cin, Cout).  g_val: the gradient w.r.t. the layer's pre-activation output as a _Val (fp32 and
        / or blocked); x_val: the forward pass' input activation as a _Val when it exists in the blocked layout.  Tensor-core mode: the
        3x3x3 layers with 16 / 32 / 64 channels run on the tcgen05 kernel in the active precision (bf16x3 by default) -- stride 1
        directly, stride 2 by phase decomposition; the one-channel ends and the 4^3-and-smaller volumes on the fp32 kernels."""
        n, cin, d, h, w = x_in.shape
        if self.tensor_cores:
            terms = {'bf16x3': 2, 'bf16': 1, 'fp32': 0}[MT.get_precision()]
            if terms and ops.wgrad_umma_eligible(n, cin, layer.filters, layer.k, layer.stride, d, h, w, terms):
                xb = x_val.as_blk(terms) if x_val is not None else ops.f32_to_blocked(x_in, terms)
                return ops.conv3d_wgrad_umma(xb, g_val.as_blk(terms), tuple(x_in.shape), layer.transposed, terms)
            if terms and layer.k == 3 and layer.stride == 2:
                dw = self._wgrad_stride2(layer, x_in, g_val, x_val, terms)
                if dw is not None:
                    return dw
        return ops.conv3d_wgrad_f32(x_in, g_val.as_f32(), layer.filters, layer.k, layer.stride, layer.transposed)

    @staticmethod
    def _wgrad_stride2(layer, x_in, g_val, x_val, terms):
        """Stride-2 layers on the stride-1 tcgen05 kernel by phase decomposition.  Both the stride-2 conv and the stride-2 transposed
        conv read their large tensor L (the conv's input / the transposed conv's output gradient) at 2b + t, t in {0, 1, 2} per axis,
        against the small tensor S at b:  dW[t] = sum_b L[2b + t] * S[b].  With the eight phase volumes L_p[b] = L[2b + p] stacked as
        channels this is a stride-1 correlation: t = 0 -> (phase 0, offset 0), t = 1 -> (phase 1, offset 0), t = 2 -> (phase 0, offset
        +1).  The phase channels are cut into chunks of C = channels of S (the kernel wants C in == C out); each chunk is one launch
        and the (phase, offset) blocks that exist are picked from its 27 x C x C result."""
        if isinstance(g_val, torch.Tensor):
            g_val = MT._Val(f32=g_val, shape=tuple(g_val.shape))
        small_shape = tuple(x_in.shape) if layer.transposed else g_val.shape
        large_shape = g_val.shape if layer.transposed else tuple(x_in.shape)
        n, c, sd, sh, sw = small_shape
        cb = large_shape[1]
        if c % cb or (8 * cb) % c or tuple(large_shape[2:]) != (2 * sd, 2 * sh, 2 * sw) or cb % 8 or c % 16:
            return None
        if not ops.wgrad_umma_eligible(n, c, c, 3, 1, sd, sh, sw, terms):
            return None
        ppc, nch = c // cb, 8 * cb // c          # phases per chunk, chunks
        large = g_val.as_f32() if layer.transposed else x_in
        phb = ops.f32_phases_to_blocked(large.contiguous(), c, terms)     # phase split + blocked layout in one pass
        if layer.transposed:
            sb = x_val.as_blk(terms) if x_val is not None else ops.f32_to_blocked(x_in, terms)
        else:
            sb = g_val.as_blk(terms)
        outs = [ops.conv3d_wgrad_umma(phb[j], sb, (n, c, sd, sh, sw), False, terms) for j in range(nch)]
        dw = torch.empty((27, x_in.shape[1], layer.filters), device=x_in.device, dtype=torch.float32)
        for t in range(27):
            tz, ty, tx = t // 9, (t // 3) % 3, t % 3
            phase = (int(tz == 1) * 2 + int(ty == 1)) * 2 + int(tx == 1)
            k = ((int(tz == 2) + 1) * 3 + int(ty == 2) + 1) * 3 + int(tx == 2) + 1
            blk = outs[phase // ppc][k, (phase % ppc) * cb:(phase % ppc + 1) * cb, :]      # (channels of L, channels of S)
            dw[t] = blk.t() if layer.transposed else blk
        return dw

    def _backward(self, tape, g_out, grads, need_input_grad=True):
        """Backward pass of one transform.  The gradients travel as _Val: in tensor-core mode they stay in the blocked bf16 layout
        between the layers (the data-gradient conv writes it, the ReLU mask / residual add / bias gradient / weight gradient read
        it); an fp32 copy is made only where an fp32 kernel needs one."""
        steps, out_id, vals, kv = tape
        V = MT._Val
        terms = {'bf16x3': 2, 'bf16': 1, 'fp32': 0}[MT.get_precision()] if self.tensor_cores else 0
        g = {out_id: V(f32=g_out, shape=tuple(g_out.shape))}

        def blocked(v):
            return terms and v.blk is not None and v.terms == terms

        def accumulate(vid, t):
            if vid not in g:
                g[vid] = t
            elif blocked(g[vid]) and blocked(t):
                g[vid] = V(blk=ops.add_blocked(g[vid].blk, t.blk, t.shape, terms), shape=t.shape, terms=terms)
            else:
                g[vid] = V(f32=ops.axpby(g[vid].as_f32(), t.as_f32(), 1.0, 1.0), shape=t.shape)

        for s in reversed(steps):
            if s[0] == 'add':
                gd = g.pop(s[3])
                accumulate(s[1], gd)
                accumulate(s[2], gd)
                continue
            _, layer, src, dst, _ = s
            gv = g.pop(dst)
            if layer.relu:
                yv = kv.get(dst)
                if yv is not None and blocked(yv) and (blocked(gv) or terms):
                    gv = V(blk=ops.relu_mask_blocked(gv.as_blk(terms), yv.blk, gv.shape, terms), shape=gv.shape, terms=terms)
                else:
                    gv = V(f32=ops.relu_bwd(gv.as_f32(), vals[dst]), shape=gv.shape)
            p = self.params[layer]
            x_in = vals[src]
            if p['b'] is None:
                db = None
            elif blocked(gv):
                db = ops.bias_grad_blocked(gv.blk, gv.shape, terms)
            else:
                db = ops.bias_grad_f32(gv.as_f32())
            grads[layer] = {'w': self._wgrad(layer, x_in, gv, kv.get(src)), 'b': db}
            if src != 0 or need_input_grad:
                # data gradient = the adjoint layer: conv <-> transposed conv, tap-major weights with the channel axes swapped
                if self.tensor_cores:
                    g[src] = MT.run_layer(self.twins[layer], gv, g.pop(src, None), return_val=True)
                else:
                    w_t = p['w'].transpose(1, 2).contiguous()
                    res = g.pop(src, None)
                    gx = ops.conv3d_f32(gv.as_f32(), w_t, None, layer.in_channels, layer.k, layer.stride, not layer.transposed, False,
                                        None if res is None else res.as_f32())
                    g[src] = V(f32=gx, shape=tuple(gx.shape))
        return g[0].as_f32() if 0 in g else None

    # -- entropy bottleneck helpers (host float64; C x 3 values) ----------------------------------------
    def _aux_loss_and_grad(self):
        """EntropyBottleneck.losses[0] = sum |logits(quantiles) - (-T,0,T)| and its gradient w.r.t. the quantiles
        (matrices / biases / factors are stop-gradient'ed in tfc)."""
        eb = self.model.entropy_bottleneck
        q = eb.quantiles.astype(np.float64)             # (C,1,3)
        target = math.log(2.0 / eb.tail_mass - 1.0)
        tgt = np.array([-target, 0.0, target])
        logits = q
        dl = np.ones_like(q)                            # d logits / d q, carried through the chain (per element)
        for i in range(len(eb.matrices)):
            M = _softplus(eb.matrices[i].astype(np.float64))
            pre = np.matmul(M, logits) + eb.biases[i].astype(np.float64)
            dpre = np.matmul(M, dl) if i > 0 else M * dl  # first layer: (C,3,1) x (C,1,3)
            if i < len(eb.factors):
                F = np.tanh(eb.factors[i].astype(np.float64))
                th = np.tanh(pre)
                logits = pre + F * th
                dl = dpre * (1.0 + F * (1.0 - th * th))
            else:
                logits, dl = pre, dpre
        diff = logits - tgt
        return float(np.abs(diff).sum()), (np.sign(diff) * dl).astype(np.float64)

    def _eb_raw_grads(self, dparams):
        """kernel gradients w.r.t. softplus(M), B, tanh(F) (C,44) -> gradients of the raw tfc variables."""
        eb = self.model.entropy_bottleneck
        C = eb.channels
        d = dparams.cpu().numpy().astype(np.float64)
        spl = [d[:, 0:3].reshape(C, 3, 1), d[:, 3:12].reshape(C, 3, 3), d[:, 12:21].reshape(C, 3, 3), d[:, 21:24].reshape(C, 1, 3)]
        gm = [g * _sigmoid(m.astype(np.float64)) for g, m in zip(spl, eb.matrices)]
        gb = [d[:, 24:27].reshape(C, 3, 1), d[:, 27:30].reshape(C, 3, 1), d[:, 30:33].reshape(C, 3, 1), d[:, 33:34].reshape(C, 1, 1)]
        gf = [d[:, 34 + 3 * i:37 + 3 * i].reshape(C, 3, 1) * (1.0 - np.tanh(eb.factors[i].astype(np.float64)) ** 2) for i in range(3)]
        return gm + gb + gf

    # -- data parallelism ------------------------------------------------------------------------------
    def _world(self):
        import torch.distributed as dist
        on = self.distributed if self.distributed is not None else (dist.is_available() and dist.is_initialized())
        return dist if (on and dist.get_world_size() > 1) else None

    @staticmethod
    def _allreduce_scalars(dist, *vals):
        t = torch.tensor(vals, dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    # -- one step ----------------------------------------------------------------------------------------
    def forward_backward(self, x, noise_y=None, noise_z=None):
        """Returns (values dict, grads dict): loss / fl / mbpov as python floats; grads[layer] = {'w','b'} (tap-major),
        grads['entropy_bottleneck'] = list of raw-variable gradients (matrices, biases, factors)."""
        import time
        m = self.model
        x = x.contiguous().float()
        eb = m.entropy_bottleneck
        grads = {}
        self.marks = [('start', time.perf_counter())]
        mark = lambda label: self.marks.append((label, time.perf_counter()))
        if self.tensor_cores:
            if not self.params:   # first step: materialise the master weights (layers are built by the model's train())
                for name, (steps, _) in self.traces.items():
                    for s in steps:
                        if s[0] == 'conv':
                            self._p(s[1], s[1].in_channels)
            self._refresh_layers()
        mark('weights refreshed')
        y, tape_a = self._forward('analysis', x)
        dist = self._world()
        n_occ = float(x.sum(dtype=torch.float64))
        if dist is not None:
            (n_occ,) = self._allreduce_scalars(dist, n_occ)   # the GLOBAL batch's occupied voxels (model_types.py:347)
        c = 1.0 / (-math.log(2.0) * n_occ)                 # d mbpov / d (sum ln p)
        if noise_y is None:
            noise_y = torch.rand_like(y) - 0.5
        y_tilde = ops.axpby(y, noise_y.contiguous().float(), 1.0, 1.0)
        ebp = eb.device_params()
        if self.v2:
            z, tape_ha = self._forward('hyper_analysis', y)
            if noise_z is None:
                noise_z = torch.rand_like(z) - 0.5
            z_tilde = ops.axpby(z, noise_z.contiguous().float(), 1.0, 1.0)
            _, sum_z = ops.eb_likelihood(z_tilde, ebp, want_likelihood=False)
            sigma, tape_hs = self._forward('hyper_synthesis', z_tilde)
            smin = float(np.float32(m.scale_table[0]))
            _, sum_y = ops.gc_likelihood(y_tilde, sigma, smin, want_likelihood=False)
        else:
            _, sum_y = ops.eb_likelihood(y_tilde, ebp, want_likelihood=False)
        x_tilde, tape_s = self._forward('synthesis', y_tilde)
        # the reported scalars are read back after the backward pass has been queued: a float() here would drain the GPU mid-step
        fl_t = ops.focal_loss_sum(x, x_tilde, self.gamma, self.alpha)
        mark('forward queued')
        # ---- backward
        g_xt = ops.focal_loss_bwd(x, x_tilde, self.gamma, self.alpha, self.lmbda)
        g_y = self._backward(tape_s, g_xt, grads)
        if self.v2:
            dv, dsig = ops.gc_likelihood_bwd(y_tilde, sigma, smin, c)
            g_y = ops.axpby(g_y, dv, 1.0, 1.0)
            g_z = self._backward(tape_hs, dsig, grads)
            dz, dpar = ops.eb_likelihood_bwd(z_tilde, ebp, c)
            g_z = ops.axpby(g_z, dz, 1.0, 1.0)
            g_y = ops.axpby(g_y, self._backward(tape_ha, g_z, grads), 1.0, 1.0)
        else:
            dv, dpar = ops.eb_likelihood_bwd(y_tilde, ebp, c)
            g_y = ops.axpby(g_y, dv, 1.0, 1.0)
        self._backward(tape_a, g_y, grads, need_input_grad=False)
        mark('backward queued')
        fl = float(fl_t[0])
        mark('backward done')
        ly, lz = float(sum_y[0]), float(sum_z[0]) if self.v2 else 0.0
        if dist is not None:
            fl, ly, lz = self._allreduce_scalars(dist, fl, ly, lz)   # reported values are those of the global batch
        mb_y = ly * c
        mb_z = lz * c if self.v2 else 0.0
        values = {'fl': fl, 'mbpov_y': mb_y, 'mbpov_z': mb_z, 'mbpov': mb_y + mb_z, 'loss': self.lmbda * fl + mb_y + mb_z,
                  'num_occupied_voxels': n_occ, 'x_tilde': x_tilde, 'y': y}
        if dist is not None:
            # one flat all-reduce (sum) of every gradient: conv kernels, biases and the entropy bottleneck's (C,44) block
            parts = [dpar.reshape(-1)]
            for layer in self.params:
                parts.append(grads[layer]['w'].reshape(-1))
                if grads[layer]['b'] is not None:
                    parts.append(grads[layer]['b'].reshape(-1))
            flat = torch.cat(parts)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            pos = 0
            for t in parts:
                t.copy_(flat[pos:pos + t.numel()])
                pos += t.numel()
        grads['entropy_bottleneck'] = self._eb_raw_grads(dpar)
        return values, grads

    def step(self, x, noise_y=None, noise_z=None):
        """sess.run(train_op): main Adam step, auxiliary Adam step on the quantiles, CDF-table refresh."""
        values, grads = self.forward_backward(x, noise_y, noise_z)
        self.t += 1
        for layer, p in self.params.items():
            gl = grads[layer]
            ops.adam_step(p['w'], gl['w'], p['mw'], p['vw'], self.lr, self.t)
            if p['b'] is not None:
                ops.adam_step(p['b'], gl['b'], p['mb'], p['vb'], self.lr, self.t)
        eb = self.model.entropy_bottleneck
        arrays = self._eb_arrays()
        if self.eb_adam is None:
            self.eb_adam = _HostAdam(arrays, self.lr)
            self.aux_adam = _HostAdam([eb.quantiles], self.aux_lr)
        aux, gq = self._aux_loss_and_grad()
        self.eb_adam.step(arrays, grads['entropy_bottleneck'])
        self.aux_adam.step([eb.quantiles], [gq])
        eb._invalidate()    # entropy_bottleneck.updates[0] (the quantised-CDF refresh) is lazy here: `tables` rebuilds them from the
                            # current variables on their next use (compress / decompress), not 64 host CDF quantisations per step
        values['aux_loss'] = aux
        values['step'] = self.t
        self.marks.append(('optimisers done', __import__('time').perf_counter()))
        self.dirty = True
        return values
