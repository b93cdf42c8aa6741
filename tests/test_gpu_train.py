"""tr_train.py path: forward loss, every parameter gradient and one optimiser step of the CUDA implementation against
torch autograd on the oracle (same weights, same U(-.5,.5) noise).

Two levels.  Kernel level (identical inputs): every backward kernel within 1e-5 of float64 autograd.  End to end: within 2e-3
of each gradient's scale -- the focal-loss gradient is ill-conditioned (g ~ -alpha/p for the many voxels with small p), so the
fp32 forward's ~1e-7 absolute error in x_tilde alone is a ~1e-4 relative change of the upstream gradient against a float64
forward; the hyper path, which does not see the focal loss, agrees to ~1e-5."""
import math

import numpy as np
import pytest
import torch

from oracle import entropy as E
from oracle import transforms as T
from oracle.model import CONFIGS, OracleModel, focal_loss as oracle_focal_loss, sparse_to_dense
from pcc_geo_cnn_v2_b200 import ops, synthetic
from pcc_geo_cnn_v2_b200.model_configs import ModelConfigType
from pcc_geo_cnn_v2_b200.training import Trainer

pytestmark = pytest.mark.gpu
DT = torch.float64


def _oracle_loss_and_grads(config, w, x, ny, nz, gamma, alpha, lmbda):
    """float64 autograd restatement of model_types.py:327-355 (V2) / 250-274 (V1)."""
    cfg = CONFIGS[config]
    f = cfg['num_filters']
    names = ['analysis', 'synthesis'] + (['hyper_analysis', 'hyper_synthesis'] if cfg['version'] == 2 else [])
    spec = {n: T.build_transform(cfg[n], f) for n in names}
    leaves = {n: [{'kernel': torch.tensor(l['kernel'], dtype=DT, requires_grad=True),
                   'bias': None if l['bias'] is None else torch.tensor(l['bias'], dtype=DT, requires_grad=True)} for l in w[n]]
              for n in names}
    ebw = w['entropy_bottleneck']
    eb = {k: [torch.tensor(a, dtype=DT, requires_grad=True) for a in ebw[k]] for k in ('matrices', 'biases', 'factors')}
    eb['quantiles'] = ebw['quantiles']
    xt = torch.tensor(x, dtype=DT)
    y = T.apply_transform(spec['analysis'], leaves['analysis'], xt)
    denom = -math.log(2) * xt.sum()
    if cfg['version'] == 2:
        z = T.apply_transform(spec['hyper_analysis'], leaves['hyper_analysis'], y)
        z_t, z_lik = E.eb_forward(eb, z, True, nz.to(DT), DT)
        sigma = T.apply_transform(spec['hyper_synthesis'], leaves['hyper_synthesis'], z_t)
        y_t, y_lik = E.gc_forward(y, sigma, E.make_scale_table(), True, ny.to(DT), DT)
        mbpov = torch.log(y_lik).sum() / denom + torch.log(z_lik).sum() / denom
    else:
        y_t, y_lik = E.eb_forward(eb, y, True, ny.to(DT), DT)
        mbpov = torch.log(y_lik).sum() / denom
    x_t = T.apply_transform(spec['synthesis'], leaves['synthesis'], y_t)
    fl = oracle_focal_loss(xt, x_t, gamma, alpha)
    loss = lmbda * fl + mbpov
    loss.backward()
    return {'loss': float(loss), 'fl': float(fl), 'mbpov': float(mbpov)}, leaves, eb


def _rel(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / (np.abs(want).max() + 1e-30))


@pytest.mark.parametrize('config,size,lmbda,batch,tc', [('c3p', 32, 3e-3, 2, False), ('c1', 32, 1e-3, 2, False),
                                                         ('c3p', 16, 3e-3, 32, False), ('c3p', 32, 3e-3, 2, True)])
def test_gradients_match_oracle_autograd(config, size, lmbda, batch, tc):
    """The third case runs the reference's training batch size (tr_train.py --batch_size 32): 32 x 64 latent channels; the
    fourth the tensor-core (bf16x3) forward / data-gradient convs, whose 1e-5 forward differences the focal loss amplifies
    to ~1e-2 in the gradients."""
    m = ModelConfigType[config].build()
    w = synthetic.trained_like_weights(m, seed=11, output_bias=-0.3)
    m.set_weights(w)
    blocks = synthetic.surface_blocks(batch, size=size, seed=5)
    x = np.concatenate([sparse_to_dense(b, (1, 1, size, size, size)) for b in blocks])
    g = torch.Generator().manual_seed(1)
    f = m.num_filters
    ny = torch.rand((batch, f) + (size // 8,) * 3, generator=g) - 0.5
    nz = torch.rand((batch, f) + (size // 16,) * 3, generator=g) - 0.5
    ref, leaves, eb = _oracle_loss_and_grads(config, w, x, ny, nz, 2, 0.75, lmbda)

    tr = Trainer(m, gamma=2, alpha=0.75, lmbda=lmbda, tensor_cores=tc)
    vals, grads = tr.forward_backward(torch.from_numpy(x).cuda(), ny.cuda(), nz.cuda())
    for k in ('loss', 'fl', 'mbpov'):
        assert abs(vals[k] - ref[k]) < 1e-4 * abs(ref[k]), (k, vals[k], ref[k])
    worst, report = 0.0, []
    for name, tf in m.transforms().items():
        for li, (layer, leaf) in enumerate(zip(tf.leaf_layers(), leaves[name])):
            gw = grads[layer]['w'].cpu().numpy().reshape(layer.k, layer.k, layer.k, layer.in_channels, layer.filters)
            if layer.transposed:
                gw = gw.transpose(0, 1, 2, 4, 3)          # back to the Keras Conv3DTranspose layout
            r = _rel(gw, leaf['kernel'].grad.numpy())
            rb = 0.0 if leaf['bias'] is None else _rel(grads[layer]['b'].cpu().numpy(), leaf['bias'].grad.numpy())
            worst = max(worst, r, rb)
            report.append(f'{name}[{li}] k{layer.k} s{layer.stride} {layer.in_channels}->{layer.filters}: w {r:.2e} b {rb:.2e}')
    print('\n'.join(report))
    assert worst < (2e-2 if tc else 2e-3), report
    ge = grads['entropy_bottleneck']
    want = [t.grad.numpy() for t in eb['matrices'] + eb['biases'] + eb['factors']]
    assert len(ge) == len(want) == 11
    for i, (a, b) in enumerate(zip(ge, want)):
        assert a.shape == b.shape
        assert _rel(a, b) < (1e-2 if tc else 1e-3), ('entropy bottleneck variable', i, _rel(a, b))
    print('worst conv-kernel gradient error', worst)


def test_backward_kernels_on_identical_inputs():
    rng = np.random.default_rng(0)
    # ---- focal loss
    xt = (rng.random((2, 1, 16, 16, 16)) < 0.05).astype(np.float32)
    xp = rng.uniform(0.0, 1.3, size=xt.shape).astype(np.float32)
    xp[rng.random(xt.shape) < 0.4] = 0.0
    p = torch.tensor(xp, dtype=DT, requires_grad=True)
    oracle_focal_loss(torch.tensor(xt, dtype=DT), p, 2, 0.75).backward()
    got = ops.focal_loss_bwd(torch.from_numpy(xt).cuda(), torch.from_numpy(xp).cuda(), 2, 0.75, 0.5).cpu().numpy()
    assert _rel(got, 0.5 * p.grad.numpy()) < 1e-6
    # ---- Gaussian conditional (values and scales on both sides of the bounds)
    st = E.make_scale_table()
    v = (rng.normal(size=(2, 8, 4, 4, 4)) * 6).astype(np.float32)
    sg = np.exp(rng.uniform(np.log(0.01), np.log(60), size=v.shape)).astype(np.float32)
    sg[rng.random(v.shape) < 0.3] = 0.0                       # ReLU'd scales: below the 0.11 bound
    vt, sgt = torch.tensor(v, dtype=DT, requires_grad=True), torch.tensor(sg, dtype=DT, requires_grad=True)
    (0.37 * torch.log(E.gc_likelihood(vt, sgt, st, DT)).sum()).backward()
    dv, ds = ops.gc_likelihood_bwd(torch.from_numpy(v).cuda(), torch.from_numpy(sg).cuda(), float(np.float32(st[0])), 0.37)
    assert _rel(dv.cpu().numpy(), vt.grad.numpy()) < 1e-5 and _rel(ds.cpu().numpy(), sgt.grad.numpy()) < 1e-5
    # ---- entropy bottleneck (values + all 11 variables)
    m = ModelConfigType['c2'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=9))
    eb = m.entropy_bottleneck
    w = eb.get_weights()
    leaves = {k: [torch.tensor(a, dtype=DT, requires_grad=True) for a in w[k]] for k in ('matrices', 'biases', 'factors')}
    leaves['quantiles'] = w['quantiles']
    z = (rng.normal(size=(3, 32, 2, 2, 2)) * 7).astype(np.float32)
    zt = torch.tensor(z, dtype=DT, requires_grad=True)
    lik = E.eb_likelihood_c1m(leaves, zt.transpose(0, 1).reshape(32, 1, -1), DT)
    (-0.21 * torch.log(lik).sum()).backward()
    tr = Trainer(m)
    dz, dpar = ops.eb_likelihood_bwd(torch.from_numpy(z).cuda(), eb.device_params(), -0.21)
    assert _rel(dz.cpu().numpy(), zt.grad.numpy()) < 2e-5
    for a, b in zip(tr._eb_raw_grads(dpar), [t.grad.numpy() for t in leaves['matrices'] + leaves['biases'] + leaves['factors']]):
        assert _rel(a, b) < 2e-5
    # ---- weight / bias / data gradients of the convolutions
    for transposed, k, s, cin, cout, shape in [(False, 3, 1, 16, 16, (8, 8, 8)), (False, 3, 2, 16, 32, (8, 8, 8)), (True, 3, 2, 32, 16, (4, 4, 4)),
                                               (True, 3, 1, 16, 1, (8, 8, 8)), (False, 3, 2, 1, 16, (8, 8, 8)), (True, 3, 1, 16, 1, (5, 6, 40)), (False, 9, 2, 1, 8, (8, 8, 8)), (True, 5, 2, 8, 8, (4, 4, 4))]:
        x = torch.tensor(rng.normal(size=(2, cin) + shape), dtype=DT, requires_grad=True)
        kshape = (k, k, k, cout, cin) if transposed else (k, k, k, cin, cout)
        kern = torch.tensor(rng.normal(size=kshape) / np.sqrt(k ** 3 * cin), dtype=DT, requires_grad=True)
        y = (T.conv3d_transpose_same if transposed else T.conv3d_same)(x, kern, None, s, False)
        gy = rng.normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.tensor(gy, dtype=DT))
        xd, gd = x.detach().float().cuda(), torch.from_numpy(gy).cuda()
        dw = ops.conv3d_wgrad_f32(xd, gd, cout, k, s, transposed).cpu().numpy().reshape(k, k, k, cin, cout)
        want_w = kern.grad.numpy().transpose(0, 1, 2, 4, 3) if transposed else kern.grad.numpy()
        assert _rel(dw, want_w) < 1e-5, ('wgrad', transposed, k, s)
        w_tap = (kern.detach().permute(0, 1, 2, 4, 3) if transposed else kern.detach()).reshape(k ** 3, cin, cout).float().cuda()
        dx = ops.conv3d_f32(gd, w_tap.transpose(1, 2).contiguous(), None, cin, k, s, not transposed, False)
        assert _rel(dx.cpu().numpy(), x.grad.numpy()) < 1e-5, ('dgrad', transposed, k, s)
        assert _rel(ops.bias_grad_f32(gd).cpu().numpy(), gy.astype(np.float64).sum((0, 2, 3, 4))) < 1e-6
    yv = torch.from_numpy(rng.normal(size=(1000,)).astype(np.float32)).cuda()
    gv = torch.from_numpy(rng.normal(size=(1000,)).astype(np.float32)).cuda()
    assert torch.equal(ops.relu_bwd(gv, yv), torch.where(yv > 0, gv, torch.zeros_like(gv)))
    assert torch.allclose(ops.axpby(gv, yv, 2.0, -0.5), 2 * gv - 0.5 * yv)


def test_aux_loss_gradient_and_adam_step():
    m = ModelConfigType['c2'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=3))
    tr = Trainer(m, gamma=2, alpha=0.9, lmbda=1e-3)
    ebw = m.entropy_bottleneck.get_weights()
    q = torch.tensor(ebw['quantiles'], dtype=DT, requires_grad=True)
    p = dict(ebw)
    p['quantiles'] = q
    loss = E.eb_aux_loss(p, DT)
    loss.backward()
    aux, gq = tr._aux_loss_and_grad()
    assert abs(aux - float(loss)) < 1e-9 * abs(float(loss))
    assert _rel(gq, q.grad.numpy()) < 1e-9
    # Adam kernel == TF1 formula
    rng = np.random.default_rng(0)
    theta0, grad = rng.normal(size=1000).astype(np.float32), rng.normal(size=1000).astype(np.float32)
    th, g_ = torch.from_numpy(theta0.copy()).cuda(), torch.from_numpy(grad).cuda()
    mm, vv = torch.zeros_like(th), torch.zeros_like(th)
    m_ref, v_ref, t_ref = np.zeros(1000), np.zeros(1000), theta0.astype(np.float64)
    for t in (1, 2, 3):
        ops.adam_step(th, g_, mm, vv, 1e-4, t)
        m_ref = 0.9 * m_ref + 0.1 * grad
        v_ref = 0.999 * v_ref + 0.001 * grad.astype(np.float64) ** 2
        t_ref -= 1e-4 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m_ref / (np.sqrt(v_ref) + 1e-8)
    assert np.abs(th.cpu().numpy() - t_ref).max() < 1e-6


def test_training_steps_reduce_the_loss_and_keep_the_codec_consistent():
    m = ModelConfigType['c3p'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=5))
    blocks = synthetic.surface_blocks(4, size=32, seed=2)
    x = torch.from_numpy(np.concatenate([sparse_to_dense(b, (1, 1, 32, 32, 32)) for b in blocks])).cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    ny = torch.rand((4, 64, 4, 4, 4), generator=g, device='cuda') - 0.5
    nz = torch.rand((4, 64, 2, 2, 2), generator=g, device='cuda') - 0.5
    tr = Trainer(m, gamma=2, alpha=0.75, lmbda=1e-2)           # Adam 1e-4 / 1e-3 like the reference
    losses = [tr.step(x, ny, nz)['loss'] for _ in range(8)]
    assert losses[-1] < losses[0], losses
    assert all(math.isfinite(v) for v in losses)
    # trained parameters flow back into the codec path
    tr.sync_to_model()
    m.compress((1, 1, 32, 32, 32))
    data, meta, _ = m.compress_blocks(None, blocks, None, None, 32, 0, fixed_threshold=True)
    dec, _ = m.decompress_blocks(None, data[0], (32, 32, 32))
    for a, b in zip(meta[0]['x_hat_list'], dec):
        assert np.array_equal(a, b)


def test_train_op_weights_are_visible_to_the_model_without_manual_sync():
    """tr_train.py flow (tr_train.py:91-134): m.train(...) builds the graph, m.train_op(x) steps it, and the SAME variables are
    then read by the validation forward, by get_weights() (checkpoint) and by the codec loops -- no explicit sync call."""
    m = ModelConfigType['c3p'].build()
    m.set_weights(synthetic.trained_like_weights(m, seed=5))
    blocks = synthetic.surface_blocks(4, size=32, seed=2)
    x = torch.from_numpy(np.concatenate([sparse_to_dense(b, (1, 1, 32, 32, 32)) for b in blocks])).cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    ny = torch.rand((4, 64, 4, 4, 4), generator=g, device='cuda') - 0.5
    nz = torch.rand((4, 64, 2, 2, 2), generator=g, device='cuda') - 0.5
    m.train(x, 2, 0.75, 1e-2, noise_y=ny, noise_z=nz)
    loss0 = float(m.train_loss)
    w0 = m.get_weights()['synthesis'][0]['kernel'].copy()
    out = [m.train_op(x, ny, nz) for _ in range(6)]
    assert abs(out[0]['loss'] - loss0) < 1e-3 * abs(loss0)        # train_op's forward == the model's forward before any step
    w1 = m.get_weights()['synthesis'][0]['kernel']
    assert np.abs(w1 - w0).max() > 1e-5                             # get_weights() returns the trained variables
    m.train(x, 2, 0.75, 1e-2, noise_y=ny, noise_z=nz)               # validation forward on the trained variables
    nxt = m.train_op.__self__.forward_backward(x, ny, nz)[0]['loss']
    assert abs(float(m.train_loss) - nxt) < 1e-3 * abs(nxt), (float(m.train_loss), nxt)
    assert float(m.train_loss) < loss0
    m.compress((1, 1, 32, 32, 32))                                  # and the codec loops run on them, encoder == decoder
    data, meta, _ = m.compress_blocks(None, blocks, None, None, 32, 0, fixed_threshold=True)
    dec, _ = m.decompress_blocks(None, data[0], (32, 32, 32))
    for a, b in zip(meta[0]['x_hat_list'], dec):
        assert np.array_equal(a, b)
    o = OracleModel('c3p')
    w = m.get_weights()
    o.set_params({k: v for k, v in w.items() if k != 'entropy_bottleneck'}, w['entropy_bottleneck'])
    ref = o.train_forward(x.cpu().numpy(), 2, 0.75, 1e-2, ny.cpu(), nz.cpu())
    assert abs(float(m.train_loss) - float(ref['loss'])) < 1e-3 * abs(float(ref['loss']))
