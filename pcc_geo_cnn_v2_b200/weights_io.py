"""Weight import / export: what `saver.restore(sess, checkpoint)` does for the reference's scripts
(src/compress_octree.py:82-92, src/decompress_octree.py:40-49, src/tr_train.py:58-76).

    load_weights(model, 'weights.npz' | dict)        save_weights(model, 'weights.npz')        variable_names(model)

There is no TensorFlow here, so a TF1 checkpoint is read elsewhere (three lines with TF installed:
`r = tf.train.load_checkpoint(ckpt); np.savez(out, **{k: r.get_tensor(k) for k in r.get_variable_to_shape_map()})`) and
arrives as an .npz / dict of arrays keyed by the TF variable names.  Names follow Keras' auto-naming rules (SURVEY.md
Appendix B.4; unverifiable offline -- no checkpoint is shipped with the reference):

  * layers are numbered per class in creation order over the whole model: `conv3d, conv3d_1, ...`,
    `conv3d_transpose, conv3d_transpose_1, ...`; the reference builds analysis, synthesis, hyper-analysis, hyper-synthesis
    in that order (src/model_types.py:252-254, 329-333);
  * variables are `<scopes>/<layer>/kernel` and `<scopes>/<layer>/bias` with the Keras layouts
    (Conv3D (kd,kh,kw,Cin,Cout), Conv3DTranspose (kd,kh,kw,Cout,Cin));
  * tfc 1.3 EntropyBottleneck: `entropy_bottleneck/matrix_i`, `bias_i`, `factor_i`, `quantiles`.

Matching is by the trailing `<layer>/<variable>` part, so the enclosing scopes (`analysis_transform_progressive_v2/
analysis_block/...`), a `:0` suffix and optimizer slots (`.../Adam`, `.../Adam_1`) do not matter.  Every model variable must be
found with the right shape, otherwise nothing is changed and a KeyError / ValueError names the first problem.
"""
import re

import numpy as np


def _snake(name):
    s = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', name)
    return re.sub('([a-z])([A-Z])', r'\1_\2', s).lower()


def variable_names(model):
    """[(tf variable name, kind, owner, expected shape)] in creation order.  kind: 'kernel' | 'bias' (owner = conv layer),
    'matrix' | 'ebias' | 'factor' | 'quantiles' (owner = (entropy bottleneck, i))."""
    out = []
    counters = {}

    def auto(cls):
        k = counters.get(cls, 0)
        counters[cls] = k + 1
        return cls if k == 0 else f'{cls}_{k}'

    from .model_transforms import ResidualLayer
    for tname, tf in model.transforms().items():
        scope = _snake(type(tf).__name__)
        bcount = {}
        for sub in tf._layers:
            if isinstance(sub, ResidualLayer):
                b = _snake(type(sub).__name__)
                k = bcount.get(b, 0)
                bcount[b] = k + 1
                prefix = f'{scope}/{b if k == 0 else f"{b}_{k}"}'
            else:
                prefix = scope
            for layer in sub.leaf_layers():
                name = auto('conv3d_transpose' if layer.transposed else 'conv3d')
                if layer.kernel is None:
                    raise ValueError(f'{tname}: layers are not built yet (call the transform once or set_weights first)')
                out.append((f'{prefix}/{name}/kernel', 'kernel', layer, tuple(layer.kernel.shape)))
                if layer.use_bias:
                    out.append((f'{prefix}/{name}/bias', 'bias', layer, (layer.filters,)))
    eb = model.entropy_bottleneck
    for i, m in enumerate(eb.matrices):
        out.append((f'entropy_bottleneck/matrix_{i}', 'matrix', (eb, i), tuple(m.shape)))
        out.append((f'entropy_bottleneck/bias_{i}', 'ebias', (eb, i), tuple(eb.biases[i].shape)))
        if i < len(eb.factors):
            out.append((f'entropy_bottleneck/factor_{i}', 'factor', (eb, i), tuple(eb.factors[i].shape)))
    out.append(('entropy_bottleneck/quantiles', 'quantiles', (eb, 0), tuple(eb.quantiles.shape)))
    return out


def _tail(name):
    name = name.split(':')[0]
    parts = name.split('/')
    return '/'.join(parts[-2:])


def _ensure_built(model):
    """Layers create their variables at first call (Keras semantics): derive every layer's input channels from the config."""
    f = model.num_filters
    in_ch = {'analysis': 1, 'synthesis': f, 'hyper_analysis': f, 'hyper_synthesis': f}
    from .model_transforms import ResidualLayer, _ConvBase

    def rec(l, c):
        if isinstance(l, _ConvBase):
            l.build(c)
            return l.filters
        if isinstance(l, ResidualLayer):
            c1 = rec(l._layers[0], c)
            c2 = c1
            for sub in l._layers[1:]:
                c2 = rec(sub, c2)
            return c1 if l.residual_mode == 'add' else c1 + c2
        for sub in l._layers:
            c = rec(sub, c)
        return c

    for name, tf in model.transforms().items():
        rec(tf, in_ch[name])
    model.entropy_bottleneck.build(f)


def state_dict(model):
    """{tf variable name: numpy array} of every model variable (trained values included: see CompressionModel.get_weights)."""
    model._sync_trainer()
    _ensure_built(model)
    sd = {}
    for name, kind, owner, _ in variable_names(model):
        if kind == 'kernel':
            sd[name] = owner.kernel
        elif kind == 'bias':
            sd[name] = owner.bias
        else:
            eb, i = owner
            sd[name] = {'matrix': eb.matrices, 'ebias': eb.biases, 'factor': eb.factors}[kind][i] if kind != 'quantiles' else eb.quantiles
    return sd


def save_weights(model, path):
    np.savez(path, **state_dict(model))


def load_weights(model, source):
    """source: path of an .npz, or a dict {variable name: array}.  Returns the list of variable names that were set."""
    _ensure_built(model)
    if not isinstance(source, dict):
        with np.load(source) as z:
            source = {k: z[k] for k in z.files}
    by_tail = {}
    for k, v in source.items():
        if re.search(r'/Adam(_\d+)?(:\d+)?$', k) or k.split(':')[0].split('/')[-1] in ('beta1_power', 'beta2_power', 'global_step'):
            continue   # optimizer slots of a training checkpoint
        by_tail.setdefault(_tail(k), []).append(k)
    plan = []
    for name, kind, owner, shape in variable_names(model):
        keys = by_tail.get(_tail(name))
        if not keys:
            raise KeyError(f'no variable matching {_tail(name)!r} (for {name}) in the weights source')
        if len(keys) > 1:
            exact = [k for k in keys if k.split(':')[0] == name]
            if len(exact) != 1:
                raise KeyError(f'{_tail(name)!r} is ambiguous in the weights source: {keys}')
            keys = exact
        arr = np.asarray(source[keys[0]], np.float32)
        if tuple(arr.shape) != shape:
            raise ValueError(f'{keys[0]}: shape {tuple(arr.shape)}, the model expects {shape}')
        plan.append((name, kind, owner, arr))
    kernels, biases = {}, {}
    ebw = {k: list(v) if isinstance(v, list) else v for k, v in model.entropy_bottleneck.get_weights().items()}
    for name, kind, owner, arr in plan:
        if kind == 'kernel':
            kernels[owner] = arr
        elif kind == 'bias':
            biases[owner] = arr
        elif kind == 'quantiles':
            ebw['quantiles'] = arr
        else:
            ebw[{'matrix': 'matrices', 'ebias': 'biases', 'factor': 'factors'}[kind]][owner[1]] = arr
    for layer, k in kernels.items():
        layer.set_weights(k, biases.get(layer))
    model.entropy_bottleneck.set_weights(ebw)
    model.trainer = None
    return [p[0] for p in plan]
