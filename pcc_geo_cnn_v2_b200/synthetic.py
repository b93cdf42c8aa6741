"""Seeded synthetic inputs and "trained-like" synthetic parameters (no datasets / checkpoints are available
offline).  Used by bench.py, the tests and __graft_entry__.smoke().

Keras-default initialisation (Glorot-uniform, zero bias) drives sparse occupancy inputs to all-zero latents,
which exercises nothing; `trained_like_weights` rescales the same Glorot draws (gain 1.8, synthesis 1.6), adds
small biases and a calibrated output bias so that y spreads over roughly +-6, z over +-13 (beyond the factorized
prior's +-10 table -> escape codes), scale indexes cover half of the table and ~2.5 % of x_hat exceeds the 0.5
threshold (a decoded point count of the order of the input's, like a trained codec)."""
import math

import numpy as np


def surface_blocks(n_blocks, size=64, seed=42):
    """Voxelised random surfaces (sphere shells and planes), ~2-3 % occupancy like the ModelNet40 blocks the
    reference trains on.  Returns a list of float32 (n_i, 3) arrays of unique coords."""
    rng = np.random.default_rng(seed)
    g = np.indices((size, size, size)).astype(np.float32)
    out = []
    for _ in range(n_blocks):
        occ = np.zeros((size, size, size), bool)
        for _ in range(int(rng.integers(1, 3))):
            if rng.random() < 0.5:
                c = rng.uniform(0.2 * size, 0.8 * size, 3).astype(np.float32)
                r = rng.uniform(0.2 * size, 0.45 * size)
                d = np.sqrt(((g - c[:, None, None, None]) ** 2).sum(0))
                occ |= np.abs(d - r) < 0.6
            else:
                nrm = rng.normal(size=3).astype(np.float32)
                nrm /= np.linalg.norm(nrm)
                off = rng.uniform(0.3 * size, 0.7 * size)
                occ |= np.abs(np.tensordot(nrm, g - size / 2, axes=(0, 0)) + size / 2 - off) < 0.5
        pts = np.argwhere(occ).astype(np.float32)
        if len(pts) == 0:
            pts = np.array([[size // 2] * 3], np.float32)
        out.append(pts)
    return out


def trained_like_weights(model, seed=42, gain=1.8, synthesis_gain=1.6, bias_scale=0.03, output_bias=-0.7):
    """Deterministic parameter set for a pcc_geo_cnn_v2_b200 model (dict accepted by model.set_weights)."""
    rng = np.random.default_rng(seed)
    f = model.num_filters
    in_ch = {'analysis': 1, 'synthesis': f, 'hyper_analysis': f, 'hyper_synthesis': f}
    w = {}
    for name, tf in model.transforms().items():
        c = in_ch[name]
        g = synthesis_gain if name == 'synthesis' else gain
        layers = []
        for layer in tf.leaf_layers():
            k, fo = layer.k, layer.filters
            limit = g * math.sqrt(6.0 / (k ** 3 * (c + fo)))
            shape = (k, k, k, fo, c) if layer.transposed else (k, k, k, c, fo)
            kern = rng.uniform(-limit, limit, size=shape).astype(np.float32)
            bias = rng.uniform(-bias_scale, bias_scale, size=(fo,)).astype(np.float32) if layer.use_bias else None
            layers.append({'kernel': kern, 'bias': bias})
            c = fo
        if name == 'synthesis' and layers[-1]['bias'] is not None:
            layers[-1]['bias'] = layers[-1]['bias'] + np.float32(output_bias)  # ~2.5 % of x_hat above the 0.5 threshold
        w[name] = layers
    eb = model.entropy_bottleneck
    eb.build(f)
    ew = eb.get_weights()
    ew = {'matrices': [m.copy() for m in ew['matrices']],
          'biases': [rng.uniform(-0.5, 0.5, size=b.shape).astype(np.float32) for b in ew['biases']],
          'factors': [rng.uniform(-0.2, 0.2, size=t.shape).astype(np.float32) for t in ew['factors']],
          'quantiles': ew['quantiles'].copy()}
    ew['quantiles'][:, 0, 1] = rng.uniform(-0.4, 0.4, size=f).astype(np.float32)  # non-trivial medians
    w['entropy_bottleneck'] = ew
    return w


def codec_like_weights(model, seed=42, latent_gain=0.15, hyper_gain=0.15, sigma_gain=0.1, sigma_bias=0.2, prior_scale=0.1,
                       prior_support=3.0, synthesis_in_gain=6.0, output_bias=-0.277):
    """A rate-realistic operating point: `trained_like_weights` rescaled so that the bitstream looks like a trained codec's
    (the reference works at 0.27-0.88 bits per input point, /root/reference/data.csv:243-246): ~98 % of the latents
    quantise to 0, the rest to +-1, no escape codes; predicted scales sit around 0.2 (low table rows); the factorized prior is
    narrow (logistic scale ~0.1, support +-3).  On the synthetic surface blocks this gives 0.7-1.4 KB per block (0.7-1.0 bpp)
    instead of trained_like_weights' ~30 KB of escape-heavy symbols, with a decoded point count of the order of the input's
    (the first synthesis layer is scaled back up and the output bias re-calibrated: ~2.5 % of x_hat above 0.5).
    Calibrated on the c3p (paper c4) network with the CPU oracle; bench.py's `e2e_realistic` uses it."""
    w = trained_like_weights(model, seed=seed)
    w['analysis'][-1]['kernel'] = w['analysis'][-1]['kernel'] * np.float32(latent_gain)
    w['synthesis'][0]['kernel'] = w['synthesis'][0]['kernel'] * np.float32(synthesis_in_gain)
    w['synthesis'][-1]['bias'] = np.full_like(w['synthesis'][-1]['bias'], output_bias)
    if 'hyper_analysis' in w:
        w['hyper_analysis'][-1]['kernel'] = w['hyper_analysis'][-1]['kernel'] * np.float32(hyper_gain)
        w['hyper_synthesis'][-1]['kernel'] = w['hyper_synthesis'][-1]['kernel'] * np.float32(sigma_gain)
        w['hyper_synthesis'][-1]['bias'] = np.full_like(w['hyper_synthesis'][-1]['bias'], sigma_bias)
    eb = w['entropy_bottleneck']
    r = (1, 3, 3, 3, 1)
    scale = prior_scale ** 0.25
    for i in range(4):
        eb['matrices'][i] = np.full_like(eb['matrices'][i], math.log(math.expm1(1.0 / scale / r[i + 1])))
    eb['quantiles'] = eb['quantiles'].copy()
    eb['quantiles'][:, 0, 0] = eb['quantiles'][:, 0, 1] - np.float32(prior_support)
    eb['quantiles'][:, 0, 2] = eb['quantiles'][:, 0, 1] + np.float32(prior_support)
    return w
