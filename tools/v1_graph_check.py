"""Graph vs eager block loops over several batches in flight, for the V1 (one latent) and V2 model types, with an output
bias that makes the decoded point sets non-empty."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_geo_cnn_v2_b200 as P
from pcc_geo_cnn_v2_b200 import synthetic
for cfg in ('c1', 'c2'):
    for ob in (0.4, 0.48, 0.5, 0.55):
        m = P.ModelConfigType[cfg].build(batch_size=3)
        m.set_weights(synthetic.trained_like_weights(m, seed=7, output_bias=ob))
        m.compress((1, 1, 64, 64, 64))
        blocks = synthetic.surface_blocks(8, size=64, seed=21)
        res = {}
        for g in (True, False):
            m.use_graphs = g
            dl, meta, _ = m.compress_blocks(None, blocks, None, None, 64, 0, fixed_threshold=True)
            dec, _ = m.decompress_blocks(None, dl[0], (64, 64, 64))
            res[g] = (dl, meta[0]['x_hat_list'], dec)
        same_dec = [np.array_equal(a, b) for a, b in zip(res[True][2], res[False][2])]
        same_enc = [np.array_equal(a, b) for a, b in zip(res[True][1], res[False][1])]
        print(cfg, 'output_bias', ob, 'enc==', all(same_enc), 'dec==', same_dec, 'pts', [len(d) for d in res[False][2]])
