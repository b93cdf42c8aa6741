"""Host logic of compress_blocks against the REFERENCE's own compress_blocks / select_best_per_opt_metric run with a fake
session (tests/golden/make_reference_block_loop_fixture.py: prescribed x_hat volumes and strings instead of the networks):
thresholds per block, the selected opt_metric, whole-cloud metrics, the per-block point arrays and the data_list structure."""
import os

import numpy as np
import pytest
import torch

from pcc_geo_cnn_v2_b200 import ModelConfigType, ops
from pcc_geo_cnn_v2_b200.model_types import threshold_f32

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_block_loops.npz'))


def _split(flat, lens):
    out, pos = [], 0
    for n in lens:
        out.append(flat[pos:pos + int(n)])
        pos += int(n)
    return out


@pytest.mark.parametrize('tag', ['fixed', 'adaptive'])
def test_compress_blocks_host_logic_equals_the_reference(tag):
    res, level, bs = int(G['res']), int(G['level']), int(G['bs'])
    blocks = _split(G['blocks'], G['block_len'])
    sb = _split(G['str_bytes'].tobytes(), G['str_lens'])
    strings = [(sb[2 * j], sb[2 * j + 1]) for j in range(len(blocks))]
    x_hat = torch.from_numpy(G['x_hats'].astype(np.float32)[:, None]).cuda()
    m = ModelConfigType['c3p'].build(batch_size=5)
    m.compress((1, 1, bs, bs, bs))

    def fake_encode_blocks(blks, x_shape=None, thr_idx=None, keep_x_hat=True):   # the networks' outputs are prescribed
        pts = None
        if thr_idx is not None:
            bits, _ = ops.threshold_pack(x_hat, torch.from_numpy(threshold_f32(m.thresholds, thr_idx)).cuda())
            pts = ops.bits_to_points(bits.cpu().numpy(), (bs, bs, bs))
        return strings, (x_hat if keep_x_hat else None), pts

    m.encode_blocks = fake_encode_blocks
    kw = dict(fixed_threshold=True) if tag == 'fixed' else dict(fixed_threshold=False, opt_metrics=['d1_mse', 'd1_sum_mean'], max_deltas=[np.inf, 1.3])
    data_list, metadata, _ = m.compress_blocks(None, blocks, list(G['binstr']), G['points'], res, level, **kw)
    assert len(metadata) == 1 and len(data_list) == 1
    md = metadata[0]
    assert int(md['idx']) == int(G[f'{tag}_idx'])
    assert [int(t) for _, t in data_list[0]] == list(G[f'{tag}_thr'])
    assert [tuple(s) for s, _ in data_list[0]] == strings
    got_pts = md['x_hat_list']
    assert [len(p) for p in got_pts] == list(G[f'{tag}_pts_len'])
    assert np.array_equal(np.vstack(got_pts), G[f'{tag}_pts']) and got_pts[0].dtype == np.float32
    assert np.array_equal(np.asarray(md['blocks_full'], np.float32), G[f'{tag}_full'])
    want = G[f'{tag}_metrics']
    got = [md['metrics'][k] for k in ('d1_sum_AB', 'd1_sum_BA', 'd1_mse', 'd1_psnr')]
    assert np.allclose(got, want, rtol=1e-9), (got, want)
