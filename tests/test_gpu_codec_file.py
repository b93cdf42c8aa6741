"""compress_octree.py -> decompress_octree.py data flow on this implementation: partition_octree -> compress_blocks ->
save_compressed_file (+gzip) -> load_compressed_file -> decompress_blocks -> departition_octree
(reference src/compress_octree.py:86-113, src/decompress_octree.py:30-60,127-140)."""
import gzip
import io

import numpy as np
import pytest

from pcc_geo_cnn_v2_b200 import ModelConfigType, synthetic
from pcc_geo_cnn_v2_b200 import model_syntax as MS
from pcc_geo_cnn_v2_b200 import octree_coding as OC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fixed_threshold', [True, False])
def test_point_cloud_file_round_trip(fixed_threshold):
    res, level, bs = 256, 2, 64
    rng = np.random.default_rng(4)
    u = rng.random((60000, 2))
    pts = np.stack([u[:, 0] * (res - 1), (np.sin(u[:, 0] * 5) * 0.3 + 0.5) * (res - 1) * (0.3 + 0.7 * u[:, 1]), u[:, 1] * (res - 1)], 1)
    pts = np.unique(pts.astype(np.int64), axis=0).astype(np.float64)
    blocks, binstr = OC.partition_octree(pts, [0, 0, 0], [res] * 3, level)
    assert len(blocks) > 10 and all(b.max() < bs for b in blocks)

    m = ModelConfigType['c3p'].build(batch_size=8)
    m.set_weights(synthetic.trained_like_weights(m, seed=3, output_bias=-0.45))
    m.compress((1, 1, bs, bs, bs))
    data_list, metadata, _ = m.compress_blocks(None, blocks, binstr, pts, res, level, opt_metrics=('d1_mse',), max_deltas=(np.inf,),
                                               fixed_threshold=fixed_threshold)
    blob = gzip.compress(MS.save_compressed_file(binstr, data_list[0], res, level))

    res2, level2, binstr2, blocks2 = MS.load_compressed_file(io.BytesIO(gzip.decompress(blob)))
    assert (int(res2), int(level2)) == (res, level) and list(binstr2) == list(binstr)
    m2 = ModelConfigType['c3p'].build(batch_size=8)
    m2.set_weights(synthetic.trained_like_weights(m2, seed=3, output_bias=-0.45))
    m2.decompress()
    dec_blocks, _ = m2.decompress_blocks(None, blocks2, (bs, bs, bs))
    # the decoder reproduces the encoder's own reconstruction, block by block (decompress_octree.py --debug contract)
    for a, b in zip(metadata[0]['x_hat_list'], dec_blocks):
        assert np.array_equal(a, b)
    if not fixed_threshold:
        assert len({int(t) for _, t in data_list[0]}) > 1   # adaptive thresholds really differ between blocks
    nonempty = [b for b in dec_blocks if len(b)]
    cloud = np.vstack(OC.departition_octree(dec_blocks, list(binstr2), [0, 0, 0], [res] * 3, level)) if nonempty else np.zeros((0, 3))
    assert cloud.shape[1] == 3 and (cloud >= 0).all() and (cloud < res).all()
    assert len(blob) < 40 * len(pts)   # a plausible size for untrained weights; the exact rate is checked elsewhere


def test_codec_functions_and_ply_files(tmp_path):
    """compress_point_cloud / decompress_point_cloud (the script bodies as functions) from and to PLY files, adaptive thresholds."""
    from pcc_geo_cnn_v2_b200 import codec, pc_io
    res, level = 128, 1
    rng = np.random.default_rng(9)
    u = rng.random((20000, 2))
    pts = np.stack([u[:, 0] * (res - 1), (0.5 + 0.3 * np.cos(u[:, 1] * 4)) * (res - 1), u[:, 1] * (res - 1)], 1)
    pts = np.unique(pts.astype(np.int64), axis=0).astype(np.float32)
    src = tmp_path / 'in.ply'
    pc_io.write_pc(str(src), pts)
    m = ModelConfigType['c3p'].build(batch_size=4)
    m.set_weights(synthetic.trained_like_weights(m, seed=3, output_bias=-0.45))
    blobs, data = codec.compress_point_cloud(m, pc_io.load_pc(str(src)), res, level, opt_metrics=('d1_mse',), max_deltas=(np.inf,))
    assert len(blobs) == 1 and data[0]['metrics']['d1_psnr'] > 0
    out = codec.decompress_point_cloud(m, blobs[0])
    assert out.dtype == np.float32 and np.array_equal(out, np.asarray(data[0]['blocks_full'], np.float32))
    dst = tmp_path / 'out.ply'
    pc_io.write_pc(str(dst), out)
    assert np.array_equal(pc_io.load_pc(str(dst)), out)


def test_compress_blocks_with_normals_and_d2_metrics():
    """with_normals=True, one D1 and one D2 opt_metric (compress_octree.py --input_normals ... --opt_metrics d1_mse d2_mse): two
    metric groups -> two data lists, each decodable to the encoder's own reconstruction."""
    res, level, bs = 64, 1, 32
    rng = np.random.default_rng(6)
    u = rng.random((6000, 2))
    pts = np.unique(np.stack([u[:, 0] * (res - 1), (0.5 + 0.3 * np.sin(u[:, 0] * 5)) * (res - 1), u[:, 1] * (res - 1)], 1).astype(np.int64), axis=0)
    nrm = rng.normal(size=(len(pts), 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    cloud = np.concatenate([pts.astype(np.float64), nrm], axis=1)
    blocks, binstr = OC.partition_octree(cloud, [0, 0, 0], [res] * 3, level)
    m = ModelConfigType['c3p'].build(batch_size=4)
    m.set_weights(synthetic.trained_like_weights(m, seed=3, output_bias=-0.45))
    m.compress((1, 1, bs, bs, bs))
    data_list, metadata, _ = m.compress_blocks(None, blocks, binstr, cloud, res, level, with_normals=True,
                                               opt_metrics=('d1_mse', 'd2_mse'), max_deltas=(np.inf,), fixed_threshold=False)
    assert len(data_list) == 2 and [md['idx'] for md in metadata] == [0, 1]
    assert 'd2_psnr' in metadata[1]['metrics'] and np.isfinite(metadata[1]['metrics']['d2_psnr'])
    m.decompress()
    for dl, md in zip(data_list, metadata):
        dec, _ = m.decompress_blocks(None, dl, (bs, bs, bs))
        assert all(np.array_equal(a, b) for a, b in zip(md['x_hat_list'], dec))
