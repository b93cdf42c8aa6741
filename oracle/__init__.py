"""CPU oracle for the pcc_geo_cnn_v2 hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (torch fp32/fp64 + numpy + pure Python), the arithmetic
of the reference's per-block transforms and entropy models:

  * oracle/transforms.py  <- reference src/model_transforms.py:7-169  (Keras Conv3D / Conv3DTranspose,
                             'same' padding, residual blocks, the 8 transform classes)
  * oracle/entropy.py     <- tensorflow-compression==1.3 EntropyBottleneck / GaussianConditional
                             (requirements.txt:7; un-vendored dependency, algorithm restated from its
                             published source as recalled + src/utils/patch_gaussian_conditional.py:49-125)
  * oracle/range_coder.py <- tfc 1.3 range_coding_ops (pmf_to_quantized_cdf, unbounded_index_range_*)
  * oracle/model.py       <- reference src/model_types.py:42-46,108-125,179-411, src/utils/focal_loss.py:5-12,
                             src/model_configs.py:16-42

PARITY UNPINNED: the reference ships no golden vectors, checkpoints or bitstreams for this path
(its only test of the path, src/test_model_transforms.py, checks output SHAPES), and TensorFlow 1.15 /
tensorflow-compression 1.3 are not installable here, so the oracle cannot be validated against the
reference's own numbers.  What pins it instead: the reference's shape tests (ported), closed-form
identities (tests/test_oracle_*.py) and hand-computed known-answer cases under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker / CPU baseline.  The product (pcc_geo_cnn_v2_b200) never imports it.
"""
