/*
 * pccgeo.h -- C ABI of libpccgeo.so, the B200 (sm_100a) implementation of the pcc_geo_cnn_v2 hot path.
 *
 * The reference (mauriceqch/pcc_geo_cnn_v2) has no FFI for this path: it reaches the arithmetic through
 * Python call signatures (Keras layers, tfc entropy models) that bottom out in TensorFlow 1.15 / cuDNN and
 * tensorflow-compression 1.3 kernels.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference repo).  The Python host shim (pcc_geo_cnn_v2_b200/) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all `const T* x` / `T* y` below are DEVICE pointers unless the name
 *     ends in `_host`; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - the caller owns every buffer; the library never allocates device memory behind the caller's back
 *     (workspaces are passed in) and never synchronises the device.
 *   - return value: 0 on success, negative PCCGEO_E* on failure; pccgeo_last_error() gives the text.
 *   - re-entrant per stream.
 *   - tensors are channels_first, contiguous: (N, C, D, H, W) -- the reference's only working layout
 *     (src/model_types.py:180).
 */
#ifndef PCCGEO_H_
#define PCCGEO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCCGEO_OK 0
#define PCCGEO_EINVAL (-1)   /* bad argument / unsupported shape */
#define PCCGEO_ECUDA (-2)    /* CUDA runtime error (text in pccgeo_last_error) */
#define PCCGEO_ENOSPC (-3)   /* output buffer / workspace too small */
#define PCCGEO_EDATA (-4)    /* corrupt bitstream */

/* ---- library ------------------------------------------------------------------------------------ */
const char* pccgeo_last_error(void);
int pccgeo_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
long long pccgeo_launch_count(void);
/* debug / tuning knobs ("umma_swap_lbo_sbo", "umma_max_ctas"); unknown names return PCCGEO_EINVAL */
int pccgeo_set_option(const char* name, long long value);

/* ---- 3D convolutions, fp32 CUDA-core path ---------------------------------------------------------
 * Replaces tf.keras.layers.Conv3D / Conv3DTranspose with padding='same' as instantiated at
 * src/model_transforms.py:45-47,56-58,67-69,78-80,93,107,121,135,144-146,155-157 (TF Conv3D /
 * Conv3DBackpropInputV2 + BiasAdd + Relu), plus the ResidualLayer add (src/model_transforms.py:35-36).
 *
 *   y = [residual +] [relu]( conv(x, w) [+ bias] )
 *
 * x (N,Cin,D,H,W) fp32; y (N,Cout,D',H',W'): D' = ceil(D/stride) (conv) or D*stride (transposed).
 * w is in the library's tap-major layout (k^3, Cin, Cout) -- i.e. the Keras Conv3D kernel as is, and the
 * Keras Conv3DTranspose kernel (k,k,k,Cout,Cin) with its last two axes swapped.
 * bias (Cout) or NULL; residual (shape of y) or NULL. */
int pccgeo_conv3d_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                      int n, int cin, int d, int h, int wd, int cout, int k, int stride, int transposed,
                      int relu, void* stream);

/* ---- 3D convolutions, tcgen05 tensor-core path (3x3x3 kernels) ------------------------------------
 * Same reference call sites as above, for the 3x3x3 layers of the V2 / progressive / hyper transforms
 * (src/model_transforms.py:62-158).  Activations live in a blocked channels layout
 *   (N, C/8, D, H, W, 8) bf16   [one "plane set" per precision term]
 * with `terms` = 1 (bf16 operands) or 2 (hi + lo bf16 split; products a_hi*w_hi + a_hi*w_lo + a_lo*w_hi
 * accumulated in fp32 -- "bf16x3", fp32-class accuracy).  Term t of a tensor starts at element offset
 * t * N*C*D*H*W.  Channel counts are padded up to a multiple of 16 by the pack/convert kernels. */

/* fp32 (N,C,D,H,W) -> blocked bf16 (terms, N, Cp/8, D, H, W, 8), Cp = round_up(C,16), zero padded */
int pccgeo_f32_to_blocked(const float* x, void* xb, int n, int c, int d, int h, int wd, int terms, void* stream);
/* The eight stride-2 phase volumes x[.., 2b + p] of x (N, Cb, 2d, 2h, 2wd) stacked as channels, cut into 8*Cb/C chunks of C channels and
 * written in the blocked layout, chunk after chunk (each chunk = one pccgeo_f32_to_blocked image of shape (N, C, d, h, wd)): what the
 * stride-2 layers' weight gradient feeds to pccgeo_conv3d_wgrad_umma (reference src/model_types.py:364-369, the stride-2 Conv3D /
 * Conv3DTranspose kernels of AnalysisBlock / SynthesisBlock, src/model_transforms.py:62-81). */
int pccgeo_f32_phases_to_blocked(const float* x, void* xb, int n, int cb, int c, int d, int h, int wd, int terms, void* stream);
/* blocked bf16 -> fp32 (N,C,D,H,W) (sums the terms) */
int pccgeo_blocked_to_f32(const void* xb, float* x, int n, int c, int d, int h, int wd, int terms, void* stream);

/* First layer of the V2 / progressive analysis transforms -- Conv3D(F <= 16, (3,3,3), strides 2, 'same') + BiasAdd + Relu on the
 * one-channel occupancy volume (src/model_transforms.py:67) -- writing the blocked bf16 layout directly (no fp32 intermediate).
 * x: fp32 (N,1,D,H,W), even dims; w: tap-major fp32 (27, 1, cout); yb: blocked (terms, N, 2, D/2, H/2, W/2, 8). */
int pccgeo_conv3d_first(const float* x, const float* w, const float* bias, void* yb, int n, int d, int h, int wd, int cout,
                        int relu, int terms, void* stream);

/* Pack tap-major fp32 weights (27, Cin, Cout) into the UMMA B-operand image used by pccgeo_conv3d_umma.
 * Returns the image size in bytes when wpacked == NULL.  `transposed`, `stride` select the layer type
 * (they change which taps are stacked together); HOST pointers in, HOST image out. */
long long pccgeo_umma_pack_weights_host(const float* w_host, void* wpacked_host, int cin, int cout,
                                        int stride, int transposed, int terms);

/* y = [residual +] [relu](conv(x,w) [+bias]) on blocked tensors, stride 1 or 2, conv or transposed conv.
 * bias: fp32 (Cout) or NULL.  residual: blocked like y or NULL.  Spatial dims must be multiples of 8
 * (16 for H when stride==1). */
int pccgeo_conv3d_umma(const void* xb, const void* wpacked, const float* bias, const void* residual_b,
                       void* yb, int n, int cin, int d, int h, int wd, int cout, int stride, int transposed,
                       int relu, int terms, void* stream);

/* hi/lo-stacked variant of pccgeo_conv3d_umma for the stride-1 3x3x3 layers with <= 16 input and output channels in two-term
 * (bf16x3-class) precision -- the second and third layer of AnalysisBlock / SynthesisBlock at 16 filters
 * (src/model_transforms.py:62-81), the dominant layers of the c3p synthesis transform: the bf16 hi and lo halves of the weights
 * are stacked in the MMA N dimension next to the three z-taps (N = 96), so each tap costs two MMAs (a_hi, a_lo) instead of
 * three; all four partial products are accumulated.  Same tensors as pccgeo_conv3d_umma(terms = 2); own weight image. */
long long pccgeo_umma_hl_pack_weights_host(const float* w_host, void* wpacked_host, int cin, int cout, int transposed);
int pccgeo_conv3d_umma_hl(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                          int n, int cin, int d, int h, int wd, int cout, int transposed, int relu, void* stream);

/* zy-ring form of pccgeo_conv3d_umma for the same layers (stride-1 3x3x3, <= 16 channels in and out; conv or transposed conv):
 * the M tile is 16 blocks x 8 x-voxels of one input row and BOTH the y and the z taps are accumulated by the tensor core in a
 * two-dimensional TMEM ring, so one MMA of N = 144 per x-offset and precision product replaces nine of N = 48 (a third of the
 * A-operand fetches per MAC).  Same tensors as pccgeo_conv3d_umma; own weight image.  W % 8 == 0; efficient for N % 16 == 0. */
long long pccgeo_umma_zy_pack_weights_host(const float* w_host, void* wpacked_host, int cin, int cout, int transposed, int terms);
int pccgeo_conv3d_umma_zy(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                          int n, int cin, int d, int h, int wd, int cout, int relu, int terms, void* stream);

/* y-stacked variant of pccgeo_conv3d_umma for stride-1 3x3x3 layers with <= 16 input and output channels (the second and
 * third layer of AnalysisBlock / SynthesisBlock at 16 filters, src/model_transforms.py:62-81): z AND y taps are stacked in
 * the MMA N dimension (3 MMAs of N=144 per input plane and precision pair instead of 9 of N=48); the epilogue adds the
 * three row partials.  Own weight image (pccgeo_umma_ys_pack_weights_host).  W % 8 == 0, any H. */
long long pccgeo_umma_ys_pack_weights_host(const float* w_host, void* wpacked_host, int cin, int cout, int transposed, int terms);
int pccgeo_conv3d_umma_ys(const void* xb, const void* wpacked, const float* bias, const void* residual_b, void* yb,
                          int n, int cin, int d, int h, int wd, int cout, int relu, int terms, void* stream);

/* Last layer of the V2 synthesis transforms, Conv3DTranspose(1, (3,3,3), 'same') + BiasAdd + Relu
 * (src/model_transforms.py:107,135), fused with what compress_blocks / decompress_blocks do with x_hat: clip to [0,1],
 * compare with the block's threshold, pack (src/model_types.py:201-202,209,233-234).  Scatter-form tcgen05 kernel: the 27
 * taps are the MMA N dimension, the partial sums are gathered per output voxel in a fixed order.
 * w: tap-major fp32 (27, Cin, 1), Cin <= 16.  Outputs (either may be NULL, not both): x_hat fp32 (N,1,D,H,W) after ReLU
 * (unclipped, as the reference's graph returns it); bits = packed occupancy (N, D*H*W/32) of min(x_hat,1) > thresholds[n]
 * in argwhere (C) order, bit i of word j = voxel 32*j+i, plus optional per-block popcounts.  H % 16 == 0, W % 8 == 0. */
long long pccgeo_out1_pack_weights_host(const float* w_host, void* wpacked_host, int cin, int transposed, int terms);
int pccgeo_conv3d_out1(const void* xb, const void* wpacked, const float* bias, float* x_hat, uint32_t* bits,
                       const float* thresholds, int32_t* counts, int n, int cin, int d, int h, int wd, int relu,
                       int terms, void* stream);

/* General tensor-core path (gather-im2col -> tcgen05): any cubic kernel <= 9, stride 1 or 2, conv or transposed conv,
 * 16..64 channels, any volume size (the batch is folded into the GEMM rows).  Same reference call sites as above, plus
 * the 5^3 / 9^3 layers of AnalysisTransformV1 / SynthesisTransformV1 (src/model_transforms.py:41-59).
 * pccgeo_gemm_pack_weights_host: tap-major fp32 (k^3, Cin, Cout) HOST weights -> self-describing image (tap table +
 * per-tap bf16 B-operand chunks); returns the size when out == NULL.  The caller uploads the image and also passes its
 * first 128 bytes from HOST memory (`wimg_header_host`) so that the launch needs no device read-back. */
long long pccgeo_gemm_pack_weights_host(const float* w_host, void* out_host, int cin, int cout, int k, int stride,
                                        int transposed, int terms);
int pccgeo_conv3d_gemm(const void* xb, const void* wimg_dev, const void* wimg_header_host, const float* bias,
                       const void* residual_b, void* yb, int n, int cin, int d, int h, int wd, int cout, int relu,
                       void* stream);

/* ---- entropy models ------------------------------------------------------------------------------
 * tfc.EntropyBottleneck (factorized prior; used at src/model_types.py:254,258,287,291-292,300,306,333,338,
 * 377,382-383,397,404).  `eb_params` is the packed per-channel parameter block, C x 58 floats:
 *   [0:3]  softplus(matrix_0) (3x1)   [3:12] softplus(matrix_1) (3x3 row-major)   [12:21] softplus(matrix_2)
 *   [21:24] softplus(matrix_3) (1x3)  [24:27] bias_0  [27:30] bias_1  [30:33] bias_2  [33] bias_3
 *   [34:37] tanh(factor_0) [37:40] tanh(factor_1) [40:43] tanh(factor_2)  [43] median  [44:58] reserved */
#define PCCGEO_EB_PARAM_STRIDE 58

/* symbols = floor(x + .5 - median[c]) (int32), x_hat = symbols + median[c]; either output may be NULL */
int pccgeo_eb_quantize(const float* x, const float* eb_params, int32_t* symbols, float* x_hat,
                       int n, int c, int spatial, void* stream);
/* x_hat = symbols + median[c] (decoder side of EntropyBottleneck.decompress) */
int pccgeo_eb_dequantize(const int32_t* symbols, const float* eb_params, float* x_hat,
                         int n, int c, int spatial, void* stream);
/* likelihood = max(|sigmoid(s*u) - sigmoid(s*l)|, 1e-9) of `values` (already noised or dequantised).
 * sum_log (device double[1], may be NULL) receives sum(ln likelihood), reduced in a fixed order
 * (deterministic); partials: device workspace of >= pccgeo_reduce_ws_doubles() doubles. */
int pccgeo_eb_likelihood(const float* values, const float* eb_params, float* likelihood, double* sum_log,
                         double* partials, int n, int c, int spatial, void* stream);

/* tfc.GaussianConditional + the reference patch (src/utils/patch_gaussian_conditional.py:49-125; used at
 * src/model_types.py:340-341,385-387,406-407).  scale_table: device fp32 (levels).
 * symbols = rint(y) (int32), y_hat = float(symbols), indexes = (levels-1) - #{t in table[:-1]: max(sigma,table[0]) <= t}
 * any of y/symbols/y_hat may be NULL (decoder calls it with only sigma -> indexes). */
int pccgeo_gc_quantize(const float* y, const float* sigma, const float* scale_table, int levels,
                       int32_t* symbols, float* y_hat, int32_t* indexes, long long count, void* stream);
/* likelihood = max(Phi((.5-|v|)/s) - Phi((-.5-|v|)/s), 1e-9), s = max(sigma, scale_min) */
int pccgeo_gc_likelihood(const float* values, const float* sigma, float scale_min, float* likelihood,
                         double* sum_log, double* partials, long long count, void* stream);
/* int32 symbols -> fp32 (GaussianConditional.decompress dequantise, mean=None) */
int pccgeo_i32_to_f32(const int32_t* symbols, float* out, long long count, void* stream);
size_t pccgeo_reduce_ws_doubles(void);

/* ---- voxel-grid helpers ---------------------------------------------------------------------------
 * sparse_to_dense (src/model_types.py:108-114): scatter 1.0f at integer coords.  coords: int16 (npts,4) =
 * (block, z, y, x); x must be zero-initialised by the caller (cudaMemsetAsync). */
int pccgeo_densify(const int16_t* coords, long long npts, float* x, int n, int d, int h, int wd, void* stream);
/* same, for rows that carry GLOBAL block indexes (the output of pccgeo_octree_partition_scatter): rows of blocks
 * [block0, block0 + n) are written, the others skipped */
int pccgeo_densify_from(const int16_t* coords, long long npts, int block0, float* x, int n, int d, int h, int wd, void* stream);
/* decompress_blocks / compress_blocks thresholding (src/model_types.py:201-202,209,233-234):
 * bit = min(x_hat,1) > threshold[block]; packed little-endian, 32 voxels per uint32 word in C order;
 * counts[block] = number of set bits.  voxels per block must be a multiple of 32. */
int pccgeo_threshold_pack(const float* x_hat, const float* thresholds, uint32_t* bits, int32_t* counts,
                          int n, long long voxels_per_block, void* stream);
/* focal_loss (src/utils/focal_loss.py:5-12), sum-reduced, deterministic order; out: device double[1] */
int pccgeo_focal_loss(const float* x_true, const float* x_pred, float gamma, float alpha, double* out,
                      double* partials, long long count, void* stream);

/* ---- training path (tr_train.py; reference src/model_types.py:327-369, TF autodiff + two AdamOptimizers) ---------
 * fp32, deterministic.  Data gradients of a conv are pccgeo_conv3d_f32 with conv <-> transposed conv swapped and the
 * tap-major weights' last two axes swapped; the entry points below add what has no forward counterpart. */
/* dx = dy where y > 0 else 0 (Relu backward; y = the layer's post-activation output) */
int pccgeo_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream);
/* out = alpha*a + beta*b (b may be NULL): ResidualLayer add (src/model_transforms.py:35) and gradient accumulation */
int pccgeo_axpby(const float* a, const float* b, float alpha, float beta, float* out, long long n, void* stream);
/* d(scale * focal_loss)/d y_pred (src/utils/focal_loss.py:5-12; clip passes gradient on [1e-3,.999] only) */
int pccgeo_focal_loss_bwd(const float* x_true, const float* x_pred, float gamma, float alpha, float scale, float* dx_pred,
                          long long n, void* stream);
/* d(c * sum ln p)/d values and /d sigma of the Gaussian conditional, tfc lower_bound gradient semantics */
int pccgeo_gc_likelihood_bwd(const float* values, const float* sigma, float scale_min, float c, float* dvalues,
                             float* dsigma, long long n, void* stream);
/* d(c * sum ln p)/d values of the entropy bottleneck + per-channel gradients w.r.t. the first 44 entries of the packed
 * parameter block (softplus'ed matrices, biases, tanh'ed factors): dparams (C,44).  ws: pccgeo_eb_bwd_ws_doubles(C). */
size_t pccgeo_eb_bwd_ws_doubles(int c);
int pccgeo_eb_likelihood_bwd(const float* values, const float* eb_params, float c, float* dvalues, float* dparams,
                             double* ws, int n, int ch, int spatial, void* stream);
/* weight gradient of a 'same' conv / transposed conv in the tap-major layout (k^3, Cin, Cout) of pccgeo_conv3d_f32.
 * x: layer input (N,Cin,D,H,W); g: gradient w.r.t. the pre-activation output.  ws: pccgeo_wgrad_ws_floats(). */
size_t pccgeo_wgrad_ws_floats(int cin, int cout, int k);
int pccgeo_conv3d_wgrad_f32(const float* x, const float* g, float* dw, float* ws, int n, int cin, int d, int h, int wd,
                            int cout, int k, int stride, int transposed, void* stream);
/* The same weight gradient on the tensor cores (tcgen05, bf16 or bf16x3 operands, fp32 accumulation in TMEM) for the 3x3x3 stride-1
 * layers with c channels in and out, c in {16, 32, 64}, W in {16, 32, 64} (and 8 for c >= 32) -- the backward-filter pass of the Conv3D /
 * Conv3DTranspose layers that tf.train.AdamOptimizer.minimize differentiates (reference src/model_types.py:364-369).
 * xb, gb: layer input and pre-activation output gradient in the blocked bf16 layout (pccgeo_f32_to_blocked, `terms` terms).
 * pccgeo_wgrad_umma_ws_floats() returns 0 when the geometry is not supported (use pccgeo_conv3d_wgrad_f32). */
long long pccgeo_wgrad_umma_ws_floats(int c, int n, int d, int h, int wd, int terms);
int pccgeo_conv3d_wgrad_umma(const void* xb, const void* gb, float* dw, float* ws, int n, int c, int d, int h, int wd,
                             int transposed, int terms, void* stream);
/* The elementwise steps of the backward pass on the blocked bf16 layout (pccgeo_f32_to_blocked; `groups` = N * ceil16(C)/8 * voxels
 * 16-byte items per term): ReLU mask from the activation's hi term (tf.nn.relu's gradient), residual add (ResidualLayer 'add',
 * reference src/model_transforms.py:35-36), bias gradient (BiasAdd; ws: pccgeo_bias_grad_blocked_ws_doubles(c) doubles). */
size_t pccgeo_bias_grad_blocked_ws_doubles(int c);
int pccgeo_relu_mask_blocked(const void* gb, const void* yb, void* out, long long groups, int terms, void* stream);
int pccgeo_add_blocked(const void* a, const void* b, void* out, long long groups, int terms, void* stream);
int pccgeo_bias_grad_blocked(const void* gb, float* db, double* ws, int n, int c, long long spatial, int terms, void* stream);
/* db[c] = sum over batch and voxels of g; ws: pccgeo_reduce_ws_doubles() doubles */
int pccgeo_bias_grad_f32(const float* g, float* db, double* ws, int n, int c, long long spatial, void* stream);
/* tf.train.AdamOptimizer step t (1-based): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps) */
int pccgeo_adam_step(float* theta, const float* grad, float* m, float* v, float lr, float beta1, float beta2, float eps,
                     long long step, long long n, void* stream);

/* ---- range coder (HOST) ---------------------------------------------------------------------------
 * Replaces tfc's range_coding_ops.unbounded_index_range_encode / _decode (precision 16, overflow_width 4;
 * src/utils/patch_gaussian_conditional.py:27-31 and EntropyBottleneck/GaussianConditional.compress/.decompress).
 * All pointers are HOST memory.  Streams are independent; `threads` worker threads split them.
 *   symbols/indexes: concatenated per-stream arrays, stream i covers [sym_offsets[i], sym_offsets[i+1])
 *   cdf: (rows, cdf_stride) int32, cdf_length (rows), offset (rows)
 *   encode: out_bytes capacity out_capacity; out_offsets (nstreams+1) receives the byte ranges.
 *   index_mode 0: indexes given per symbol; 1: index = (position / channel_stride) % rows (per-channel
 *   tables, EntropyBottleneck) -- `indexes` may then be NULL. */
int pccgeo_range_encode_host(const int32_t* symbols, const int32_t* indexes, const long long* sym_offsets,
                             int nstreams, const int32_t* cdf, int cdf_stride, const int32_t* cdf_length,
                             const int32_t* offset, int rows, int index_mode, long long channel_stride,
                             uint8_t* out_bytes, long long out_capacity, long long* out_offsets, int threads);
int pccgeo_range_decode_host(const uint8_t* bytes, const long long* byte_offsets, const int32_t* indexes,
                             const long long* sym_offsets, int nstreams, const int32_t* cdf, int cdf_stride,
                             const int32_t* cdf_length, const int32_t* offset, int rows, int index_mode,
                             long long channel_stride, int32_t* symbols_out, int threads);
/* ---- range coder (DEVICE) -------------------------------------------------------------------------
 * The same coder as above, byte for byte, run on the GPU with one warp per stream so that the block loops of
 * compress_blocks / decompress_blocks (src/model_types.py:184-238) need no host entropy coding.  All pointers are DEVICE
 * memory; every stream has per_stream symbols (the codec's streams are the latents of equal-sized blocks).  Tables as
 * above, uploaded by the caller; `err` (device int, caller-zeroed) is set to 1 on an out-of-range table index or a
 * corrupt escape code.
 *   encode: ws of pccgeo_rc_encode_ws_bytes(); streams are packed back to back into `packed` (bytes beyond
 *           packed_capacity are dropped: compare offsets[nstreams] with the capacity), lengths (nstreams) and offsets
 *           (nstreams+1) receive the byte ranges; lengths[i] == -1 if stream i outgrew 4 bytes per symbol.
 *   decode: byte_offsets (nstreams+1) into `bytes`; cdf16 / row_start from pccgeo_range_compact_tables_host, uploaded. */
size_t pccgeo_rc_encode_ws_bytes(int nstreams, long long per_stream);
int pccgeo_range_encode_device(const int32_t* symbols, const int32_t* indexes, int nstreams, long long per_stream,
                               const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, const int32_t* offset,
                               int rows, int index_mode, long long channel_stride, void* ws, uint8_t* packed,
                               long long packed_capacity, int32_t* lengths, long long* offsets, int* err, void* stream);
int pccgeo_range_decode_device(const uint8_t* bytes, const long long* byte_offsets, const int32_t* indexes, int nstreams,
                               long long per_stream, const uint16_t* cdf16, const int32_t* row_start, const int32_t* cdf_length,
                               const int32_t* offset, int rows, int total_entries, int index_mode, long long channel_stride,
                               int32_t* symbols_out, int* err, void* stream);
/* HOST: the decoder's compact tables (it keeps them in shared memory): the rows' first cdf_length[r]-1 entries as 16 bits, back
 * to back (a row's last entry, 2^16, is implied), row_start (rows) = the rows' positions.  Returns the number of entries, -1 if
 * a row is not a 16-bit CDF; cdf16 == NULL queries the size. */
long long pccgeo_range_compact_tables_host(const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, int rows,
                                           uint16_t* cdf16, int32_t* row_start);
/* HOST test hook: the device encoder's arithmetic (the same inline functions the kernels call) run sequentially over
 * HOST arrays, so that the CPU test-suite can pin it against pccgeo_range_encode_host without a GPU. */
int pccgeo_range_encode_emulate_host(const int32_t* symbols, const int32_t* indexes, int nstreams, long long per_stream,
                                     const int32_t* cdf, int cdf_stride, const int32_t* cdf_length, const int32_t* offset,
                                     int rows, int index_mode, long long channel_stride, uint8_t* packed,
                                     long long packed_capacity, int32_t* lengths, long long* offsets);
/* HOST test hook: the device encoder's carry out of its 4-byte register into the n bytes already stored in buf. */
unsigned pccgeo_rc_carry_probe_host(uint8_t* buf, int n);
/* HOST: packed occupancy words from pccgeo_threshold_pack (copied to the host) -> float32 (z,y,x) rows in np.argwhere
 * order, the host half of the reference's `np.argwhere(x_hat > t).astype(float32)` (src/model_types.py:209,234).
 * offsets (n_blocks+1) receives the prefix sum of per-block point counts; pass points == NULL to query sizes only. */
int pccgeo_bits_to_points_host(const uint32_t* bits, int n_blocks, int d, int h, int w, long long* offsets,
                               float* points, long long capacity_points, int threads);

/* Host half of sparse_to_dense / pc_to_tf for a batch (src/model_types.py:23-39,108-114): per-block point arrays (rows of
 * >= 3 float32 or float64 coordinates, row pitch row_bytes[b]) -> int16 (sum counts, 4) rows (block, c0, c1, c2), the
 * input of pccgeo_densify.  Values are truncated like numpy's float -> int16 cast. */
int pccgeo_blocks_to_coords_host(const void* const* blocks, const long long* counts, const long long* row_bytes,
                                 int n_blocks, int is_f64, int16_t* out, int threads);

/* partition_octree (src/utils/octree_coding.py:68-113) on the GPU: stable counting sort of the points by the Morton key of their
 * block.  Two asynchronous stages around one small read-back (the number of occupied blocks): see csrc/octree.cu.
 * rows: device float64 (n, cols >= 3), coordinates in [0, block_size * 2^level); out_rows: grouped rows in local coordinates
 * (points of a block keep their input order) and / or out_coords: int16 (block, i0, i1, i2) rows for pccgeo_densify_from;
 * offsets: device int64 (n_blocks + 1). */
size_t pccgeo_octree_ws_bytes(long long n, int level);
size_t pccgeo_octree_ws2_bytes(long long n, int n_blocks);
int pccgeo_octree_partition_keys(const double* rows, long long n, int cols, double block_size, int level, void* ws, void* stream);
int pccgeo_octree_partition_scatter(const double* rows, long long n, int cols, double block_size, int level, int n_blocks,
                                    const void* ws, void* ws2, double* out_rows, int16_t* out_coords, long long* offsets, void* stream);

/* Host half of partition_octree (src/utils/octree_coding.py:103-111): stable counting sort of float64 point rows (n, cols)
 * by block index, block origins (n_blocks, 3) subtracted from the first three columns; offsets = (n_blocks + 1) prefix sums. */
int pccgeo_group_points_host(const double* rows, const int32_t* block_idx, long long n, int cols, int n_blocks,
                             const double* origins, double* out, long long* offsets);

/* ---- per-block threshold optimisation ("next" #1) -----------------------------------------------
 * Replaces the kd-tree loop of src/model_opt.py:9-44 (one compute_metrics, src/utils/pc_metric.py:76-108, per threshold and
 * block): exact integer D1 sums between a block's points A and B_i = { x_hat > thresholds[i] } for every threshold at once.
 * Stage 1 fills the workspace (rank volume + per-slice EDT of A) and the (N, T+1) histograms over rank of EDT2_A and of the
 * voxel counts (exclusive suffix sums over rank give sum_BA[i] and |B_i|); stage 2 computes sum_AB[n][i] (-1: B_i empty,
 * -2: B_i == B_{i-1}).  x_hat fp32 (N,1,D,H,W), H,W <= 128; points int16 (P,4) rows (block,z,y,x) sorted by block;
 * offsets int64 (N+1); all device pointers. */
size_t pccgeo_threshold_opt_ws_bytes(int n, int d, int h, int wd);
int pccgeo_threshold_hist(const float* x_hat, const float* thresholds, int t, const int16_t* points, const long long* offsets,
                          void* ws, unsigned long long* hist, unsigned long long* cnt, int n, int d, int h, int wd, void* stream);
int pccgeo_threshold_sum_ab(const void* ws, const int16_t* points, const long long* offsets, const long long* counts_b,
                            long long* sum_ab, int n, int t, int d, int h, int wd, int max_points, void* stream);
/* tfc pmf_to_quantized_cdf (precision 16): pmf (len) doubles -> cdf (len+1) int32, HOST */
int pccgeo_pmf_to_quantized_cdf_host(const double* pmf, int len, int precision, int32_t* cdf);

#ifdef __cplusplus
}
#endif
#endif /* PCCGEO_H_ */
