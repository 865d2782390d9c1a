/* iunet_b200: C ABI of the B200-native volume-prediction engine.
 *
 * This is the drop-in boundary for the full-volume prediction path of
 * laprade117/interactive-unet.  The reference has no FFI of its own: the seam is the Python
 * callables in `interactive_unet/predict.py` and the `UNet` module in `interactive_unet/unet.py`
 * (SURVEY.md section 8b).  Each entry point below names the reference lines it replaces; the Python
 * host side (`interactive-unet_b200/predict.py`, `unet.py`) keeps the reference's signatures and
 * binds these symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C types only; every function returns IU_OK (0) or an IU_ERR_* code, the message is
 *     available from iu_last_error();
 *   - the caller owns every input / output buffer; the engine owns weights and workspace;
 *   - pointers documented as "host or device" are classified with cudaPointerGetAttributes;
 *   - volumes are C-order [z][y][x]; a slice along axis a is image (y,x) | (z,x) | (z,y);
 *   - all work is issued on the engine's own non-blocking stream; calls return after that stream
 *     has drained unless IU_FLAG_ASYNC is passed (then call iu_engine_synchronize);
 *   - an engine is not re-entrant: serialise calls per handle (the Python side holds a lock);
 *   - there is no CPU fallback: every compute entry fails with IU_ERR_CUDA on a machine
 *     without an sm_100 GPU.
 */
#ifndef IUNET_B200_H_
#define IUNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IU_ABI_VERSION 2

#define IU_OK 0
#define IU_ERR_INVALID 1 /* bad argument / unsupported configuration            */
#define IU_ERR_CUDA 2    /* CUDA runtime / driver failure                        */
#define IU_ERR_OOM 3     /* device allocation failed: message contains "out of memory" (predict.py:67-72) */
#define IU_ERR_STATE 4   /* weights not loaded, etc.                             */

#define IU_FLAG_ASYNC 1u

#define IU_DTYPE_U8 0
#define IU_DTYPE_F32 1

/* 16-bit storage format of weights and inter-layer activations (accumulation is always fp32 in TMEM).
 * Both run the same tcgen05 kind::f16 instruction at the same rate and move the same bytes. */
#define IU_PRECISION_FP16 0 /* default: 11-bit significand, ~8x smaller rounding error than bf16 */
#define IU_PRECISION_BF16 1

typedef struct iu_engine iu_engine;

int iu_abi_version(void);

/* Message of the last failure on `e` (or of the last failed iu_engine_create when e == NULL). */
const char* iu_last_error(const iu_engine* e);

/* Create / destroy an engine bound to CUDA device `device`. */
int iu_engine_create(int device, iu_engine** out);
void iu_engine_destroy(iu_engine* e);

/* Stream handle (cudaStream_t) the engine launches on, for callers that record their own events. */
void* iu_engine_stream(iu_engine* e);
int iu_engine_synchronize(iu_engine* e);

/* Load the weights of smp.Unet('resnet34', in_channels=1, classes=num_classes) -- the model
 * `unet.py:56-61` builds for architecture='U-Net', encoder_name='resnet34' -- from host fp32
 * tensors named with smp's state_dict keys WITHOUT the Lightning `model.` prefix
 * (e.g. "encoder.layer1.0.conv1.weight", "decoder.blocks.0.conv1.1.running_var",
 * "segmentation_head.0.bias"; `num_batches_tracked` entries may be omitted).  The engine folds every
 * eval-mode BatchNorm into the preceding conv, converts to the 16-bit storage format and packs for the tensor cores.
 * Replaces `UNet.load_from_checkpoint(...).to(device).eval()` (predict.py:22-27,130-135). */
int iu_engine_load_weights(iu_engine* e, int num_classes, int n_tensors, const char* const* names,
                           const float* const* data, const int64_t* numel);
int iu_engine_num_classes(const iu_engine* e);

/* Select the storage format; call BEFORE iu_engine_load_weights (changing it drops loaded weights). */
int iu_engine_set_precision(iu_engine* e, int precision);
int iu_engine_precision(const iu_engine* e);

/* Upper bound on the slices per internal batch (0 = automatic).  Results do not depend on it. */
int iu_engine_set_max_batch(iu_engine* e, int max_batch);
/* Slices the engine runs per network pass when asked for `count` slices of h x w (the automatic choice, capped by
 * iu_engine_set_max_batch): callers that pipeline work against the engine (the multi-GPU exchange) chunk by it. */
int iu_engine_auto_batch(const iu_engine* e, int h, int w, int count);
/* Device bytes the engine would hold for `batch` slices of h x w (weights + workspace). */
int64_t iu_engine_workspace_bytes(iu_engine* e, int batch, int h, int w);

/* `UNet.forward` (unet.py:65-69): x fp32 [batch,1,h,w] -> softmax probabilities fp32
 * [batch,num_classes,h,w] (NCHW).  h, w multiples of 32.  x / probs: host or device. */
int iu_engine_forward(iu_engine* e, const float* x, int batch, int h, int w, float* probs, unsigned flags);

/* One axis of `predict_block` (predict.py:87-108) on a cubic volume of edge n (uint8 or fp32, host
 * or device): runs slices [slice_begin, slice_begin+slice_count) along `axis` through the network
 * and stores their softmax probabilities (fp32, device) slice-major:
 *   probs[ (((row / row_block) * slice_total + slice_offset + i) * row_block + row % row_block) * n + col ][c]
 * for slice i (relative to slice_begin), image pixel (row, col).  With row_block == n and
 * slice_offset == 0 this is simply probs[i][row][col][c]; row_block = n / G lays the buffer out
 * destination-major for the z-slab all-to-all. */
int iu_engine_predict_axis(iu_engine* e, const void* volume, int dtype, int n, int axis, int slice_begin,
                           int slice_count, float* probs_dev, int slice_offset, int slice_total, int row_block,
                           unsigned flags);

/* The same for slices of ANY strided device source (uint8 or fp32): element (slice i, row r, col c) of the `count`
 * h x w images is base_dev[i*stride_slice + r*stride_row + c*stride_col] (strides in elements; h, w multiples of 32,
 * row_block divides h).  A cubic volume [z][y][x] of edge n is (n*n, n, 1) | (n, n*n, 1) | (1, n*n, n) for axis
 * 0 | 1 | 2; the multi-GPU path reads its z-slab and the two exchanged strips this way
 * (`np.moveaxis` + batch slicing of predict.py:91-97 for a source that is not the whole cube).  Output layout as
 * for iu_engine_predict_axis with n replaced by w. */
int iu_engine_predict_slices(iu_engine* e, const void* base_dev, int dtype, int count, int h, int w,
                             int64_t stride_slice, int64_t stride_row, int64_t stride_col, float* probs_dev,
                             int slice_offset, int slice_total, int row_block, unsigned flags);

/* K1 alone (predict.py:91,95,97,237): gather + normalise slices into fp32 [count][n][n] (device). */
int iu_engine_gather_slices(iu_engine* e, const void* volume_dev, int dtype, int n, int axis, int start, int count,
                            float* out_dev, unsigned flags);

/* K4 alone: cross-axis accumulate in `order`, divide by n_axes (predict.py:101-110), blend with the
 * Gaussian window and quantise (predict.py:244-245,255), argmax (predict.py:38).
 *   p0/p1/p2: per-axis probabilities (device, fp32, layouts as written by iu_engine_predict_axis
 *             with row_block == n for a single slab; see aux_kernels.cuh), NULL when unused;
 *   n: volume edge, t: slab thickness in z (t == n on one GPU), z0: first global z of the slab;
 *   g1d: host pointer to the n-entry 1-D Gaussian factor or NULL (no window: q = trunc(255*mean));
 *   out_u8 [t][n][n][C], out_labels [t][n][n], out_mean fp32 [t][n][n][C]: device, any may be NULL. */
int iu_engine_reduce(iu_engine* e, const float* p0, const float* p1, const float* p2, const int* order, int n_axes,
                     int n, int t, int z0, int num_classes, const float* g1d_host, float gmax, float lo,
                     uint8_t* out_u8, uint8_t* out_labels, float* out_mean, unsigned flags);

/* The same for the planes [zoff, zoff + zcount) of the slab only (outputs stay indexed by the slab's z): lets a caller
 * reduce a z range as soon as its last axis has been predicted and copy it out while the next range is in the network. */
int iu_engine_reduce_planes(iu_engine* e, const float* p0, const float* p1, const float* p2, const int* order, int n_axes,
                            int n, int t, int z0, int zoff, int zcount, int num_classes, const float* g1d_host,
                            float gmax, float lo, uint8_t* out_u8, uint8_t* out_labels, float* out_mean, unsigned flags);

/* Whole single-GPU path (predict.py:79-112 + 244-245,255): volume (uint8 or fp32, host or device,
 * cubic edge n) -> uint8 probabilities [n][n][n][C], uint8 labels [n][n][n], fp32 mean probabilities
 * [n][n][n][C]; each output host or device, any may be NULL.  `axes`: n_axes entries from {0,1,2}. */
int iu_engine_predict_volume(iu_engine* e, const void* volume, int dtype, int n, const int* axes, int n_axes,
                             const float* g1d_host, float gmax, float lo, uint8_t* out_u8, uint8_t* out_labels,
                             float* out_mean, unsigned flags);

/* Tiled / blended mode of `predict_volumes` (predict.py:153,201,235-256): a uint8 volume
 * [d][h][w] of any shape (host or device) is predicted in n_blocks cubic blocks of edge s (a multiple of 32) whose
 * voxel (0,0,0) sits at volume coordinates origins[3*b .. 3*b+2] -- the first three columns of the reference's
 * `padded_block_coords` from `get_block_coordinates` (predict.py:362-411); origins may be negative / overhang.  Per block,
 * in the given order: reflect-padded extraction (`get_padded_block`, predict.py:291-316), `predict_block` over `axes`,
 * `pred += mean * window`, `weight += window` over the part of the block inside the volume (predict.py:244-245); then
 * out_u8[d][h][w][C] = uint8(255 * pred / max(weight, 1e-3)) (predict.py:255) and, optionally, labels = argmax_c pred.
 * g1d: host pointer to the s-entry 1-D factor of `gaussian_3d(s)`, gmax / lo as for iu_engine_reduce.
 * out_u8 / out_labels: host or device, either may be NULL. */
int iu_engine_predict_tiled(iu_engine* e, const uint8_t* volume, int d, int h, int w, int s, int n_blocks,
                            const int* origins, const int* axes, int n_axes, const float* g1d_host, float gmax,
                            float lo, uint8_t* out_u8, uint8_t* out_labels, unsigned flags);

/* The three steps of the tiled mode on their own (device buffers; used by the parity tests):
 *   extract_block: `get_padded_block` (predict.py:291-316) -> uint8 [s][s][s];
 *   blend_block:   per-axis probabilities of one block (layouts of iu_engine_predict_axis, row_block == s) ->
 *                  pred[d][h][w][C] += mean * window, weight[d][h][w] += window (predict.py:244-245), fp32;
 *   finalise:      `normalize_shard` (predict.py:252-255) over `voxels` voxels (+ optional argmax labels). */
int iu_engine_extract_block(iu_engine* e, const uint8_t* volume_dev, int d, int h, int w, int i0, int j0, int k0, int s,
                            uint8_t* out_dev, unsigned flags);
int iu_engine_blend_block(iu_engine* e, const float* p0, const float* p1, const float* p2, const int* order, int n_axes,
                          int s, int num_classes, const float* g1d_host, float gmax, float lo, float* pred_dev,
                          float* weight_dev, int d, int h, int w, const int* origin, unsigned flags);
int iu_engine_finalise(iu_engine* e, const float* pred_dev, const float* weight_dev, int64_t voxels, int num_classes,
                       uint8_t* out_u8_dev, uint8_t* out_labels_dev, unsigned flags);

/* Zarr staging, the steps either side of the path (device buffers).
 *   A volume is C-order [d][h][w] voxels of `elem` bytes (classes x item size).  `staged` is the same data as the
 *   store's inner chunks (predict.py:174-179: chunks (128,128,128,C)), chunk-major: chunk (gz,gy,gx) of the
 *   ceil(d/cz) x ceil(h/cy) x ceil(w/cx) grid in C order, each a contiguous [cz][cy][cx] block of voxels of
 *   `chunk_elem` >= `elem` bytes (pyramid levels keep level 0's chunk shape while their class axis is halved,
 *   utils.py:66-71, so a chunk's class extent can exceed the array's; the excess is padding).
 *   to_chunks:   volume -> staged, edge-chunk padding zero-filled (the arrays' fill value) -- what
 *                `final_predictions[i0:i1, j0:j1, k0:k1] = ...` (predict.py:255) makes zarr do on the host;
 *   from_chunks: staged -> volume, padding ignored -- `zarr.open(f)['0'][...]` (predict.py:167, :299).
 *   zoom_nearest: one pyramid level, `utils.resize_volume` (utils.py:29-48: scipy.ndimage.zoom(block, 0.5, order=0)
 *                per shard-sized block) as one gather dst[i][j][k][l] = src[t0[i]][t1[j]][t2[k]][t3[l]] over 4-D
 *                arrays of `item_bytes`-byte items; tables are HOST int32 arrays of dst_dims[k] entries, -1 = scipy's
 *                constant fill (0).  3-D arrays pass a trailing extent of 1. */
int iu_engine_to_chunks(iu_engine* e, const void* volume_dev, int d, int h, int w, int elem, int chunk_elem, int cz,
                        int cy, int cx, void* staged_dev, unsigned flags);
int iu_engine_from_chunks(iu_engine* e, const void* staged_dev, int d, int h, int w, int elem, int chunk_elem, int cz,
                          int cy, int cx, void* volume_dev, unsigned flags);
int iu_engine_zoom_nearest(iu_engine* e, const void* src_dev, const int* src_dims, void* dst_dev, const int* dst_dims,
                           const int* t0, const int* t1, const int* t2, const int* t3, int item_bytes, unsigned flags);

/* Test hook: one tensor-core convolution exactly as the engine runs it.
 *   src0/src1: device 16-bit (engine precision) NHWC [batch][h_in][w_in][cin]; src1 may be NULL (cin1 = 0);
 *   weight: host fp32 [cout][cin0+cin1][k][k] (PyTorch layout, source 0's channels first),
 *   bias: host fp32 [cout]; residual: device 16-bit NHWC at output geometry or NULL;
 *   out: device 16-bit NHWC [batch][h_out*(1+up2x)][w_out*(1+up2x)][cout];
 *   up2x bit 1: src0 is stored at half resolution and read through a 2x nearest upsample (decoder conv1). */
int iu_engine_conv_test(iu_engine* e, const void* src0, int cin0, const void* src1, int cin1, int batch, int h_in,
                        int w_in, int ksize, int stride, const float* weight, const float* bias, int cout,
                        const void* residual, int relu, int up2x, void* out);

/* Number of kernels the engine has launched since creation (bench.py's `gpu_launches`). */
int64_t iu_engine_launch_count(const iu_engine* e);

/* Device memory the engine keeps between calls: packed weights, the activation plans of the last few (batch, h, w)
 * shapes (cached so that `predict_slice` and `predict_volumes` can alternate without re-allocating) and pooled
 * scratch buffers.  iu_engine_release_workspace drains the stream and frees everything but the weights -- the
 * drop-in `predict_volumes` calls it when it is done (the reference frees its tensors per volume, predict.py:257-259,
 * and the trainer shares the GPU); iu_engine_held_bytes reports the current total. */
int iu_engine_release_workspace(iu_engine* e);
int64_t iu_engine_held_bytes(const iu_engine* e);

/* Per-kernel-class device timing: while enabled every launch is bracketed by CUDA events recorded on
 * the engine's stream.  iu_engine_profile_read drains the stream and returns, per class, the summed
 * kernel milliseconds and launch counts accumulated so far (arrays of IU_PROF_CLASSES entries). */
#define IU_PROF_GATHER 0 /* K1 slice gather + normalise                */
#define IU_PROF_STEM 1   /* 7x7/s2 stem conv + BN + ReLU               */
#define IU_PROF_POOL 2   /* 3x3/s2 max-pool                            */
#define IU_PROF_CONV 3   /* tcgen05 implicit-GEMM convs (incl. head)   */
#define IU_PROF_REDUCE 4 /* K4 accumulate / blend / quantise / argmax  */
#define IU_PROF_CLASSES 5
/* Development aid (engines created with env IU_CONV_DEBUG=1): 16 clock-cycle counters per conv layer written
 * by the halo kernel's warp roles (layout in csrc/conv_tc.cuh, ConvArgs::debug); n_values <= 1024. */
int iu_engine_debug_counters(iu_engine* e, unsigned long long* out, int n_values, int reset);
int iu_engine_profile(iu_engine* e, int enable);
int iu_engine_profile_read(iu_engine* e, double* ms, int64_t* count, int reset);

#ifdef __cplusplus
}
#endif
#endif /* IUNET_B200_H_ */
