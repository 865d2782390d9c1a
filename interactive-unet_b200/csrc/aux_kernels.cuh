// Non-tensor-core kernels of the volume-prediction path (HBM-bound byte/element work).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace iu {

// K1: gather `count` slices [start, start+count) of a cubic volume along `axis` into contiguous
// fp32 images out[b][row][col]; uint8 input is normalised with a true IEEE division by 255
// (`predict.py:237` / `:30`), float input is copied.  Image (row,col) = (y,x) | (z,x) | (z,y).
cudaError_t launch_gather_slices(const void* vol, int vol_is_f32, int n, int axis, int start, int count, float* out,
                                 cudaStream_t stream);
// The same for any slice source given by element strides: element (slice b, row r, col c) of the `count` h x w
// images is base[b*s_slice + r*s_row + c*s_col] (uint8 or fp32).  Vector path when rows are contiguous (s_col == 1,
// 16-byte aligned), smem transpose when slices are (s_slice == 1), scalar gather otherwise.
cudaError_t launch_gather_strided(const void* base, int is_f32, int count, int h, int w, long long s_slice,
                                  long long s_row, long long s_col, float* out, cudaStream_t stream);

// 3x3 stride-2 pad-1 max pool on NHWC 16-bit NON-NEGATIVE values (post-ReLU): for those the unsigned
// integer order of the bit patterns equals the numeric order in both fp16 and bf16.
cudaError_t launch_maxpool(const __nv_bfloat16* in, int batch, int h, int w, int c, __nv_bfloat16* out,
                           cudaStream_t stream);

// K4: fused cross-axis accumulate + average + (Gaussian window blend) + uint8 quantise + argmax.
//   p[a] : per-axis probabilities, fp32, slice-major, or nullptr if axis a is not used:
//          p[0][((z*n + y)*n + x)*C + c],  p[1][((y*t + z)*n + x)*C + c],  p[2][((x*t + z)*n + y)*C + c]
//          with z in [0,t) local to the slab, y,x in [0,n)
//   order: the axes in accumulation order (`predict.py:87`), n_axes of them
//   g1d  : the 1-D Gaussian factor (n floats, device) or nullptr for "no window";
//          window = clip((g[z0+z]*g[y])*g[x] / gmax, lo, 1)   (`predict.py:327-347`)
//   outputs (any may be nullptr): uint8 probs [t][n][n][C], uint8 labels [t][n][n], fp32 mean [t][n][n][C]
struct ReduceArgs {
  const float* p[3];
  int order[3];
  int n_axes;
  int n, t, z0, num_classes;
  int zoff, zcount;     // planes [zoff, zoff + zcount) of the slab are reduced (zcount == 0: all t planes)
  const float* g1d;
  float gmax, lo;
  uint8_t* out_u8;
  uint8_t* out_labels;
  float* out_mean;
  // tiled / blended mode (`predict.py:235-245`): when blend_pred != nullptr the block's voxels inside the local crop
  // [l0, l1) are accumulated into the whole-volume buffers
  //   pred[g][c] += mean[c] * window,  weight[g] += window,   g = (block origin b0) + (z, y, x)
  // (fp32, one thread per voxel, blocks processed one after the other on the stream: the reference's add order)
  float* blend_pred;    // [gd][gh][gw][C]
  float* blend_weight;  // [gd][gh][gw]
  int gd, gh, gw;
  int gd_ring;          // > 0: the accumulators hold only gd_ring z planes, plane z of the volume lives at z % gd_ring
                        // (the tiled mode streams finished z ranges out instead of holding the whole volume in fp32)
  int b0[3];            // volume coordinates of the block's voxel (0,0,0); may be negative (padding)
  int l0[3], l1[3];     // local crop of the block that lies inside the volume
};
cudaError_t launch_reduce(const ReduceArgs& args, cudaStream_t stream);

// `get_padded_block` (`predict.py:291-316`): out[s][s][s] = the box [b0, b0+s)^3 of the uint8 volume [d][h][w], clipped
// to the volume and padded back to s^3 with numpy's 'reflect' mode (mirror without repeating the edge voxel).
cudaError_t launch_extract_block(const uint8_t* vol, int d, int h, int w, int i0, int j0, int k0, int s, uint8_t* out,
                                 cudaStream_t stream);

// `normalize_shard` (`predict.py:252-255`): out = uint8(trunc(255 * pred / max(weight, 1e-3))) per voxel and class;
// optional labels = argmax_c pred (first maximum wins).
cudaError_t launch_finalise(const float* pred, const float* weight, size_t voxels, int num_classes, uint8_t* out_u8,
                            uint8_t* out_labels, cudaStream_t stream);

// ---- Zarr staging: the byte shuffling either side of the store, on the device.
// A volume is C-order [d][h][w] voxels of `elem` bytes (classes x item size); the store keeps it as inner chunks of
// cz x cy x cx voxels.  `staged` is chunk-major: chunk (gz, gy, gx) of the ceil(d/cz) x ceil(h/cy) x ceil(w/cx) grid,
// C order, each chunk a contiguous [cz][cy][cx] block of voxels.  to_chunks zero-fills the padding of edge chunks
// (the arrays' fill value); from_chunks ignores it.
// `celem` >= `elem`: bytes per voxel inside a chunk when the chunk's class extent exceeds the array's (padding zeroed).
cudaError_t launch_chunk_layout(const uint8_t* src, uint8_t* dst, int d, int h, int w, int elem, int celem, int cz, int cy,
                                int cx, bool to_chunks, cudaStream_t stream);

// `scipy.ndimage.zoom(block, scale, order=0)` applied block by block as `utils.resize_volume` does
// (`utils.py:29-48`), collapsed into one separable gather: dst[z][y][x][c] = src[tz[z]][ty[y]][tx[x]][tc[c]], or 0
// where any table entry is -1 (scipy's `mode='constant'` fill for a coordinate that rounds past the last sample).
// Items are `item` bytes (1, 2, 4 or 8); the tables live on the device.
cudaError_t launch_zoom_gather(const uint8_t* src, const int* sdims, uint8_t* dst, const int* ddims, const int* tz,
                               const int* ty, const int* tx, const int* tc, int item, cudaStream_t stream);

}  // namespace iu
