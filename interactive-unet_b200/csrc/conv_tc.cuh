// Host-visible description of one tensor-core convolution launch (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace iu {

constexpr int kTileM = 128;  // output pixels per CTA tile = UMMA M

enum EpilogueMode : int {
  kEpiBf16 = 0,         // bias (+residual) (+ReLU) -> 16-bit NHWC (optionally written 2x nearest-upsampled)
  kEpiSoftmaxNHWC = 1,  // bias -> softmax over the first num_classes columns -> fp32 [slice][row][col][C]
  kEpiSoftmaxNCHW = 2,  // bias -> softmax -> fp32 [n][C][row][col]
};

// One K segment of the implicit GEMM: a source activation tensor (NHWC bf16) read through a
// k x k window with the given stride / padding.  A conv has 1 or 2 segments:
//   cat([a, b], dim=1) -> 3x3   = {a, 3, 1, 1} + {b, 3, 1, 1}
//   conv3x3(t) + downsample(x)  = {t, 3, 1, 1} + {x, 1, 2, 0}
struct ConvSegment {
  int cin;     // channels of this source (multiple of the K chunk)
  int ksize;   // 1 or 3
  int stride;  // 1 or 2
  int pad;     // 0 or 1
  int up;      // 1: the source is stored at half the output resolution and read through a 2x nearest upsample
               //    (decoder `F.interpolate(x, scale_factor=2)` fused into the consumer; halo kernel only)
};

struct alignas(64) ConvArgs {
  CUtensorMap amap[2];  // 4-D maps (C, W, H, N) over the segment sources, box = (KC, tw*stride, th*stride, nb)
  CUtensorMap bmap;     // 2-D map (K, Cout_pad) over the packed weights, box = (KC, BN)
  CUtensorMap bmap2;    // the same weights with box = (64, 64): one CTA's half tile in the CTA-pair kernel
  CUtensorMap bmap256;  // the same weights with box = (64, 256): per-tap kernel with BN = 256 (valid when use_bn256)
  int use_bn256;
  // Latency shapes (a handful of pixel tiles): the same weights boxed (64, small_bn) with small_bn = 64 or 32, so that
  // the layer's Cout is spread over 2-8x as many CTAs (each streams that much less of the weights); 0 = not planned.
  CUtensorMap bmap_small;
  int small_bn;
  // row-folded kernel (conv_row.cu), valid when use_row != 0:
  CUtensorMap bmapf;    // 2-D map (3*sum(cin), 3*Cout_pad) over the fold-packed weights, box = (row KC, 3*Cout_pad)
  CUtensorMap bmapu;    // 2-D map (3*cin0, 4*Cout_pad) over the four-slot weights of an upsampled segment 0,
                        // box = (row KC, 4*Cout_pad)
  CUtensorMap bmapi;    // 2-D map over the [Cout][Cout] identity (residual segment), box = (row KC, Cout)
  CUtensorMap omap;     // 4-D map (Cout, W, H, N) over the output tensor, box = (Cout, 128, store rows, 1): TMA store
  int use_row;
  // row-folded kernel, Cout 64 with a residual: 4-D map over the RESIDUAL tensor with the output map's box, and
  // res_tma != 0 to add the residual in the epilogue from a TMA-loaded staging buffer instead of as an identity K segment
  CUtensorMap rmap;
  int res_tma;
  // row-folded kernel, Cin = Cout = 64 identity source: 4-D map (64, W, H, N) over the SOURCE, box (64, 130, 3, 1),
  // 128B swizzle -- the A ring is filled by TMA (zero fill outside the image = the conv's padding) instead of by the
  // gather warps' cp.async; row_tma != 0 selects that variant (2: descriptors carry a base offset, development switch)
  CUtensorMap rowmap;
  int row_tma;
  // halo kernel with TMA-filled tiles (identity sources, 64-channel chunks, Cout % 128 == 0): 4-D maps (C, W, H, N)
  // over the segment sources, box (64, 18, 18, 1), 128B swizzle; halo_tma != 0 selects that variant and is the number
  // of halo pixels per buffer row (the box width: 18, or 24 as a development switch)
  CUtensorMap hmap[2];
  int halo_tma;
  ConvSegment seg[2];
  const __nv_bfloat16* src_ptr[2];  // raw pointers of the segment sources (halo kernel: cp.async gathers)
  int nseg;
  int batch, out_h, out_w;  // output geometry (before the optional 2x upsample)
  int cout;                 // real output channels (bf16 mode: multiple of BN)
  int tw, th, nb;           // tile = nb images x th rows x tw cols = 128 pixels
  int tiles_x, tiles_y;
  int ntiles_n, total_tiles;  // filled by launch_conv_tc: Cout tiles and (pixel tiles x Cout tiles)
  const float* bias;              // [Cout_pad] folded BatchNorm shift (or conv bias)
  const __nv_bfloat16* residual;  // optional, same geometry as the output
  void* out;
  int relu;
  int fp16;         // 16-bit storage format of activations / weights: 1 = IEEE fp16, 0 = bf16
  int up2x;         // 16-bit mode: write every pixel to its 2x2 nearest-upsampled positions
  int mode;         // EpilogueMode
  int num_classes;  // softmax modes
  // softmax NHWC addressing (lets one buffer be laid out destination-major for the multi-GPU exchange):
  //   off(n, y, x) = (((y / row_block) * slice_count + slice0 + n) * row_block + y % row_block) * out_w + x
  int slice0, slice_count, row_block;
  // optional (nullptr = off): 16 cycle counters per layer filled by the halo kernel's roles (development aid)
  //   [0] MMA wait accumulator  [1] MMA wait A  [2] MMA wait B  [3] MMA total  [4] gather wait empty
  //   [5] gather issue  [6] gather wait landing  [7] gather total  [8] epilogue wait  [9] epilogue body  [10] CTAs
  //   [11] CTA lifetime (first instruction to after the TMEM dealloc)
  unsigned long long* debug;
};

// Launch on `stream`; KC = min(64, cin), BN = min(128, Cout_pad).  Returns cudaGetLastError().
// bm = 2 (64-channel chunks, BN 128 / 256 only): CTA tiles of two M tiles that share every weight box.
// cluster = 2 (kc 64 with bn 256 / bm 1 or bn 128 / bm 2, layers with a single Cout tile): pairs of CTAs multicast the
// weight boxes to each other; needs `bmap` (bn 256) / `bmap2` (bn 128) as the half-box maps.
cudaError_t launch_conv_tc(const ConvArgs& args, int kc, int bn, cudaStream_t stream, int bm = 1, int cluster = 1);

// Shared memory the kernel variant needs (for occupancy planning / tests).
int conv_tc_smem_bytes(int kc, int bn);

// Halo-tile variant (conv_halo.cu) for stride-1 3x3 convs on images of at least 16x16: every CTA tile is
// a 16x16 output block whose 18x18 input halo is gathered ONCE per channel chunk into a planar layout,
// the nine taps being shifted views of it.  Same ConvArgs; tiling fields are overwritten (16x16, nb=1).
constexpr int kHaloTile = 16;
bool conv_halo_applicable(const ConvArgs& args);
cudaError_t launch_conv_halo(const ConvArgs& args, int kc, int bn, cudaStream_t stream);
// ... with the halo tile fetched by ONE TMA box per chunk into a 128B-swizzled buffer (needs `hmap`); for the layers
// the per-tap kernel serves otherwise: identity sources, 64-channel chunks, Cout a multiple of 128
bool conv_halo_tma_applicable(const ConvArgs& args);
cudaError_t launch_conv_halo_tma(const ConvArgs& args, cudaStream_t stream);

// CTA-pair variant (tcgen05 cta_group::2, M = 256, N = 128) for stride-1 3x3 convs with 64-channel chunks and
// Cout a multiple of 128; needs `bmap2`.
bool conv_pair_applicable(const ConvArgs& args);
cudaError_t launch_conv_pair(const ConvArgs& args, cudaStream_t stream);

// Per-tap TMA kernel on CTA pairs (conv_tc2.cu, tcgen05 cta_group::2, M = 256): 64-channel chunks, 16-bit outputs,
// Cout a multiple of bn = 256 (one pixel tile per CTA) or 128 (two pixel tiles per CTA); each CTA holds half of every
// weight tile.  Needs `bmap` (bn 256: half box (64, 128)) / `bmap2` (bn 128: half box (64, 64)).
bool conv_tc2_applicable(const ConvArgs& args, int bn);
cudaError_t launch_conv_tc2(const ConvArgs& args, int bn, cudaStream_t stream);

// Row-folded variant (conv_row.cu) for stride-1 3x3 convs with Cout_pad in {16, 32, 64} on images whose width is a
// multiple of 128: vertical taps folded into the MMA's N, residual as an identity K segment, TMA-store epilogue.
// Needs `bmapf` / `omap` (and `bmapi` with a residual).  conv_row_kc: the K chunk its weight maps are boxed with.
bool conv_row_applicable(const ConvArgs& args);
int conv_row_mode(const ConvArgs& args);   // 0 = not applicable, 1 = resident weights, 2 = streamed weights
int conv_row_kc(int cout_pad, int mode);
bool conv_row_tma_applicable(const ConvArgs& args);  // Cout 64 from one identity 64-channel source: A ring by TMA
int conv_row_store_rows(int cout_pad);  // output rows per TMA store box (the `omap` box height)
cudaError_t launch_conv_row(const ConvArgs& args, cudaStream_t stream);

// Fused decoder tail (conv_chain.cu): decoder block 4 conv1 (32 channels at half resolution, read through the 2x
// nearest upsample, -> 16) -> conv2 (16 -> 16) -> head (16 -> classes, softmax, oriented fp32 store) in one kernel; the
// two 16-channel full-resolution intermediates stay in shared-memory line buffers.
struct alignas(64) ChainArgs {
  CUtensorMap w1;   // conv1, four-slot fold (ConvArgs::bmapu of decoder block 4 conv1), box (16, 64)
  CUtensorMap w2;   // conv2, three-slot fold (ConvArgs::bmapf), box (16, 48)
  CUtensorMap w3;   // head, three-slot fold (ConvArgs::bmapf), box (16, 48)
  const __nv_bfloat16* src;  // decoder block 3 output [batch][h/2][w/2][32]
  const float* b1;  // [16] folded-BN shifts of conv1 / conv2
  const float* b2;
  int batch, h, w;  // full-resolution geometry
  int strips, total_items, groups;  // filled by launch_conv_chain
  int fp16;
  ConvArgs head;    // the head's epilogue description (bias, mode, out, num_classes, slice0, slice_count, row_block)
};
bool conv_chain_applicable(int h, int w);
cudaError_t launch_conv_chain(const ChainArgs& args, cudaStream_t stream);

// Stem (conv_stem.cu): Conv2d(1, 64, 7, stride 2, padding 3) + bias + ReLU on the tensor cores.
//   bmap: 2-D map over the packed 16-bit weights [64 cout][64 k], k = filter_row * 8 + filter_col (col 7 and
//         k >= 56 are zero), box (64, 64), 128-byte swizzle;  x: fp32 [batch][h][w];
//   omap: 4-D map (64, w/2, h/2, batch) over the output, box (64, 8, 16, 1), 128-byte swizzle (the epilogue's TMA store);
//   epi:  epilogue description (mode kEpiBf16, out = [batch][h/2][w/2][64], bias, relu, fp16, cout = 64, out_h, out_w).
//   pool_out: optional [batch][h/4][w/4][64], zeroed by the caller: the 3x3/s2/p1 max-pool of the output, fused into
//         the epilogue (interior windows stored, windows shared with a neighbouring tile combined with red.max).
cudaError_t launch_conv_stem(const CUtensorMap& bmap, const CUtensorMap& omap, const float* x, int batch, int h, int w,
                             const ConvArgs& epi, cudaStream_t stream, __nv_bfloat16* pool_out = nullptr);

}  // namespace iu
