// Stem of the encoder on the tensor cores: Conv2d(1, 64, 7, stride 2, padding 3) + folded BatchNorm + ReLU
// (`smp.Unet('resnet34').encoder.conv1/bn1/relu`, reached from `/root/reference/interactive_unet/unet.py:67`).
//
// One input channel makes the conv a [pixels x 49] x [49 x 64] GEMM.  K is laid out as 7 filter rows x 8 columns
// (the 8th column has zero weights) + 8 zero pad = 64, so that the A row of an output pixel is seven 16-byte
// chunks, each being 8 CONSECUTIVE input pixels of one input row (stride 2 makes every chunk start at an even
// column).  Per 16x16 block of output pixels a CTA
//   * stages the 37x38 fp32 input patch as fp16 in shared memory (zero outside the image = the conv's padding),
//   * 8 builder warps expand it into two 128-row A tiles in the 128B-swizzled K-major layout tcgen05.mma reads
//     (one thread per output pixel: 7 x (4 LDS.32 + 1 STS.128), both bank-conflict free),
//   * one thread issues 2 x 4 tcgen05.mma (M=128, N=64, K=16) into double-buffered TMEM accumulators,
//   * 8 epilogue warps apply bias + ReLU and store 16-bit NHWC (conv_epilogue.cuh).
// The [64 x 64] weight tile is loaded once per CTA.  Persistent over the tile list like the other conv kernels.
//
// Fused max-pool (`encoder.maxpool`, 3x3 / stride 2 / padding 1): the tile's 16x16 outputs sit in the epilogue's
// staging buffers anyway (for the TMA store), so the epilogue warps also reduce them to the 9x9 pooled pixels the tile
// touches.  The 7x7 interior windows lie wholly inside the tile and are stored directly; the windows along the
// tile's edges also cover a neighbouring tile's last row / column, so both tiles contribute their partial maximum
// with `red.global.max` on packed 16-bit pairs (post-ReLU values are >= 0, the pooled tensor is zeroed beforehand).
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

constexpr int kStemTileEdge = 16;                          // output pixels per tile edge
constexpr int kStemPatchRows = 2 * kStemTileEdge + 5;      // 37 input rows per tile
constexpr int kStemPatchCols = 2 * kStemTileEdge + 6;      // 38 input columns (7 taps + the zero-weight 8th)
constexpr int kStemPitch = 48;                             // halfs per patch row: rows 2*py apart land 16 banks apart
constexpr int kStemPatchBytes = (kStemPatchRows * kStemPitch * 2 + 127) / 128 * 128;
constexpr int kStemAStages = 3;
constexpr int kStemATile = 128 * 128;                      // one M tile: 128 rows x 64 fp16
constexpr int kStemAStage = 2 * kStemATile;
constexpr int kStemBBytes = 64 * 128;
constexpr int kStemBuilders = 256;
constexpr int kStemThreads = 320 + kStemBuilders;          // producer warp, MMA warp, 8 epilogue warps, 8 builder warps
constexpr int kStemStage = 128 * 128;                      // epilogue staging of one M tile: 128 pixels x 64 ch x 2 B
constexpr int kStemSmem = kStemAStages * kStemAStage + 4 * kStemStage + kStemBBytes + 2 * kStemPatchBytes + 64 * 4 + 16 * 8 + 16 + 1024;

struct StemArgs {
  CUtensorMap omap;  // 4-D map (64, W/2, H/2, N) over the output, box (64, 8, 16, 1), 128B swizzle: the epilogue's TMA store
  CUtensorMap bmap;  // 2-D map (K = 64, Cout = 64) over the packed weights, box = (64, 64), 128B swizzle
  const float* x;    // [batch][h][w] fp32 in [0, 1]
  __nv_bfloat16* pool_out;  // optional: the 3x3 / stride-2 / pad-1 max-pool of the output, [batch][h/4][w/4][64], ZEROED
                            // by the caller before the launch (tiles combine their border windows with red.max)
  int batch, h, w;
  ConvArgs epi;      // epilogue description (mode kEpiBf16, out, relu, fp16, cout = 64, out_h = h / 2, out_w = w / 2)
};

__global__ void __launch_bounds__(kStemThreads, 1) conv_stem_kernel(const __grid_constant__ StemArgs s) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t stg_base = a_base + kStemAStages * kStemAStage;  // 2 epilogue groups x 2 buffers, 1024-aligned
  const uint32_t b_base = stg_base + 4 * kStemStage;
  const uint32_t patch_base = b_base + kStemBBytes;
  const uint32_t bias_base = patch_base + 2 * kStemPatchBytes;
  const uint32_t bar_base = bias_base + 64 * 4;
  auto a_full = [&](int st) { return bar_base + 8u * st; };
  auto a_empty = [&](int st) { return bar_base + 8u * (kStemAStages + st); };
  const uint32_t b_full = bar_base + 16u * kStemAStages;
  auto acc_full = [&](int b) { return bar_base + 16u * kStemAStages + 8u + 8u * b; };
  auto acc_empty = [&](int b) { return bar_base + 16u * kStemAStages + 24u + 8u * b; };
  const uint32_t tmem_slot = bar_base + 16u * kStemAStages + 40u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_base - raw));
  uint8_t* a_ptr = smem_raw + (a_base - raw);
  uint16_t* patch_ptr = reinterpret_cast<uint16_t*>(smem_raw + (patch_base - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvArgs& a = s.epi;
  const int tiles_x = (a.out_w + kStemTileEdge - 1) / kStemTileEdge;
  const int tiles_y = (a.out_h + kStemTileEdge - 1) / kStemTileEdge;
  const int total_tiles = tiles_x * tiles_y * s.batch;

  if (threadIdx.x < 64) bias_s[threadIdx.x] = a.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&s.bmap);
    tma_prefetch_desc(&s.omap);
    for (int st = 0; st < kStemAStages; ++st) {
      mbar_init(a_full(st), kStemBuilders / 32);
      mbar_init(a_empty(st), 1);
    }
    mbar_init(b_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ weights: one TMA load for the whole kernel
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, kStemBBytes);
      tma_load_2d(b_base, &s.bmap, b_full, 0, 0);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_f16(kTileM, 64, a.fp16);
    const uint64_t bdesc = umma_smem_desc<128>(b_base);
    const uint32_t b_lo = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
    if (lane == 0) mbar_wait(b_full, 0);
    __syncwarp();
    operand_ready_fence();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      const int st = it % kStemAStages;
      if (lane == 0) {
        mbar_wait(acc_empty(buf), ((it >> 1) & 1u) ^ 1u);
        mbar_wait(a_full(st), (it / kStemAStages) & 1);
      }
      __syncwarp();
      tc_fence_after();
      const uint64_t adesc = umma_smem_desc<128>(a_base + st * kStemAStage);
      const uint32_t a_lo = (uint32_t)adesc, a_hi = (uint32_t)(adesc >> 32);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_lohi(tmem_base + (buf * 2u + j) * 64u, a_lo + j * (kStemATile >> 4) + 2u * kk, a_hi, b_lo + 2u * kk,
                          b_hi, idesc, kk ? 1u : 0u);
        }
        umma_commit(a_empty(st));
        umma_commit(acc_full(buf));
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue: group j owns M tile j (columns 8j..8j+7)
    const int j = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // TMEM lane = pixel (row >> 3, row & 7) of the 16 x 8 M tile
    const bool lead = ((warp - 2) & 3) == 0 && lane == 0;
    const uint32_t stg_j = stg_base + j * 2 * kStemStage;
    const uint32_t swz = (uint32_t)(row & 7);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
      if (lane == 0) mbar_wait(acc_full(buf), (it >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (buf * 2u + j) * 64u + ((uint32_t)(quarter * 32) << 16);
      // bias + ReLU + 16-bit pack of this pixel's 64 channels, then the accumulator is free again
      uint4 pk[8];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t acc[32];
        tmem_ld_32x32(taddr + ch * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; q += 8) {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 b2 = *reinterpret_cast<const float2*>(bias_s + ch * 32 + q + 2 * k);
            o[k] = relu16x2(pack16_sat(__uint_as_float(acc[q + 2 * k]) + b2.x, __uint_as_float(acc[q + 2 * k + 1]) + b2.y, a.fp16),
                            a.fp16);
          }
          pk[(ch * 32 + q) / 8] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      // staged TMA store of the 16 x 8 pixel tile (a per-lane 16-byte global store touches 32 lines per instruction
      // and made this kernel LSU bound): swizzled rows of 128 B, one box (64 ch, 8 px, 16 rows), clipped at the edges
      const uint32_t stg = stg_j + (it & 1u) * kStemStage;
      if (lead) bulk_wait_read_1();
      group_bar(2 + j);  // named barriers 2 / 3: id 1 belongs to the builder warps
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t dst = stg + (uint32_t)row * 128u + (((uint32_t)c ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[c].x), "r"(pk[c].y), "r"(pk[c].z),
                     "r"(pk[c].w)
                     : "memory");
      }
      group_bar(2 + j);
      if (lead) {
        fence_proxy_async();
        tma_store_4d(&s.omap, stg, 0, tx * kStemTileEdge + 8 * j, ty * kStemTileEdge, n);
        bulk_commit();
      }
      if (s.pool_out != nullptr) {
        // both groups' halves of the tile are staged: pool the 9 x 9 windows this tile touches, 8 channels per item
        asm volatile("bar.sync 4, 256;" ::: "memory");
        const int ph = a.out_h >> 1, pw = a.out_w >> 1;
        const uint32_t stg_it = stg_base + (it & 1u) * kStemStage;
        for (int item = (warp - 2) * 32 + lane; item < 81 * 8; item += 256) {
          const int chunk = item & 7, pp = item >> 3;
          const int py = pp / 9, px = pp - py * 9;
          const int gy = ty * (kStemTileEdge / 2) + py, gx = tx * (kStemTileEdge / 2) + px;
          if (gy >= ph || gx >= pw) continue;
          const int r_lo = py == 0 ? 0 : 2 * py - 1, r_hi = py == 8 ? 15 : 2 * py + 1;
          const int c_lo = px == 0 ? 0 : 2 * px - 1, c_hi = px == 8 ? 15 : 2 * px + 1;
          uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;   // +0.0: identity of max over non-negative values
          for (int r = r_lo; r <= r_hi; ++r)
            for (int cc = c_lo; cc <= c_hi; ++cc) {
              const int cg = cc & 7;
              const uint32_t src = stg_it + (uint32_t)(cc >> 3) * (2 * kStemStage) + (uint32_t)(r * 8 + cg) * 128u +
                                   (((uint32_t)chunk ^ (uint32_t)cg) << 4);
              uint32_t v0, v1, v2, v3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(src));
              m0 = __vmaxu2(m0, v0);   // unsigned order of the 16-bit patterns == numeric order for values >= 0
              m1 = __vmaxu2(m1, v1);
              m2 = __vmaxu2(m2, v2);
              m3 = __vmaxu2(m3, v3);
            }
          uint32_t* dst = reinterpret_cast<uint32_t*>(s.pool_out + (((size_t)n * ph + gy) * pw + gx) * 64 + chunk * 8);
          if (py >= 1 && py <= 7 && px >= 1 && px <= 7) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(m0, m1, m2, m3);
          } else if (a.fp16) {
            asm volatile("red.global.max.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
                         : "memory");
          } else {
            asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
                         : "memory");
          }
        }
      }
    }
    if (lead) bulk_wait_all();
  } else {
    // ------------------------------------------------------------ A builders (8 warps, one thread per output pixel)
    const int t = threadIdx.x - 320;
    const int py = t >> 4, px = t & 15;
    const int m = py * 8 + (px & 7);  // row inside the M tile; the tile's two halves are columns 0-7 / 8-15
    const uint32_t row_off = (uint32_t)(px >> 3) * kStemATile + (uint32_t)m * 128u;
    // the zero K columns 56..63 (chunk 7) never change: write them once in every stage
#pragma unroll
    for (int st = 0; st < kStemAStages; ++st)
      *reinterpret_cast<uint4*>(a_ptr + st * kStemAStage + row_off + ((7 ^ (m & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    constexpr int kPatchElems = kStemPatchRows * kStemPatchCols;
    constexpr int kPerThread = (kPatchElems + kStemBuilders - 1) / kStemBuilders;
    // Patches are fetched TWO tiles ahead into registers (`pre` = the next tile's, `pre2` = the one after): the global
    // latency of a fetch then hides under a whole tile's expansion instead of being waited for in store_patch.
    float pre[kPerThread], pre2[kPerThread];
    auto load_patch = [&](int tile, float (&dst)[kPerThread]) {
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
      const int iy0 = 2 * ty * kStemTileEdge - 3, ix0 = 2 * tx * kStemTileEdge - 3;
      const float* img = s.x + (size_t)n * s.h * s.w;
#pragma unroll
      for (int i = 0; i < kPerThread; ++i) {
        const int e = t + i * kStemBuilders;
        const int r = e / kStemPatchCols, c = e - r * kStemPatchCols;
        const int iy = iy0 + r, ix = ix0 + c;
        const bool ok = e < kPatchElems && (unsigned)iy < (unsigned)s.h && (unsigned)ix < (unsigned)s.w;
        dst[i] = ok ? __ldg(img + (size_t)iy * s.w + ix) : 0.0f;
      }
    };
    auto store_patch = [&](int which, const float (&src)[kPerThread]) {
      uint16_t* p = patch_ptr + which * (kStemPatchBytes / 2);
#pragma unroll
      for (int i = 0; i < kPerThread; ++i) {
        const int e = t + i * kStemBuilders;
        const int r = e / kStemPatchCols, c = e - r * kStemPatchCols;
        if (e < kPatchElems)
          p[r * kStemPitch + c] = a.fp16 ? __half_as_ushort(__float2half_rn(src[i]))
                                         : __bfloat16_as_ushort(__float2bfloat16_rn(src[i]));
      }
    };
    auto builders_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kStemBuilders) : "memory"); };
    uint32_t it = 0;
    int tile = blockIdx.x;
    if (tile < total_tiles) {
      load_patch(tile, pre);
      store_patch(0, pre);
      if (tile + (int)gridDim.x < total_tiles) load_patch(tile + gridDim.x, pre);
    }
    builders_sync();
    for (; tile < total_tiles; tile += gridDim.x, ++it) {
      const int st = it % kStemAStages;
      const int next = tile + gridDim.x, next2 = tile + 2 * gridDim.x;
      if (next2 < total_tiles) load_patch(next2, pre2);  // lands during this tile's and the next tile's expansion
      if (lane == 0) mbar_wait(a_empty(st), ((it / kStemAStages) & 1) ^ 1u);
      __syncwarp();
      const uint32_t* prow =
          reinterpret_cast<const uint32_t*>(patch_ptr + (it & 1u) * (kStemPatchBytes / 2) + (2 * py) * kStemPitch + 2 * px);
      uint8_t* arow = a_ptr + st * kStemAStage + row_off;
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const uint32_t* q = prow + r * (kStemPitch / 2);
        *reinterpret_cast<uint4*>(arow + ((r ^ (m & 7)) << 4)) = make_uint4(q[0], q[1], q[2], q[3]);
      }
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async();  // generic-proxy writes (ordered by the warp sync) -> visible to the tensor core's reads
        mbar_arrive(a_full(st));
      }
      if (next < total_tiles) store_patch((it + 1) & 1u, pre);  // fetched one iteration ago
#pragma unroll
      for (int i = 0; i < kPerThread; ++i) pre[i] = pre2[i];
      builders_sync();  // next patch complete; everyone is done reading the current one
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

cudaError_t launch_conv_stem(const CUtensorMap& bmap, const CUtensorMap& omap, const float* x, int batch, int h, int w,
                             const ConvArgs& epi, cudaStream_t stream, __nv_bfloat16* pool_out) {
  static_assert(kStemSmem <= 227 * 1024, "stem kernel exceeds the shared memory of an SM");
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured_dev = dev;
  }
  StemArgs s;
  s.bmap = bmap;
  s.omap = omap;
  s.x = x;
  s.pool_out = pool_out;
  s.batch = batch;
  s.h = h;
  s.w = w;
  s.epi = epi;
  const int tiles = ((h / 2 + kStemTileEdge - 1) / kStemTileEdge) * ((w / 2 + kStemTileEdge - 1) / kStemTileEdge) * batch;
  const int grid = tiles < num_sms ? tiles : num_sms;
  conv_stem_kernel<<<grid, kStemThreads, kStemSmem, stream>>>(s);
  return cudaGetLastError();
}

}  // namespace iu
