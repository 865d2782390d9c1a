// Halo-tile implicit-GEMM 3x3 convolution on the sm_100a tensor cores.
//
// conv_tc.cu fetches one shifted 128-pixel box per filter tap, i.e. every input pixel crosses L2->SM nine
// times; on B200 that makes the wide layers L2-bandwidth bound (~6300 B/clk chip-wide) and the narrow ones
// (16/32 channels = 32/64-byte rows) TMA row-rate bound.  Here a CTA owns a 16x16 block of output pixels and
//   * gathers its 18x18 input halo ONCE per 64-channel chunk (cp.async, zero-filled outside the image) into a
//     PLANAR shared-memory layout: plane k holds channels [8k, 8k+8) of all halo pixels, 16 bytes per pixel;
//   * in that layout the A operand of filter tap (r,q) is the SAME buffer read through a shifted UMMA
//     descriptor (no-swizzle K-major canonical layout: 8 pixels of an image row are one 8x16B core matrix,
//     SBO = one halo row, LBO = one plane), so the nine taps cost no extra shared-memory or L2 traffic;
//   * the block is two M=128 tiles (16 rows x 8 columns each) that share every weight tile -> 2 tcgen05.mma
//     per weight tile, 4 TMEM accumulators (2 tiles x double buffer);
//   * weights: streamed per tap through a TMA ring (wide layers), or loaded ONCE per CTA and kept resident
//     (layers with <= 32 input channels: their whole [Cout x 9*Cin] matrix is a few KB);
//   * warp roles: 1 weight-TMA thread, 1 MMA thread, 8 epilogue warps (one group per M tile), 4 gather warps;
//     waits are polled by one lane per warp; the folded-BN bias vector lives in shared memory.
// Everything after the MMA (bias / residual / ReLU / upsampled store / softmax head) is conv_epilogue.cuh.
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

constexpr int kHaloW = kHaloTile + 2;                // 18 halo pixels per edge
constexpr int kHaloPix = kHaloW * kHaloW;            // 324
constexpr int kPlaneStride = kHaloPix * 16 + 16;     // 5200 B; (stride / 16) is odd -> conflict-free cp.async stores
constexpr int kHaloThreads = 448;                    // 14 warps, see roles above
constexpr int kGatherThreads = 128;
constexpr int kHaloSmemBudget = 200 * 1024;
constexpr int kMaxBias = 512;

template <int KC, int BN>
struct HaloCfg {
  static constexpr bool STATIONARY = KC <= 32;  // whole weight matrix resident in shared memory
  static constexpr int CTAS = STATIONARY ? 2 : 1;
  static constexpr int PLANES = KC / 8;
  static constexpr int A_STAGE = (PLANES * kPlaneStride + 1023) / 1024 * 1024;
  static constexpr int A_STAGES = KC == 64 ? 3 : 4;
  static constexpr int SW = KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int B_ALLOC = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_RAW = (kHaloSmemBudget - A_STAGES * A_STAGE) / B_ALLOC;
  static constexpr int B_STAGES = STATIONARY ? 9 : (B_RAW > 9 ? 9 : (B_RAW < 3 ? 3 : B_RAW));
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;
  static constexpr int TMEM_COLS = 4 * ACC_COLS;  // 2 M tiles x 2 buffers: 128 / 128 / 256 / 512 columns
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_STAGES + 4;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + B_STAGES * B_ALLOC + kMaxBias * 4 + NBAR * 8 + 16 + 1024;
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// No-swizzle K-major descriptor: rows 16 B apart inside an 8-row core matrix, core matrices `sbo` bytes
// apart along M, the two 16-byte K halves of one MMA `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc_planar(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 |
         (uint64_t)1 << 46;
}

// One lane polls, the warp follows: keeps hundreds of threads from hammering the same mbarrier.
__device__ __forceinline__ void warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

template <int KC, int BN>
__global__ void __launch_bounds__(kHaloThreads, HaloCfg<KC, BN>::CTAS)
    conv_halo_kernel(const __grid_constant__ ConvArgs a) {
  using Cfg = HaloCfg<KC, BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + Cfg::A_STAGES * Cfg::A_STAGE;
  const uint32_t bias_base = b_base + Cfg::B_STAGES * Cfg::B_ALLOC;
  const uint32_t bar_base = bias_base + kMaxBias * 4;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 16u * Cfg::A_STAGES + 8u * s; };
  auto b_empty = [&](int s) { return bar_base + 16u * Cfg::A_STAGES + 8u * (Cfg::B_STAGES + s); };
  const uint32_t acc_bars = bar_base + 16u * (Cfg::A_STAGES + Cfg::B_STAGES);
  auto acc_full = [&](int b) { return acc_bars + 8u * b; };
  auto acc_empty = [&](int b) { return acc_bars + 16u + 8u * b; };
  const uint32_t tmem_slot = acc_bars + 32u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_base - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < a.ntiles_n * BN; i += kHaloThreads) bias_s[i] = a.bias[i];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.bmap);
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      mbar_init(a_full(s), kGatherThreads / 32);  // one arrival per gather warp (after every lane fenced its copies)
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < Cfg::B_STAGES; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer (TMA)
    // The warp stays converged and issues under elect.sync: the compiler then knows exactly one lane is
    // active and emits straight-line UTMALDG instead of a per-lane election loop.
    if constexpr (Cfg::STATIONARY) {
      // single source, single channel chunk, single Cout tile: 9 tap tiles, loaded once for all pixel tiles
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full(0), 9 * Cfg::B_BYTES);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(b_base + tap * Cfg::B_ALLOC, &a.bmap, b_full(0), tap * KC, 0);
      }
    } else {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int ntile = tile % a.ntiles_n;
        int kbase = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const int cin = a.seg[s].cin;
          for (int cc = 0; cc < cin / KC; ++cc) {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap, ++it) {
              const int st = it % Cfg::B_STAGES;
              warp_wait(b_empty(st), ((it / Cfg::B_STAGES) & 1) ^ 1u, lane);
              if (elect_one()) {
                mbar_arrive_expect_tx(b_full(st), Cfg::B_BYTES);
                tma_load_2d(b_base + st * Cfg::B_ALLOC, &a.bmap, b_full(st), kbase + tap * cin + cc * KC, ntile * BN);
              }
            }
          }
          kbase += 9 * cin;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const uint32_t idesc = umma_idesc_f16(kTileM, BN, a.fp16);
    uint32_t ita = 0, itb = 0, tcount = 0;
    const uint64_t bdesc_base = umma_smem_desc<Cfg::SW>(b_base);
    const uint32_t b_lo_base = (uint32_t)bdesc_base, b_hi = (uint32_t)(bdesc_base >> 32);
    if constexpr (Cfg::STATIONARY) {
      warp_wait(b_full(0), 0, lane);
      tc_fence_after();
    }
    const bool dbg = a.debug != nullptr;
    long long w_acc = 0, w_a = 0, w_b = 0, t_begin = dbg ? clock64() : 0, t0 = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1u;
      if (dbg) t0 = clock64();
      warp_wait(acc_empty(buf), ((tcount >> 1) & 1u) ^ 1u, lane);
      if (dbg) w_acc += clock64() - t0;
      tc_fence_after();
      const uint32_t tmem_d0 = tmem_base + (buf * 2u) * Cfg::ACC_COLS;
      const uint32_t tmem_d1 = tmem_d0 + Cfg::ACC_COLS;
      uint32_t accumulate = 0;
      for (int s = 0; s < a.nseg; ++s) {
        for (int cc = 0; cc < a.seg[s].cin / KC; ++cc, ++ita) {
          const int sta = ita % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          warp_wait(a_full(sta), (ita / Cfg::A_STAGES) & 1, lane);
          if (dbg) w_a += clock64() - t0;
          tc_fence_after();
          // Descriptors: one base per stage; every (tap, k-step, M tile) is base.lo + a compile-time constant
          // (in 16-byte units), so issuing an MMA costs one integer add instead of a bit-field rebuild.
          const uint64_t adesc_base = umma_smem_desc_planar(a_base + sta * Cfg::A_STAGE, kPlaneStride, kHaloW * 16);
          const uint32_t a_lo = (uint32_t)adesc_base, a_hi = (uint32_t)(adesc_base >> 32);
          if constexpr (Cfg::STATIONARY) {
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int r = tap / 3, q = tap - 3 * r;
                const uint32_t b_lo = b_lo_base + tap * (Cfg::B_ALLOC >> 4);
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_off = (uint32_t)(r * kHaloW + q) + (uint32_t)(2 * kk) * (kPlaneStride >> 4);
                  umma_f16_lohi(tmem_d0, a_lo + a_off, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  umma_f16_lohi(tmem_d1, a_lo + a_off + 8u, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit(a_empty(sta));
            }
            accumulate = 1;
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap, ++itb) {
              const int stb = itb % Cfg::B_STAGES;
              if (dbg) t0 = clock64();
              warp_wait(b_full(stb), (itb / Cfg::B_STAGES) & 1, lane);
              if (dbg) w_b += clock64() - t0;
              tc_fence_after();
              const int r = tap / 3, q = tap - 3 * r;
              const uint32_t b_lo = b_lo_base + stb * (Cfg::B_ALLOC >> 4);
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_off = (uint32_t)(r * kHaloW + q) + (uint32_t)(2 * kk) * (kPlaneStride >> 4);
                  umma_f16_lohi(tmem_d0, a_lo + a_off, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  umma_f16_lohi(tmem_d1, a_lo + a_off + 8u, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  accumulate = 1;
                }
                umma_commit(b_empty(stb));
                if (tap == 8) umma_commit(a_empty(sta));
              }
              accumulate = 1;
            }
          }
        }
      }
      if (elect_one()) umma_commit(acc_full(buf));
    }
    if (dbg && lane == 0) {
      atomicAdd(a.debug + 0, (unsigned long long)w_acc);
      atomicAdd(a.debug + 1, (unsigned long long)w_a);
      atomicAdd(a.debug + 2, (unsigned long long)w_b);
      atomicAdd(a.debug + 3, (unsigned long long)(clock64() - t_begin));
      atomicAdd(a.debug + 10, 1ull);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue: group j owns M tile j (columns 8j..8j+7)
    const int j = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t tcount = 0;
    const bool dbg = a.debug != nullptr && warp == 2 && lane == 0;
    long long w_full = 0, t_body = 0, t0 = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1u;
      const TileCoord tc = decode_tile(a, tile);
      if (dbg) t0 = clock64();
      warp_wait(acc_full(buf), (tcount >> 1) & 1u, lane);
      if (dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
      tc_fence_after();
      const uint32_t taddr = tmem_base + (buf * 2u + j) * Cfg::ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      const int y = tc.y0 + (row >> 3);
      const int x = tc.x0 + 8 * j + (row & 7);
      epilogue_pixel<BN>(a, bias_s, tc.ntile, taddr, tc.n0, y, x, (y < a.out_h) && (x < a.out_w));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      if (dbg) t_body += clock64() - t0;
    }
    if (dbg) {
      atomicAdd(a.debug + 8, (unsigned long long)w_full);
      atomicAdd(a.debug + 9, (unsigned long long)t_body);
    }
  } else {
    // ------------------------------------------------------------ halo gather (4 warps, cp.async)
    // Thread t copies channel octet kc = t % PLANES of the halo pixels p0, p0 + STEP, ...: a warp reads whole
    // pixel rows (coalesced) and the (row, col) of the next pixel follows incrementally, no divisions.
    constexpr int STEP = kGatherThreads / Cfg::PLANES;  // halo pixels between two copies of one thread
    constexpr int DY = STEP / kHaloW, DX = STEP % kHaloW;
    constexpr int DEPTH = Cfg::A_STAGES - 1;            // cp.async groups kept in flight
    const int t = threadIdx.x - (kHaloThreads - kGatherThreads);
    const int kc = t % Cfg::PLANES;
    const int p0 = t / Cfg::PLANES;
    const int hy0 = p0 / kHaloW, hx0 = p0 % kHaloW;
    const uint32_t dst_off = kc * kPlaneStride + p0 * 16;
    uint32_t it = 0;
    const bool dbg = a.debug != nullptr && t == 0;
    long long g_empty = 0, g_issue = 0, g_land = 0, t0 = 0, t_begin = dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(a, tile);
      for (int s = 0; s < a.nseg; ++s) {
        const int cin = a.seg[s].cin;
        const __nv_bfloat16* src = a.src_ptr[s];
        const __nv_bfloat16* img = src + (size_t)tc.n0 * a.out_h * a.out_w * cin + kc * 8;
        for (int cc = 0; cc < cin / KC; ++cc, ++it) {
          const int st = it % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          warp_wait(a_empty(st), ((it / Cfg::A_STAGES) & 1) ^ 1u, lane);
          if (dbg) { const long long t1 = clock64(); g_empty += t1 - t0; t0 = t1; }
          uint32_t dst = a_base + st * Cfg::A_STAGE + dst_off;
          int hy = hy0, hx = hx0;
#pragma unroll 4
          for (int p = p0; p < kHaloPix; p += STEP) {
            const int gy = tc.y0 - 1 + hy, gx = tc.x0 - 1 + hx;
            const bool ok = ((unsigned)gy < (unsigned)a.out_h) && ((unsigned)gx < (unsigned)a.out_w);
            const __nv_bfloat16* g = ok ? img + ((size_t)gy * a.out_w + gx) * cin + cc * KC : src;
            cp_async_16(dst, g, ok ? 16u : 0u);  // src-size 0 = zero fill (the convolution's padding)
            dst += STEP * 16;
            hx += DX;
            hy += DY;
            if (hx >= kHaloW) {
              hx -= kHaloW;
              hy += 1;
            }
          }
          cp_async_commit();
          if (dbg) { const long long t1 = clock64(); g_issue += t1 - t0; t0 = t1; }
          if (it >= (uint32_t)(DEPTH - 1)) {
            cp_async_wait<DEPTH - 1>();  // the group issued DEPTH-1 stages ago has landed
            if (dbg) g_land += clock64() - t0;
            fence_proxy_async();         // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full((it - (DEPTH - 1)) % Cfg::A_STAGES));
          }
        }
      }
    }
    if (dbg) {
      atomicAdd(a.debug + 4, (unsigned long long)g_empty);
      atomicAdd(a.debug + 5, (unsigned long long)g_issue);
      atomicAdd(a.debug + 6, (unsigned long long)g_land);
      atomicAdd(a.debug + 7, (unsigned long long)(clock64() - t_begin));
    }
    // drain: signal the last DEPTH-1 stages
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const uint32_t first = it >= (uint32_t)(DEPTH - 1) ? it - (DEPTH - 1) : 0u;
      for (uint32_t k = first; k < it; ++k) mbar_arrive(a_full(k % Cfg::A_STAGES));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int KC, int BN>
static bool halo_variant_ok(const ConvArgs& a) {
  using Cfg = HaloCfg<KC, BN>;
  const int cout_pad = (a.mode == kEpiBf16) ? a.cout : BN;
  if (cout_pad > kMaxBias) return false;
  if (Cfg::STATIONARY && (a.nseg != 1 || a.seg[0].cin != KC || cout_pad != BN)) return false;
  return true;
}

bool conv_halo_applicable(const ConvArgs& a) {
  if (a.nseg < 1 || a.nseg > 2 || a.out_h < kHaloTile || a.out_w < kHaloTile) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].ksize != 3 || a.seg[s].stride != 1 || a.seg[s].pad != 1 || a.src_ptr[s] == nullptr) return false;
  return true;
}

template <int KC, int BN>
static cudaError_t launch_halo_one(const ConvArgs& args_in, cudaStream_t stream) {
  using Cfg = HaloCfg<KC, BN>;
  static_assert(Cfg::SMEM_BYTES * Cfg::CTAS <= 227 * 1024, "halo kernel exceeds the shared memory of an SM");
  if (!halo_variant_ok<KC, BN>(args_in)) return launch_conv_tc(args_in, KC, BN, stream);
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<KC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured_dev = dev;
  }
  ConvArgs args = args_in;
  args.tw = kHaloTile;
  args.th = kHaloTile;
  args.nb = 1;
  args.tiles_x = (args.out_w + kHaloTile - 1) / kHaloTile;
  args.tiles_y = (args.out_h + kHaloTile - 1) / kHaloTile;
  const int cout_pad = (args.mode == kEpiBf16) ? args.cout : BN;
  args.ntiles_n = cout_pad / BN;
  args.total_tiles = args.tiles_x * args.tiles_y * args.batch * args.ntiles_n;
  const int slots = num_sms * Cfg::CTAS;
  const int grid = args.total_tiles < slots ? args.total_tiles : slots;
  conv_halo_kernel<KC, BN><<<grid, kHaloThreads, Cfg::SMEM_BYTES, stream>>>(args);
  return cudaGetLastError();
}

#define IU_HALO_DISPATCH(KC_, BN_) \
  if (kc == KC_ && bn == BN_) return launch_halo_one<KC_, BN_>(args, stream);

cudaError_t launch_conv_halo(const ConvArgs& args, int kc, int bn, cudaStream_t stream) {
  IU_HALO_DISPATCH(64, 128)
  IU_HALO_DISPATCH(64, 64)
  IU_HALO_DISPATCH(64, 32)
  IU_HALO_DISPATCH(32, 32)
  IU_HALO_DISPATCH(32, 16)
  IU_HALO_DISPATCH(16, 16)
  return cudaErrorInvalidValue;
}

}  // namespace iu
