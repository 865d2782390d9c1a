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
//   * warp roles: 1 weight-TMA thread, 1 MMA thread, 8 epilogue warps (one group per M tile), 4 or 8 gather warps;
//     waits are polled by one lane per warp; the folded-BN bias vector lives in shared memory.
// Everything after the MMA (bias / residual / ReLU / upsampled store / softmax head) is conv_epilogue.cuh.
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

constexpr int kHaloW = kHaloTile + 2;                // 18 halo pixels per edge
constexpr int kHaloPix = kHaloW * kHaloW;            // 324
constexpr int kPlaneStride = kHaloPix * 16 + 16;     // 5200 B; (stride / 16) is odd -> conflict-free cp.async stores
constexpr int kHaloBaseThreads = 320;                // warps 0-9: weight producer, MMA issuer, 8 epilogue warps
#ifndef IU_HALO_BUDGET_KB
#define IU_HALO_BUDGET_KB 200
#endif
#ifndef IU_HALO_A64_STAGES
#define IU_HALO_A64_STAGES 3
#endif
#ifndef IU_PAIR_A_STAGES
#define IU_PAIR_A_STAGES 3
#endif
#ifndef IU_PAIR_B_STAGES
#define IU_PAIR_B_STAGES 8
#endif
constexpr int kHaloSmemBudget = IU_HALO_BUDGET_KB * 1024;
constexpr int kMaxBias = 512;

// STAT: the layer's whole weight matrix stays resident in shared memory (loaded once per CTA) -- every layer with
// <= 32-channel chunks, and the 64-channel layers whose [BN x K] matrix fits beside the activation ring
// (64 -> 64 and 64+64 -> 32: 72 KB).  Otherwise weights stream per (chunk, tap) through a TMA ring.
// TMA_A (identity sources, 64-channel chunks, streamed weights): the halo tile is ONE TMA box -- 18 x 18 pixels x 64
// channels in the 128B-swizzled K-major layout (a pixel = one 128-byte row, 41,472 B) -- and filter tap (r,q) of
// M tile j is that buffer read from start address + ((r * 18 + q + 8j) * 128) bytes with SBO = one halo row
// (18 * 128 B): the tensor core swizzles on address bits, so an operand may start at any 128-byte row (measured on
// the row-folded kernel, profiles/r02_findings.md 1c).  One thread issues one box per chunk where eight warps issued
// 2592 cp.async; two A stages are enough then, and the shared memory they free deepens the weight ring to 7 tiles.
template <int KC, int BN, bool STAT, bool TMA_A = false>
struct HaloCfg {
  static constexpr int CTAS = KC <= 32 ? 2 : 1;
  static constexpr int GATHER_THREADS = TMA_A ? 32 : (KC <= 32 ? 128 : 256);  // 4 gather warps per CTA (two CTAs per SM) or 8
  static constexpr int THREADS = kHaloBaseThreads + GATHER_THREADS;
  static constexpr int PLANES = KC / 8;
  static constexpr int T_PITCH_MAX = 24;             // TMA_A: halo pixels per row of the box (ConvArgs::halo_tma: 18, or
                                                     //        24 = every 8-row group starts a fresh swizzle period)
  static constexpr int T_BYTES = kHaloW * T_PITCH_MAX * KC * 2;  //  bytes the largest box lands
  static constexpr int A_STAGE = ((TMA_A ? T_BYTES : PLANES * kPlaneStride) + 1023) / 1024 * 1024;
  static constexpr int A_STAGES = TMA_A ? 2 : (KC == 64 ? (STAT ? 3 : IU_HALO_A64_STAGES) : 4);
  static_assert(!TMA_A || (KC == 64 && !STAT), "TMA-filled halo tiles: 64-channel chunks, streamed weights");
  static constexpr int SW = KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int B_ALLOC = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_RAW = ((TMA_A ? 220 * 1024 : kHaloSmemBudget) - A_STAGES * A_STAGE) / B_ALLOC;
  // STAT: number of resident [BN x KC] weight tiles; streaming: depth of the weight ring
  static constexpr int B_STAGES = STAT ? (KC <= 32 ? 9 : (B_RAW > 18 ? 18 : B_RAW)) : (B_RAW > 9 ? 9 : (B_RAW < 3 ? 3 : B_RAW));
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

// K-major 128B-swizzled descriptor whose 8-row groups are `sbo` bytes apart (a halo row instead of the dense 1024)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (uint64_t)1 << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46 |
         (uint64_t)2 << 61;
}

template <int KC, int BN, bool STAT, bool TMA_A = false>
__global__ void __launch_bounds__(HaloCfg<KC, BN, STAT, TMA_A>::THREADS, HaloCfg<KC, BN, STAT, TMA_A>::CTAS)
    conv_halo_kernel(const __grid_constant__ ConvArgs a) {
  using Cfg = HaloCfg<KC, BN, STAT, TMA_A>;
  const long long t_cta = (a.debug != nullptr && threadIdx.x == 0) ? clock64() : 0;
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

  for (int i = threadIdx.x; i < a.ntiles_n * BN; i += Cfg::THREADS) bias_s[i] = a.bias[i];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.bmap);
    if (TMA_A)
      for (int s = 0; s < a.nseg; ++s) tma_prefetch_desc(&a.hmap[s]);
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      // one arrival per gather warp (after every lane fenced its copies), or the A box's expect_tx alone
      mbar_init(a_full(s), TMA_A ? 1 : Cfg::GATHER_THREADS / 32);
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
    if constexpr (STAT) {
      // single Cout tile: every (segment, chunk, tap) weight tile is loaded once, for all pixel tiles of this CTA
      if (elect_one()) {
        int ntl = 0;
        for (int s = 0; s < a.nseg; ++s) ntl += 9 * (a.seg[s].cin / KC);
        mbar_arrive_expect_tx(b_full(0), ntl * Cfg::B_BYTES);
        int kbase = 0, idx = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const int cin = a.seg[s].cin;
          for (int cc = 0; cc < cin / KC; ++cc)
            for (int tap = 0; tap < 9; ++tap, ++idx)
              tma_load_2d(b_base + idx * Cfg::B_ALLOC, &a.bmap, b_full(0), kbase + tap * cin + cc * KC, 0);
          kbase += 9 * cin;
        }
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
    if constexpr (STAT) {
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
      uint32_t chunk = 0;  // (segment, chunk) counter of this tile: index of its 9 resident weight tiles
      for (int s = 0; s < a.nseg; ++s) {
        for (int cc = 0; cc < a.seg[s].cin / KC; ++cc, ++ita, ++chunk) {
          const int sta = ita % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          warp_wait(a_full(sta), (ita / Cfg::A_STAGES) & 1, lane);
          if (dbg) w_a += clock64() - t0;
          operand_ready_fence();
          // Descriptors: one base per stage; every (tap, k-step, M tile) is base.lo + a compile-time constant
          // (in 16-byte units), so issuing an MMA costs one integer add instead of a bit-field rebuild.
          const uint32_t pitch = TMA_A ? (uint32_t)a.halo_tma : (uint32_t)kHaloW;  // halo pixels per buffer row
          const uint64_t adesc_base = TMA_A ? umma_smem_desc_sw128(a_base + sta * Cfg::A_STAGE, pitch * 128u)
                                            : umma_smem_desc_planar(a_base + sta * Cfg::A_STAGE, kPlaneStride, kHaloW * 16);
          const uint32_t a_lo = (uint32_t)adesc_base, a_hi = (uint32_t)(adesc_base >> 32);
          // 16-byte units per halo pixel / per 16-channel k-step / between the two M tiles (8 pixels)
          constexpr uint32_t PX = TMA_A ? 8u : 1u, KSTEP = TMA_A ? 2u : 2u * (kPlaneStride >> 4), MT = 8u * PX;
          if constexpr (STAT) {
            const uint32_t b_lo_chunk = b_lo_base + chunk * 9u * (Cfg::B_ALLOC >> 4);
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int r = tap / 3, q = tap - 3 * r;
                const uint32_t b_lo = b_lo_chunk + tap * (Cfg::B_ALLOC >> 4);
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
              operand_ready_fence();
              const int r = tap / 3, q = tap - 3 * r;
              const uint32_t b_lo = b_lo_base + stb * (Cfg::B_ALLOC >> 4);
              const uint32_t a_tap = a_lo + ((uint32_t)r * pitch + (uint32_t)q) * PX;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_off = (uint32_t)kk * KSTEP;
                  umma_f16_lohi(tmem_d0, a_tap + a_off, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  umma_f16_lohi(tmem_d1, a_tap + a_off + MT, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
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
      const int y = tc.y0 + (row >> 3);
      const int x = tc.x0 + 8 * j + (row & 7);
      const bool valid = (y < a.out_h) && (x < a.out_w);
      uint4 res[EpiCfg<BN>::RV];
      residual_prefetch<BN>(a, tc.ntile, tc.n0, y, x, valid, res);
      if (dbg) t0 = clock64();
      warp_wait(acc_full(buf), (tcount >> 1) & 1u, lane);
      if (dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
      tc_fence_after();
      const uint32_t taddr = tmem_base + (buf * 2u + j) * Cfg::ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      epilogue_pixel<BN>(a, bias_s, tc.ntile, taddr, tc.n0, y, x, valid, res);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      if (dbg) t_body += clock64() - t0;
    }
    if (dbg) {
      atomicAdd(a.debug + 8, (unsigned long long)w_full);
      atomicAdd(a.debug + 9, (unsigned long long)t_body);
    }
  } else if (TMA_A) {
    // ------------------------------------------------------------ halo tiles by TMA: one 18 x 18 x 64 box per chunk
    if (warp == 10 && lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(a, tile);
        for (int s = 0; s < a.nseg; ++s) {
          for (int cc = 0; cc < a.seg[s].cin / KC; ++cc, ++it) {
            const int st = it % Cfg::A_STAGES;
            mbar_wait(a_empty(st), ((it / Cfg::A_STAGES) & 1) ^ 1u);
            mbar_arrive_expect_tx(a_full(st), (uint32_t)(kHaloW * a.halo_tma * KC * 2));
            // pixels outside the image arrive as zeros: the convolution's padding
            tma_load_4d(a_base + st * Cfg::A_STAGE, &a.hmap[s], a_full(st), cc * KC, tc.x0 - 1, tc.y0 - 1, tc.n0);
          }
        }
      }
    }
  } else if constexpr (!TMA_A) {
    // ------------------------------------------------------------ halo gather (4 warps, cp.async)
    // Thread t copies channel octet kc = t % PLANES of the halo pixels p0, p0 + STEP, ...: a warp reads whole
    // pixel rows (coalesced) and the (row, col) of the next pixel follows incrementally, no divisions.
    constexpr int STEP = Cfg::GATHER_THREADS / Cfg::PLANES;  // halo pixels between two copies of one thread
    constexpr int DY = STEP / kHaloW, DX = STEP % kHaloW;
    constexpr int DEPTH = Cfg::A_STAGES - 1;            // cp.async groups kept in flight
    const int t = threadIdx.x - kHaloBaseThreads;
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
        const int up = a.seg[s].up;  // 1: the source is half resolution and read through a 2x nearest upsample
        const int src_w = a.out_w >> up;
        const __nv_bfloat16* src = a.src_ptr[s];
        const __nv_bfloat16* img = src + (size_t)tc.n0 * (a.out_h >> up) * src_w * cin + kc * 8;
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
            const __nv_bfloat16* g = ok ? img + ((size_t)(gy >> up) * src_w + (gx >> up)) * cin + cc * KC : src;
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
            // every lane's copies have landed; ONE proxy fence per warp (after the warp-level sync that orders the
            // lanes' writes before it) makes them visible to the tensor core's async-proxy reads: a fence per thread
            // serialises ~10 cycles x 256 threads per stage
            __syncwarp();
            if (lane == 0) {
              fence_proxy_async();
              mbar_arrive(a_full((it - (DEPTH - 1)) % Cfg::A_STAGES));
            }
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
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();
      const uint32_t first = it >= (uint32_t)(DEPTH - 1) ? it - (DEPTH - 1) : 0u;
      for (uint32_t k = first; k < it; ++k) mbar_arrive(a_full(k % Cfg::A_STAGES));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (a.debug != nullptr && threadIdx.x == 0) atomicAdd(a.debug + 11, (unsigned long long)(clock64() - t_cta));
}

// =====================================================================================================
// CTA-pair variant for the wide layers (64-channel chunks, Cout a multiple of 128): two CTAs of a cluster sit
// on the two SMs of a TPC and run ONE tcgen05.mma.cta_group::2 stream (M = 256, N = 128).  Each CTA owns a
// 16x16 pixel block (its two 128-row A tiles, gathered exactly as above) and HALF of every [128 x 64] weight
// tile: the tensor cores exchange the B halves, so the weight traffic L2 -> SM and the B share of the
// shared-memory reads are halved -- the two limits the single-CTA kernel runs into on these layers
// (profiles/r01_mma_probe.txt).  Barrier protocol:
//   a_full / b_full / acc_empty : waited by the leader's MMA thread; arrivals come from both CTAs
//                                 (remote mbarrier.arrive, TMA complete_tx addressed to the leader);
//   a_empty / b_empty / acc_full: signalled in both CTAs at once by the leader's multicast tcgen05.commit.
struct PairCfg {
  static constexpr int KC = 64, BN = 128;
  static constexpr int GATHER_THREADS = 256;
  static constexpr int THREADS = kHaloBaseThreads + GATHER_THREADS;
  static constexpr int PLANES = KC / 8;
  static constexpr int A_STAGE = (PLANES * kPlaneStride + 1023) / 1024 * 1024;
  static constexpr int A_STAGES = IU_PAIR_A_STAGES;
  static constexpr int B_HALF_BYTES = (BN / 2) * KC * 2;  // 8 KB: this CTA's 64 weight rows of one (chunk, tap)
  static constexpr int B_STAGES = IU_PAIR_B_STAGES;
  static constexpr int TMEM_COLS = 4 * BN;                // 2 M tiles x 2 buffers
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_STAGES + 4;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + B_STAGES * B_HALF_BYTES + kMaxBias * 4 + NBAR * 8 + 16 + 1024;
};

__global__ void __launch_bounds__(PairCfg::THREADS, 1) conv_pair_kernel(const __grid_constant__ ConvArgs a) {
  using Cfg = PairCfg;
  constexpr int KC = Cfg::KC, BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + Cfg::A_STAGES * Cfg::A_STAGE;
  const uint32_t bias_base = b_base + Cfg::B_STAGES * Cfg::B_HALF_BYTES;
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
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // super tile s -> (Cout tile, pair of pixel blocks); this CTA's block is 2 * (s / ntiles_n) + rank
  const int n_blocks = a.tiles_x * a.tiles_y * a.batch;
  const int n_super = ((n_blocks + 1) >> 1) * a.ntiles_n;

  for (int i = threadIdx.x; i < a.ntiles_n * BN; i += Cfg::THREADS) bias_s[i] = a.bias[i];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.bmap2);
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      mbar_init(a_full(s), 2 * Cfg::GATHER_THREADS / 32);  // the gather warps of both CTAs
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < Cfg::B_STAGES; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 16);  // the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before anyone signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer: this CTA's 64 rows of every tile
    uint32_t it = 0;
    for (int st_ = pair_id; st_ < n_super; st_ += n_pairs) {
      const int ntile = st_ % a.ntiles_n;
      int kbase = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int cin = a.seg[s].cin;
        for (int cc = 0; cc < cin / KC; ++cc) {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap, ++it) {
            const int st = it % Cfg::B_STAGES;
            warp_wait(b_empty(st), ((it / Cfg::B_STAGES) & 1) ^ 1u, lane);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(b_full(st), 2 * Cfg::B_HALF_BYTES);
              tma_load_2d_pair(b_base + st * Cfg::B_HALF_BYTES, &a.bmap2, mapa_shared(b_full(st), 0),
                               kbase + tap * cin + cc * KC, ntile * BN + (int)rank * (BN / 2));
            }
          }
        }
        kbase += 9 * cin;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(2 * kTileM, BN, a.fp16);
      uint32_t ita = 0, itb = 0, tcount = 0;
      const uint64_t bdesc_base = umma_smem_desc<128>(b_base);
      const uint32_t b_lo_base = (uint32_t)bdesc_base, b_hi = (uint32_t)(bdesc_base >> 32);
      const bool dbg = a.debug != nullptr;
      long long w_acc = 0, w_a = 0, w_b = 0, t_begin = dbg ? clock64() : 0, t0 = 0;
      for (int st_ = pair_id; st_ < n_super; st_ += n_pairs, ++tcount) {
        const uint32_t buf = tcount & 1u;
        if (dbg) t0 = clock64();
        if (lane == 0) mbar_wait_cluster(acc_empty(buf), ((tcount >> 1) & 1u) ^ 1u);
        __syncwarp();
        if (dbg) w_acc += clock64() - t0;
        tc_fence_after();
        const uint32_t tmem_d0 = tmem_base + (buf * 2u) * BN;
        const uint32_t tmem_d1 = tmem_d0 + BN;
        uint32_t accumulate = 0;
        for (int s = 0; s < a.nseg; ++s) {
          for (int cc = 0; cc < a.seg[s].cin / KC; ++cc, ++ita) {
            const int sta = ita % Cfg::A_STAGES;
            if (dbg) t0 = clock64();
            if (lane == 0) mbar_wait_cluster(a_full(sta), (ita / Cfg::A_STAGES) & 1);
            __syncwarp();
            if (dbg) w_a += clock64() - t0;
            operand_ready_fence();
            const uint64_t adesc_base = umma_smem_desc_planar(a_base + sta * Cfg::A_STAGE, kPlaneStride, kHaloW * 16);
            const uint32_t a_lo = (uint32_t)adesc_base, a_hi = (uint32_t)(adesc_base >> 32);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap, ++itb) {
              const int stb = itb % Cfg::B_STAGES;
              if (dbg) t0 = clock64();
              if (lane == 0) mbar_wait_cluster(b_full(stb), (itb / Cfg::B_STAGES) & 1);
              __syncwarp();
              if (dbg) w_b += clock64() - t0;
              operand_ready_fence();
              const int r = tap / 3, q = tap - 3 * r;
              const uint32_t b_lo = b_lo_base + stb * (Cfg::B_HALF_BYTES >> 4);
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_off = (uint32_t)(r * kHaloW + q) + (uint32_t)(2 * kk) * (kPlaneStride >> 4);
                  umma_f16_pair_lohi(tmem_d0, a_lo + a_off, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  umma_f16_pair_lohi(tmem_d1, a_lo + a_off + 8u, a_hi, b_lo + 2u * kk, b_hi, idesc, accumulate);
                  accumulate = 1;
                }
                umma_commit_pair(b_empty(stb));
                if (tap == 8) umma_commit_pair(a_empty(sta));
              }
              accumulate = 1;
            }
          }
        }
        if (elect_one()) umma_commit_pair(acc_full(buf));
      }
      if (dbg && lane == 0) {
        atomicAdd(a.debug + 0, (unsigned long long)w_acc);
        atomicAdd(a.debug + 1, (unsigned long long)w_a);
        atomicAdd(a.debug + 2, (unsigned long long)w_b);
        atomicAdd(a.debug + 3, (unsigned long long)(clock64() - t_begin));
        atomicAdd(a.debug + 10, 1ull);
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue: group j owns this CTA's M tile j
    const int j = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t acc_empty_leader = mapa_shared(acc_empty(0), 0);
    uint32_t tcount = 0;
    const bool dbg = a.debug != nullptr && rank == 0 && warp == 2 && lane == 0;
    long long w_full = 0, t_body = 0, t0 = 0;
    for (int st_ = pair_id; st_ < n_super; st_ += n_pairs, ++tcount) {
      const uint32_t buf = tcount & 1u;
      const int block = 2 * (st_ / a.ntiles_n) + (int)rank;
      const TileCoord tc = decode_tile(a, block * a.ntiles_n + st_ % a.ntiles_n);
      const int y = tc.y0 + (row >> 3);
      const int x = tc.x0 + 8 * j + (row & 7);
      const bool valid = (block < n_blocks) && (y < a.out_h) && (x < a.out_w);
      uint4 res[EpiCfg<BN>::RV];
      residual_prefetch<BN>(a, tc.ntile, tc.n0, y, x, valid, res);
      if (dbg) t0 = clock64();
      warp_wait(acc_full(buf), (tcount >> 1) & 1u, lane);
      if (dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
      tc_fence_after();
      const uint32_t taddr = tmem_base + (buf * 2u + j) * BN + ((uint32_t)(quarter * 32) << 16);
      epilogue_pixel<BN>(a, bias_s, tc.ntile, taddr, tc.n0, y, x, valid, res);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8u * buf);
      if (dbg) t_body += clock64() - t0;
    }
    if (dbg) {
      atomicAdd(a.debug + 8, (unsigned long long)w_full);
      atomicAdd(a.debug + 9, (unsigned long long)t_body);
    }
  } else {
    // ------------------------------------------------------------ halo gather of this CTA's block (8 warps, cp.async)
    constexpr int STEP = Cfg::GATHER_THREADS / Cfg::PLANES;
    constexpr int DY = STEP / kHaloW, DX = STEP % kHaloW;
    constexpr int DEPTH = Cfg::A_STAGES - 1;
    const int t = threadIdx.x - kHaloBaseThreads;
    const int kc = t % Cfg::PLANES;
    const int p0 = t / Cfg::PLANES;
    const int hy0 = p0 / kHaloW, hx0 = p0 % kHaloW;
    const uint32_t dst_off = kc * kPlaneStride + p0 * 16;
    const uint32_t a_full_leader = mapa_shared(a_full(0), 0);
    uint32_t it = 0;
    const bool dbg = a.debug != nullptr && rank == 0 && t == 0;
    long long g_empty = 0, g_issue = 0, g_land = 0, t0 = 0, t_begin = dbg ? clock64() : 0;
    for (int st_ = pair_id; st_ < n_super; st_ += n_pairs) {
      const int block = 2 * (st_ / a.ntiles_n) + (int)rank;
      const TileCoord tc = decode_tile(a, block * a.ntiles_n);
      const bool live = block < n_blocks;  // an odd block count leaves the last pair's second CTA with zeros
      for (int s = 0; s < a.nseg; ++s) {
        const int cin = a.seg[s].cin;
        const int up = a.seg[s].up;
        const int src_w = a.out_w >> up;
        const __nv_bfloat16* src = a.src_ptr[s];
        const __nv_bfloat16* img = src + (size_t)tc.n0 * (a.out_h >> up) * src_w * cin + kc * 8;
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
            const bool ok = live && ((unsigned)gy < (unsigned)a.out_h) && ((unsigned)gx < (unsigned)a.out_w);
            const __nv_bfloat16* g = ok ? img + ((size_t)(gy >> up) * src_w + (gx >> up)) * cin + cc * KC : src;
            cp_async_16(dst, g, ok ? 16u : 0u);
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
            cp_async_wait<DEPTH - 1>();
            if (dbg) g_land += clock64() - t0;
            __syncwarp();
            if (lane == 0) {
              fence_proxy_async();
              mbar_arrive_cluster(a_full_leader + 8u * ((it - (DEPTH - 1)) % Cfg::A_STAGES));
            }
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
    cp_async_wait<0>();
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();
      const uint32_t first = it >= (uint32_t)(DEPTH - 1) ? it - (DEPTH - 1) : 0u;
      for (uint32_t k = first; k < it; ++k) mbar_arrive_cluster(a_full_leader + 8u * (k % Cfg::A_STAGES));
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still signal this CTA's barriers / read its weight half until here
  if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

bool conv_pair_applicable(const ConvArgs& a) {
  if (!conv_halo_applicable(a)) return false;
  const int cout_pad = (a.mode == kEpiBf16) ? a.cout : 0;
  if (cout_pad == 0 || cout_pad % PairCfg::BN || cout_pad > kMaxBias) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].cin % PairCfg::KC) return false;
  return true;
}

cudaError_t launch_conv_pair(const ConvArgs& args_in, cudaStream_t stream) {
  using Cfg = PairCfg;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "pair kernel exceeds the shared memory of an SM");
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
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
  args.ntiles_n = args.cout / Cfg::BN;
  args.total_tiles = args.tiles_x * args.tiles_y * args.batch * args.ntiles_n;
  const int n_blocks = args.tiles_x * args.tiles_y * args.batch;
  const int n_super = ((n_blocks + 1) / 2) * args.ntiles_n;
  const int max_pairs = num_sms / 2;
  const int pairs = n_super < max_pairs ? n_super : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_pair_kernel, args);
}

static int halo_weight_tiles(const ConvArgs& a, int kc) {
  int n = 0;
  for (int s = 0; s < a.nseg; ++s) n += 9 * (a.seg[s].cin / kc);
  return n;
}

template <int KC, int BN, bool STAT>
static bool halo_variant_ok(const ConvArgs& a) {
  using Cfg = HaloCfg<KC, BN, STAT>;
  const int cout_pad = (a.mode == kEpiBf16) ? a.cout : BN;
  if (cout_pad > kMaxBias) return false;
  if (STAT && (cout_pad != BN || halo_weight_tiles(a, KC) > Cfg::B_STAGES)) return false;
  return true;
}

bool conv_halo_applicable(const ConvArgs& a) {
  if (a.nseg < 1 || a.nseg > 2) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].ksize != 3 || a.seg[s].stride != 1 || a.seg[s].pad != 1 || a.src_ptr[s] == nullptr) return false;
  return true;
}

template <int KC, int BN, bool STAT, bool TMA_A = false>
static cudaError_t launch_halo_one(const ConvArgs& args_in, cudaStream_t stream) {
  using Cfg = HaloCfg<KC, BN, STAT, TMA_A>;
  static_assert(Cfg::SMEM_BYTES * Cfg::CTAS <= 227 * 1024, "halo kernel exceeds the shared memory of an SM");
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<KC, BN, STAT, TMA_A>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
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
  conv_halo_kernel<KC, BN, STAT, TMA_A><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(args);
  return cudaGetLastError();
}

// Weights resident when the variant allows it, else streamed (64-channel chunks only); layers that fit neither
// (e.g. <= 32-channel chunks with several Cout tiles -- none in this network) go to the per-tap kernel, which
// cannot read an upsampled source.
template <int KC, int BN>
static cudaError_t launch_halo_pick(const ConvArgs& args, cudaStream_t stream) {
  if constexpr (BN <= 64) {
    if (halo_variant_ok<KC, BN, true>(args)) return launch_halo_one<KC, BN, true>(args, stream);
  }
  if constexpr (KC == 64) {
    if (halo_variant_ok<KC, BN, false>(args)) return launch_halo_one<KC, BN, false>(args, stream);
  }
  for (int s = 0; s < args.nseg; ++s)
    if (args.seg[s].up) return cudaErrorInvalidValue;
  return launch_conv_tc(args, KC, BN, stream);
}

// Halo tiles filled by TMA (`hmap`): stride-1 3x3 layers whose sources are all identity tensors with 64-channel
// chunks and whose Cout is a multiple of 128 -- the layers the per-tap kernel serves otherwise.
bool conv_halo_tma_applicable(const ConvArgs& a) {
  if (!conv_halo_applicable(a) || a.mode != kEpiBf16 || a.cout % 128 != 0 || a.cout > kMaxBias || a.up2x) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].up || a.seg[s].cin % 64 != 0) return false;
  return true;
}
cudaError_t launch_conv_halo_tma(const ConvArgs& args, cudaStream_t stream) {
  if (!conv_halo_tma_applicable(args)) return cudaErrorInvalidValue;
  return launch_halo_one<64, 128, false, true>(args, stream);
}

#define IU_HALO_DISPATCH(KC_, BN_) \
  if (kc == KC_ && bn == BN_) return launch_halo_pick<KC_, BN_>(args, stream);

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
