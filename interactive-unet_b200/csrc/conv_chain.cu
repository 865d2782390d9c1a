// Fused tail of the decoder: decoder block 4 conv1 -> conv2 -> segmentation head (+ softmax + oriented store) in ONE
// kernel (`smp.Unet.decoder.blocks[4]` + `segmentation_head` + the reference's Softmax, reached from
// `/root/reference/interactive_unet/unet.py:67`; the head's store is `predict.py:98-106`).
//
// Why: run as three row-folded launches these layers are HBM bound -- the two 16-channel full-resolution intermediates
// are written and read back once each (2 x 32 B in + 2 x 32 B out per pixel against 16 B of real input and 4*C B of
// real output), 17 % of the step at ~4.2 TB/s (profiles/r01_findings.md, finding 10).  Here they never leave the SM:
//   * a work item is one image x one 124-column strip; the CTA streams down the strip in groups of 8 rows;
//   * every layer is the row-folded formulation of conv_row.cu (an M tile = 128 consecutive pixels of one image row,
//     vertical taps folded into the MMA's N: N = 4*16 for the upsampled conv1, 3*16 for conv2 and the head), all on
//     the SAME 128 columns [x0-2, x0+126); conv1's output is valid on all of them, conv2's on the inner 126, the
//     head's on the inner 124 -- the strip's own columns.  The 4 recomputed columns per strip are the whole price of
//     the fusion (128/124 = 3 % at 2048 columns, 25 % at 512 where the fifth strip is mostly empty);
//   * conv1's and conv2's outputs live in two LINE BUFFERS in shared memory (planar layout, 16 B per pixel and 8
//     channels, 18-row rings): the epilogue warps write a finished row straight into the ring in the layout the next
//     layer's A operand descriptor reads (input row r, filter column kx = ring row r shifted by kx*16 bytes), so no
//     row is ever recomputed vertically.  Rows outside the image read a permanent all-zero ring row (the convs' zero
//     padding), columns outside the image are written as zeros by the epilogue;
//   * software pipeline over row groups: iteration i issues conv1 of group i, conv2 of group i-1 (rows shifted up by
//     one) and the head of group i-2 (shifted by two), each into its own TMEM accumulator (3 x 128 columns), while the
//     epilogue warps drain the previous ones and the gather warps fetch group i+1's six source rows (cp.async, the 2x
//     nearest upsample is `src = dst >> 1` along x and the four-slot pre-summed filters along y, as in conv_row.cu).
// Warp roles (640 threads, one persistent CTA per SM): warp 0 loads the three layers' weights once (TMA, 21 KB,
// resident), warp 1 issues the MMAs, warps 2-9 are the epilogue (two sets of four, rows 0-3 / 4-7 of each group),
// eight of the warps 10-19 gather.
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

constexpr int kChainValid = 124;                  // output columns a strip owns
constexpr int kChainSlots = 130;                  // pixel slots per buffered row: columns x0-3 .. x0+126
constexpr int kChainPitch = kChainSlots * 16;     // bytes per row inside a plane
constexpr int kChainR = 8;                        // rows per group
constexpr int kChainRing = 18;                    // line-buffer depth (rows); ring row 18 is the permanent zero row
constexpr int kChainRingPlane = (kChainRing + 1) * kChainPitch + 16;  // (stride / 16) odd: planes 16 B apart mod 32 banks
constexpr int kChainRingBytes = 2 * kChainRingPlane;                  // 16 channels = 2 planes
constexpr int kChainInRows = kChainR / 2 + 2;     // source rows gathered per group of 8 upsampled rows
constexpr int kChainInPlane = kChainInRows * kChainPitch + 16;
constexpr int kChainInStage = 2 * kChainInPlane;  // one 16-channel chunk of the 32-channel source
constexpr int kChainW1Tile = 4 * 16 * 32;         // conv1: four-slot tile of one (chunk, kx): 64 rows x 16 ch
constexpr int kChainWTile = 3 * 16 * 32;          // conv2 / head: three-slot tile of one kx
constexpr int kChainW1Bytes = 6 * kChainW1Tile, kChainW2Bytes = 3 * kChainWTile;
constexpr int kChainWBytes = kChainW1Bytes + 2 * kChainW2Bytes;
constexpr int kChainThreads = 640;
constexpr int kChainGatherThreads = 256;
constexpr int kChainSmem = kChainWBytes + 2 * kChainRingBytes + 2 * kChainInStage + 3 * 64 + 256;
static_assert(kChainWBytes % 1024 == 0, "the line buffers follow the (1024-aligned) weight tiles");
static_assert((kChainRingPlane / 16) % 2 == 1 && (kChainInPlane / 16) % 2 == 1, "odd plane strides");
static_assert(kChainSmem <= 227 * 1024, "chain kernel exceeds the shared memory of an SM");

__device__ __forceinline__ void ccp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ uint64_t chain_desc_planar(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 |
         (uint64_t)1 << 46;
}
__device__ __forceinline__ void chain_warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ ChainArgs c) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  if ((raw & 1023u) != 0u) __trap();
  const uint32_t w_base = raw;
  const uint32_t a1_base = w_base + kChainWBytes;
  const uint32_t a2_base = a1_base + kChainRingBytes;
  const uint32_t in_base = a2_base + kChainRingBytes;
  const uint32_t bias_base = in_base + 2 * kChainInStage;
  const uint32_t bar_base = bias_base + 3 * 64;
  auto in_full = [&](int s) { return bar_base + 8u * s; };
  auto in_empty = [&](int s) { return bar_base + 16u + 8u * s; };
  const uint32_t w_full = bar_base + 32u;
  auto acc_full = [&](int k) { return bar_base + 40u + 8u * k; };
  auto acc_empty = [&](int k) { return bar_base + 64u + 8u * k; };
  auto a_ready = [&](int k) { return bar_base + 88u + 8u * k; };  // k = 0: conv1's rows in A1, k = 1: conv2's rows in A2
  const uint32_t tmem_slot = bar_base + 104u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_base - raw));  // [3][16]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = c.h, W = c.w, NG = c.groups;  // NG = H / 8

  if (threadIdx.x < 16) {
    bias_s[threadIdx.x] = c.b1[threadIdx.x];
    bias_s[16 + threadIdx.x] = c.b2[threadIdx.x];
    bias_s[32 + threadIdx.x] = c.head.bias[threadIdx.x];
  }
  // the permanent zero rows of both line buffers
  for (int i = threadIdx.x; i < 2 * 2 * kChainSlots; i += kChainThreads) {
    const int buf = i / (2 * kChainSlots), rem = i % (2 * kChainSlots);
    const int plane = rem / kChainSlots, slot = rem % kChainSlots;
    const uint32_t off = (buf ? a2_base : a1_base) - raw + plane * kChainRingPlane + kChainRing * kChainPitch + slot * 16;
    *reinterpret_cast<uint4*>(smem_raw + off) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&c.w1);
    tma_prefetch_desc(&c.w2);
    tma_prefetch_desc(&c.w3);
    for (int s = 0; s < 2; ++s) {
      mbar_init(in_full(s), kChainGatherThreads / 32);
      mbar_init(in_empty(s), 1);
    }
    mbar_init(w_full, 1);
    for (int k = 0; k < 3; ++k) {
      mbar_init(acc_full(k), 1);
      mbar_init(acc_empty(k), 8);
    }
    for (int k = 0; k < 2; ++k) mbar_init(a_ready(k), 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();  // the zero rows are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ weights: every tile once, resident
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kChainWBytes);
      uint32_t off = 0;
      for (int cb = 0; cb < 2; ++cb)
        for (int kx = 0; kx < 3; ++kx, off += kChainW1Tile) tma_load_2d(w_base + off, &c.w1, w_full, kx * 32 + cb * 16, 0);
      for (int kx = 0; kx < 3; ++kx, off += kChainWTile) tma_load_2d(w_base + off, &c.w2, w_full, kx * 16, 0);
      for (int kx = 0; kx < 3; ++kx, off += kChainWTile) tma_load_2d(w_base + off, &c.w3, w_full, kx * 16, 0);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc[5] = {0u, umma_idesc_f16(kTileM, 16, c.fp16), umma_idesc_f16(kTileM, 32, c.fp16),
                               umma_idesc_f16(kTileM, 48, c.fp16), umma_idesc_f16(kTileM, 64, c.fp16)};
    const uint64_t bdesc = umma_smem_desc<32>(w_base);
    const uint32_t b_lo0 = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
    const uint64_t d_in = chain_desc_planar(in_base, kChainInPlane, 128);
    const uint64_t d_a1 = chain_desc_planar(a1_base, kChainRingPlane, 128);
    const uint64_t d_a2 = chain_desc_planar(a2_base, kChainRingPlane, 128);
    const uint32_t in_lo = (uint32_t)d_in, in_hi = (uint32_t)(d_in >> 32);
    const uint32_t a1_lo = (uint32_t)d_a1, a1_hi = (uint32_t)(d_a1 >> 32);
    const uint32_t a2_lo = (uint32_t)d_a2, a2_hi = (uint32_t)(d_a2 >> 32);
    chain_warp_wait(w_full, 0, lane);
    operand_ready_fence();
    uint32_t n1 = 0, n2 = 0, n3 = 0;  // groups issued so far per layer (barrier phases)
    uint32_t r1 = 0, r2 = 0;          // line-buffer hand-offs waited for so far (conv1 -> A1, conv2 -> A2)
    // IU_CONV_DEBUG=1: role cycle counters in the slots of ConvArgs::debug ([0] MMA waits for a drained accumulator,
    // [1] for the gathered input, [2] for a line-buffer hand-off, [3] MMA warp total, [4] gather waits for a free stage,
    // [5] gather issue, [6] gather landing, [7] gather total, [8] epilogue waits for an accumulator, [9] epilogue body)
    unsigned long long* dbgp = c.head.debug;
    const bool dbg = dbgp != nullptr;
    long long w_acc = 0, w_in = 0, w_ready = 0, t0 = 0;
    const long long t_begin = dbg ? clock64() : 0;
#define IU_CHAIN_TIMED(acc_, stmt_)      \
  do {                                   \
    if (dbg) t0 = clock64();             \
    stmt_;                               \
    if (dbg) acc_ += clock64() - t0;     \
  } while (0)
    // conv2 / head of the group whose first OUTPUT row is `orow0`: input rows orow0-1 .. orow0+8 from the line buffer
    auto issue_folded = [&](uint32_t dbase, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, int orow0) {
      uint32_t a_row[kChainR + 2];  // descriptor of every input row: its ring row, or the zero row outside the image
#pragma unroll
      for (int jj = 0; jj < kChainR + 2; ++jj) {
        const int r_in = orow0 - 1 + jj;
        const int ring_row = (r_in < 0 || r_in >= H) ? kChainRing : r_in % kChainRing;
        a_row[jj] = a_lo + (uint32_t)((ring_row * kChainPitch) >> 4);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t b_k = b_lo + (uint32_t)kx * (kChainWTile >> 4);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
          for (int jj = 0; jj < kChainR + 2; ++jj) {
            // first touch of the accumulator (kx == 0): input rows 2, 5, 8 cover every output row exactly once and
            // overwrite; everything else accumulates
            const bool first = kx == 0 && (jj % 3 == 2);
            if (kx == 0 ? (first != (pass == 0)) : (pass == 1)) continue;
            const int lo_slot = jj >= 2 ? 0 : 2 - jj;
            const int hi_slot = jj <= kChainR - 1 ? 2 : kChainR + 1 - jj;
            umma_f16_lohi(dbase + (uint32_t)((jj - 2 + lo_slot) * 16), a_row[jj] + (uint32_t)kx, a_hi,
                          b_k + (uint32_t)((lo_slot * 16 * 32) >> 4), b_hi, idesc[hi_slot - lo_slot + 1],
                          first ? 0u : 1u);
          }
        }
      }
    };
    for (int item = blockIdx.x; item < c.total_items; item += gridDim.x) {
      for (int i = 0; i < NG + 3; ++i) {
        if (i < NG) {
          // ---- conv1 of group i: A1 rows [8i, 8i+8) from six source rows (four-slot pre-summed filters)
          IU_CHAIN_TIMED(w_acc, chain_warp_wait(acc_empty(0), (n1 & 1u) ^ 1u, lane));
          tc_fence_after();
          const uint32_t dbase = tmem_base;
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            IU_CHAIN_TIMED(w_in, chain_warp_wait(in_full(cb), n1 & 1u, lane));
            operand_ready_fence();
            if (elect_one()) {
              const uint32_t a_st = in_lo + (uint32_t)((cb * kChainInStage) >> 4);
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint32_t b_k = b_lo0 + (uint32_t)(((cb * 3 + kx) * kChainW1Tile) >> 4);
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                  for (int js = 0; js < kChainInRows; ++js) {
                    // first touch (chunk 0, kx 0): the odd source rows cover every output row exactly once
                    const bool first = cb == 0 && kx == 0 && (js & 1);
                    if ((cb == 0 && kx == 0) ? (first != (pass == 0)) : (pass == 1)) continue;
                    const int r_lo = 2 * js - 3;
                    const int lo_slot = r_lo < 0 ? -r_lo : 0;
                    const int hi_slot = r_lo + 3 > kChainR - 1 ? kChainR - 1 - r_lo : 3;
                    umma_f16_lohi(dbase + (uint32_t)((r_lo + lo_slot) * 16),
                                  a_st + (uint32_t)((js * kChainPitch) >> 4) + (uint32_t)kx, in_hi,
                                  b_k + (uint32_t)((lo_slot * 16 * 32) >> 4), b_hi, idesc[hi_slot - lo_slot + 1],
                                  first ? 0u : 1u);
                  }
                }
              }
              umma_commit(in_empty(cb));
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(acc_full(0));
          __syncwarp();
          ++n1;
        }
        const int j2 = i - 1;
        if (j2 >= 0 && j2 <= NG) {
          // ---- conv2 of group j2: A2 rows [8*j2 - 1, 8*j2 + 7) from A1 rows [8*j2 - 2, 8*j2 + 8)
          if (j2 < NG) IU_CHAIN_TIMED(w_ready, chain_warp_wait(a_ready(0), r1++ & 1u, lane));  // conv1's rows of group j2 are in A1
          IU_CHAIN_TIMED(w_acc, chain_warp_wait(acc_empty(1), (n2 & 1u) ^ 1u, lane));
          tc_fence_after();
          if (elect_one()) {
            issue_folded(tmem_base + 128u, a1_lo, a1_hi, b_lo0 + (uint32_t)(kChainW1Bytes >> 4), kChainR * j2 - 1);
            umma_commit(acc_full(1));
          }
          __syncwarp();
          ++n2;
        }
        const int j3 = i - 2;
        if (j3 >= 0 && j3 <= NG) {
          // ---- head of group j3: output rows [8*j3 - 2, 8*j3 + 6) from A2 rows [8*j3 - 3, 8*j3 + 7)
          IU_CHAIN_TIMED(w_ready, chain_warp_wait(a_ready(1), r2++ & 1u, lane));               // conv2's rows of group j3 are in A2
          IU_CHAIN_TIMED(w_acc, chain_warp_wait(acc_empty(2), (n3 & 1u) ^ 1u, lane));
          tc_fence_after();
          if (elect_one()) {
            issue_folded(tmem_base + 256u, a2_lo, a2_hi, b_lo0 + (uint32_t)((kChainW1Bytes + kChainW2Bytes) >> 4),
                         kChainR * j3 - 2);
            umma_commit(acc_full(2));
          }
          __syncwarp();
          ++n3;
        }
      }
    }
    if (dbg && lane == 0) {
      atomicAdd(dbgp + 0, (unsigned long long)w_acc);
      atomicAdd(dbgp + 1, (unsigned long long)w_in);
      atomicAdd(dbgp + 2, (unsigned long long)w_ready);
      atomicAdd(dbgp + 3, (unsigned long long)(clock64() - t_begin));
      atomicAdd(dbgp + 10, 1ull);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue: set s owns rows [4s, 4s+4) of every group
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;            // TMEM lanes [32q, 32q+32) = pixels [32q, 32q+32) of the 128-column tile
    const int p = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t n1 = 0, n2 = 0, n3 = 0;
    unsigned long long* dbgp = c.head.debug;
    const bool dbg = dbgp != nullptr && warp == 2 && lane == 0;
    long long e_wait = 0, e_t0 = 0;
    const long long e_begin = dbg ? clock64() : 0;
#define IU_CHAIN_EWAIT(stmt_)              \
  do {                                     \
    if (dbg) e_t0 = clock64();             \
    stmt_;                                 \
    if (dbg) e_wait += clock64() - e_t0;   \
  } while (0)
    // One layer's rows [4*set, 4*set+4) of the group in accumulator `acc_col`: TMEM -> bias + ReLU + 16-bit pack (two
    // rows at a time), accumulator handed back, then the packed rows go into the line buffer in the layout the next
    // layer's A descriptor reads.  Rows outside the image are skipped (readers take the zero row instead); columns
    // outside the image become zeros: the next conv's zero padding.
    auto rows_to_line_buffer = [&](uint32_t acc_col, uint32_t empty_bar, uint32_t ready_bar, const float* bias,
                                   uint32_t ring_base, int orow0, bool colok) {
      uint32_t o[4][8];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t acc[2][16];
#pragma unroll
        for (int q = 0; q < 2; ++q)
          tmem_ld_32x16(tmem_base + acc_col + lane_off + (uint32_t)((set * 4 + half * 2 + q) * 16), acc[q]);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 b2 = *reinterpret_cast<const float2*>(bias + 2 * k);
            const uint32_t v = relu16x2(pack16_sat(__uint_as_float(acc[q][2 * k]) + b2.x,
                                                   __uint_as_float(acc[q][2 * k + 1]) + b2.y, c.fp16), c.fp16);
            o[half * 2 + q][k] = colok ? v : 0u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int orow = orow0 + set * 4 + q;
        if (orow < 0 || orow >= H) continue;
        const uint32_t dst = ring_base + (uint32_t)((orow % kChainRing) * kChainPitch + (p + 1) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[q][0]), "r"(o[q][1]), "r"(o[q][2]), "r"(o[q][3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kChainRingPlane), "r"(o[q][4]), "r"(o[q][5]),
                     "r"(o[q][6]), "r"(o[q][7])
                     : "memory");
      }
      fence_proxy_async();  // this thread's line-buffer writes -> visible to the tensor core's (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar);
    };
    for (int item = blockIdx.x; item < c.total_items; item += gridDim.x) {
      const int strip = item % c.strips, n = item / c.strips;
      const int x = strip * kChainValid - 2 + p;     // this thread's image column in every layer
      const bool colok = x >= 0 && x < W;
      for (int i = 0; i < NG + 3; ++i) {
        if (i < NG) {
          IU_CHAIN_EWAIT(chain_warp_wait(acc_full(0), n1 & 1u, lane));
          tc_fence_after();
          rows_to_line_buffer(0u, acc_empty(0), a_ready(0), bias_s, a1_base, kChainR * i, colok);
          ++n1;
        }
        const int j2 = i - 1;
        if (j2 >= 0 && j2 <= NG) {
          IU_CHAIN_EWAIT(chain_warp_wait(acc_full(1), n2 & 1u, lane));
          tc_fence_after();
          rows_to_line_buffer(128u, acc_empty(1), a_ready(1), bias_s + 16, a2_base, kChainR * j2 - 1, colok);
          ++n2;
        }
        const int j3 = i - 2;
        if (j3 >= 0 && j3 <= NG) {
          IU_CHAIN_EWAIT(chain_warp_wait(acc_full(2), n3 & 1u, lane));
          tc_fence_after();
          const bool mine = colok && p >= 2 && p < 2 + kChainValid;
          if (c.head.num_classes <= 4) {
            uint32_t acc[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_ld_32x4(tmem_base + 256u + lane_off + (uint32_t)((set * 4 + q) * 16), acc[q]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(2));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int y = kChainR * j3 - 2 + set * 4 + q;
              if (mine && y >= 0 && y < H) {
                if (c.head.num_classes == 2) softmax_store<2>(c.head, bias_s + 32, acc[q], n, y, x);
                else if (c.head.num_classes == 4) softmax_store<4>(c.head, bias_s + 32, acc[q], n, y, x);
                else if (c.head.num_classes == 3) softmax_store<3>(c.head, bias_s + 32, acc[q], n, y, x);
                else softmax_store<1>(c.head, bias_s + 32, acc[q], n, y, x);
              }
            }
          } else {
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
              const int y = kChainR * j3 - 2 + set * 4 + q;
              uint4 unused[EpiCfg<16>::RV];
              epilogue_pixel<16>(c.head, bias_s + 32, 0, tmem_base + 256u + lane_off + (uint32_t)((set * 4 + q) * 16), n, y, x,
                                 mine && y >= 0 && y < H, unused);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(2));
          }
          ++n3;
        }
      }
    }
    if (dbg) {
      atomicAdd(dbgp + 8, (unsigned long long)e_wait);
      atomicAdd(dbgp + 9, (unsigned long long)(clock64() - e_begin - e_wait));
    }
  } else if ((warp & 3) != 1) {
    // ------------------------------------------------------------ gather: six source rows x 130 slots x 16 channels per stage
    const int gw = warp - 10;
    const int t = (gw - (gw + 1) / 4) * 32 + lane;   // rank among the gather warps (ids 13 and 17 stay idle) * 32 + lane
    const int sh = H >> 1, sw = W >> 1;
    uint32_t n1 = 0;
    unsigned long long* dbgp = c.head.debug;
    const bool dbg = dbgp != nullptr && t == 0;
    long long g_empty = 0, g_issue = 0, g_land = 0, g_t0 = 0;
    const long long g_begin = dbg ? clock64() : 0;
    // Thread t owns the (slot, plane) column t of a stage (130 slots x 2 planes = 260 columns: the last four are shared
    // out as 24 single copies) and walks the six rows: per copy one bounds test and one pointer add.
    const int my_slot = t >> 1, my_plane = t & 1;
    const int x_slot = 128 + ((t % 4) >> 1), x_plane = t & 1, x_row = t >> 2;   // threads 0..23: slots 128 / 129, rows 0..5
    const bool has_extra = t < 4 * kChainInRows;
    const uint32_t my_dst = my_plane * kChainInPlane + my_slot * 16;
    const uint32_t x_dst = x_plane * kChainInPlane + x_row * kChainPitch + x_slot * 16;
    for (int item = blockIdx.x; item < c.total_items; item += gridDim.x) {
      const int strip = item % c.strips, n = item / c.strips;
      const int ux0 = strip * kChainValid - 3;         // upsampled column of slot 0
      const __nv_bfloat16* img = c.src + (size_t)n * sh * sw * 32;
      const int my_ux = ux0 + my_slot, x_ux = ux0 + x_slot;
      const bool my_xok = (unsigned)my_ux < (unsigned)W, x_xok = (unsigned)x_ux < (unsigned)W;
      const __nv_bfloat16* my_col = img + (size_t)(my_xok ? my_ux >> 1 : 0) * 32 + my_plane * 8;
      const __nv_bfloat16* x_col = img + (size_t)(x_xok ? x_ux >> 1 : 0) * 32 + x_plane * 8;
      const size_t row_stride = (size_t)sw * 32;
      for (int i = 0; i < NG; ++i, ++n1) {
        const int sy0 = 4 * i - 1;                     // source row of stage row 0
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          if (dbg) g_t0 = clock64();
          chain_warp_wait(in_empty(cb), (n1 & 1u) ^ 1u, lane);
          if (dbg) { const long long t1 = clock64(); g_empty += t1 - g_t0; g_t0 = t1; }
          const uint32_t stage = in_base + cb * kChainInStage;
#pragma unroll
          for (int js = 0; js < kChainInRows; ++js) {
            const int sy = sy0 + js;
            const bool ok = my_xok && (unsigned)sy < (unsigned)sh;
            ccp_async_16(stage + my_dst + js * kChainPitch, ok ? my_col + (size_t)sy * row_stride + cb * 16 : c.src, ok ? 16u : 0u);
          }
          if (has_extra) {
            const int sy = sy0 + x_row;
            const bool ok = x_xok && (unsigned)sy < (unsigned)sh;
            ccp_async_16(stage + x_dst, ok ? x_col + (size_t)sy * row_stride + cb * 16 : c.src, ok ? 16u : 0u);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (dbg) g_issue += clock64() - g_t0;
        }
        if (dbg) g_t0 = clock64();
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          fence_proxy_async();
          mbar_arrive(in_full(0));
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          fence_proxy_async();
          mbar_arrive(in_full(1));
        }
        if (dbg) g_land += clock64() - g_t0;
      }
    }
    if (dbg) {
      atomicAdd(dbgp + 4, (unsigned long long)g_empty);
      atomicAdd(dbgp + 5, (unsigned long long)g_issue);
      atomicAdd(dbgp + 6, (unsigned long long)g_land);
      atomicAdd(dbgp + 7, (unsigned long long)(clock64() - g_begin));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------- host side
// The strip overhead is (strips * 128) / width: 1.25 at 512 columns, 1.125 at 1024, 1.06 at 2048, 1.5 at 256.  Against
// it stand the four 32-byte-per-pixel round trips the fusion removes; below ~500 columns the three separate launches win.
bool conv_chain_applicable(int h, int w) {
  if (h < 64 || h % kChainR || w < 128 || (w & 1) || (h & 1)) return false;
  const int strips = (w + kChainValid - 1) / kChainValid;
  return strips * 128 * 10 <= w * 13;
}

cudaError_t launch_conv_chain(const ChainArgs& args_in, cudaStream_t stream) {
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured_dev = dev;
  }
  ChainArgs args = args_in;
  args.strips = (args.w + kChainValid - 1) / kChainValid;
  args.total_items = args.strips * args.batch;
  args.groups = args.h / kChainR;
  args.head.out_h = args.h;
  args.head.out_w = args.w;
  const int grid = args.total_items < num_sms ? args.total_items : num_sms;
  conv_chain_kernel<<<grid, kChainThreads, kChainSmem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace iu
