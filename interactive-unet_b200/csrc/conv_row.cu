// Row-folded implicit-GEMM 3x3 convolution for the narrow layers (Cout = 16 / 32 / 64, image width >= 128).
//
// Why: a tcgen05.mma with M = 128 and both operands in shared memory costs ~42 cycles however small N is (it has
// to read the 4 KB A tile), so an N = 16 / 32 / 64 MMA runs at 19 / 38 / 66 % of the tensor rate
// (profiles/r01_pipe_probe.txt).  The narrow layers -- layer1, decoder blocks 2-4 and the head, 31 % of the
// network's FLOPs -- were therefore issue bound at a fraction of peak.  This kernel makes every MMA three filter
// taps wide:
//   * an M tile is 128 CONSECUTIVE PIXELS OF ONE IMAGE ROW; a CTA block is R output rows x 128 columns;
//   * the accumulators of the block's R output rows sit side by side in TMEM: row r owns columns [r*CO, (r+1)*CO);
//   * the weights of one filter column kx are packed [W(ky=2) | W(ky=1) | W(ky=0)] (3*CO rows of B), so ONE MMA of
//     input row i against that tile adds, in a single N = 3*CO instruction, input row i's contribution to output
//     rows i-1, i and i+1 -- the three vertical taps land in three neighbouring accumulators because those are
//     contiguous TMEM columns.  The block's first / last input rows use the matching 1- or 2-slot sub-tiles.
//     Per (kx, 16 channels): R+2 MMAs of N <= 3*CO instead of 3*R MMAs of N = CO.
//   * the A operand is the planar halo layout of conv_halo.cu (plane = 8 channels, 16 B per pixel) with a row pitch
//     of 130 pixels: input row i, filter column kx is the same buffer through a descriptor shifted by
//     i * pitch + kx * 16 bytes (SBO = 128 B: the eight-pixel core matrices of a row are contiguous).  The 2x nearest
//     upsample of the decoder (`src = dst >> 1`) and the channel concat are folded into the gather as before;
//   * Cout-64 layers fed by ONE identity 64-channel tensor (layer1, decoder block 2 conv2) fill the A ring by TMA instead:
//     a stage is three input rows x 130 pixels in the 128B-swizzled K-major layout (one pixel = one 128-byte row),
//     and input row j / filter column kx is the stage read through a descriptor whose START ADDRESS is shifted by
//     (j * 130 + kx) * 128 bytes -- the tensor core applies the swizzle to address bits, so an operand may begin at
//     any 128-byte row of a swizzled buffer (base offset field 0; verified bit for bit against the planar path).
//     One thread issues two 50 KB boxes per block where eight warps issued 6400 cp.async: the MMA issuer's wait for
//     operands drops from 60k to 11k cycles per CTA and pass, the layer from 184k to 110k (RowCfg TMA_A);
//   * an UPSAMPLED source (decoder conv1: `F.interpolate(x, 2x nearest)`) is never expanded vertically: the block
//     gathers its R/2+2 SOURCE rows (each pixel still written twice along x), and because upsampled rows 2s and 2s+1
//     are the same data, source row s feeds output rows 2s-1 .. 2s+2 with the pre-summed weights
//     [W(ky=2) | W(1)+W(2) | W(0)+W(1) | W(ky=0)]: one N = 4*CO MMA per source row instead of two N = 3*CO MMAs, and
//     half the gather traffic (the gather's cp.async rate, ~20 B/clk/SM, is what bounds those layers);
//   * a residual (ResNet shortcut) is one more K segment against an identity weight tile: the tensor core adds it,
//     the epilogue never touches global memory for it;
//   * epilogue: TMEM -> registers -> bias / ReLU / 16-bit pack -> swizzled shared-memory staging -> ONE TMA store
//     of the 128-pixel row (a contiguous 128*CO*2-byte run of the NHWC tensor).  The per-lane 16-byte global
//     stores of the generic epilogue cost 32 L1 wavefronts per instruction and were the bottleneck of layer1.
//   * all weight tiles of the layer stay resident in shared memory (loaded once per CTA by TMA).
// Warp roles (640 threads, one persistent CTA per SM): warp 0 weight TMA, warp 1 MMA issue, warps 2-9 epilogue
// (two groups of four, each owning half of the block's rows), eight of the warps 10-19 gather (cp.async).
//
// Replaces the same cuDNN conv2d / batch_norm / relu / add / cat / upsample / softmax launches as conv_tc.cu
// (`smp.Unet.forward` under `/root/reference/interactive_unet/unet.py:67`).
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

constexpr int kRowSeg = 128;               // output pixels per M tile (one image-row segment)
constexpr int kRowHaloPx = kRowSeg + 2;    // gathered pixels per row: x0-1 .. x0+128
constexpr int kRowPitch = kRowHaloPx * 16;  // bytes per gathered row inside a plane
constexpr int kRowGatherThreads = 256;     // 8 gather warps
// Warps are bound to the SM's four schedulers by (warp id % 4).  The MMA issuer is warp 1; the gather warps are the
// ids >= 10 with id % 4 != 1 (10,11,12,14,15,16,18,19), ids 13 and 17 idle, so that the issuer's scheduler hosts only
// itself and two epilogue warps (TMEM lane quarter 1 needs them there): tools/contention_probe.cu shows ALU-busy
// warps on the issuer's scheduler slowing a small-N MMA stream by 25 %.
constexpr int kRowThreads = 640;
constexpr int kRowSmemMax = 227 * 1024;

// KC : channels per gathered A stage          KCB: channels per weight tile (its TMA / UMMA swizzle span is KCB*2 bytes;
//      a 64-byte span is read by the tensor core with 2-way bank conflicts, so the 64-output layers use 128-byte tiles)
// CO : output channels (= cout_pad)           R  : output rows per block         STAGES: A ring depth
// RB : output rows per TMA store            NSTG: staging buffers (of RB rows) per epilogue group (1 or 2).
//      Direct 16-byte global stores from the epilogue lanes were measured SLOWER even for 32 / 64 bytes per pixel
//      (dec3/dec4 layers +7..14 %): they share the LSU / L1 wavefront queue with the gather's cp.async traffic,
//      which is the scarcer resource; the TMA store bypasses it.
// TMA_A: the A ring is filled by TMA in the 128B-swizzled K-major layout (one 64-channel pixel = one 128-byte row):
//      a stage is TROWS = 3 input rows x 130 pixels, a block's R + 2 input rows are (R + 2) / 3 stages, and input row
//      j, filter column kx is the stage read through a descriptor whose start address is shifted by
//      (j * 130 + kx) * 128 bytes (the tensor core swizzles on address bits, so any 128-byte row may start an operand).
//      Only identity (not upsampled) 64-channel sources; the gather warps idle.
template <int KC, int KCB, int CO, int R, int STAGES, int RB, int NSTG, bool STREAM, bool TMA_A = false>
struct RowCfg {
  static constexpr int PLANES = KC / 8;
  static constexpr int ROWS = R + 2;
  static constexpr int PLANE_STRIDE = ROWS * kRowPitch + 16;  // (stride / 16) odd: planes start 16 B apart mod 32 banks
  static constexpr int TROWS = 3;                              // TMA_A: input rows per stage
  static constexpr int TSUB = ROWS / TROWS;                    //        stages per block
  static constexpr int T_ROW = kRowHaloPx * KC * 2;            //        bytes per input row in a stage (130 x 128)
  static constexpr int T_STAGE = TROWS * T_ROW;                //        bytes one TMA box lands
  // no-swizzle operand: 16-byte alignment is enough; the swizzled TMA stages stay 1024-aligned
  static constexpr int A_STAGE = TMA_A ? (T_STAGE + 1023) / 1024 * 1024 : PLANES * PLANE_STRIDE;
  static_assert(!TMA_A || (KC == 64 && KCB == 64 && !STREAM && ROWS % TROWS == 0), "TMA-filled A ring: 64-channel chunks");
  static constexpr int A_STAGES = STAGES;
  static constexpr int NF = 3 * CO;         // N of a full-width (three-slot) MMA
  static constexpr int SWB = KCB * 2;       // bytes per weight row = TMA / UMMA swizzle span
  static constexpr int B_TILE = NF * SWB;   // one (segment, KCB chunk, kx) weight tile
  static constexpr int I_TILE = CO * SWB;   // one identity tile (residual segment)
  static constexpr int U_TILE = 4 * CO * SWB;  // one weight tile of an upsampled segment (four slots, see below)
  static constexpr int ROWS_UP = R / 2 + 2;    // source rows an upsampled segment gathers per block
  static constexpr int STG = RB * kRowSeg * CO * 2;  // one staging buffer: RB output row segments
  static constexpr int STG_TOTAL = 2 * NSTG * STG;   // two epilogue groups x NSTG buffers
  static constexpr int ACC_COLS = R * CO;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;  // double buffered: 512 / 512 / 256
  static constexpr int MISC = 512;  // bias (<= 256 B) + barriers + TMEM slot
  // layout: [weights W_MAX][staging][A ring][bias, barriers]; the swizzled regions come first (1024-aligned base)
  // STREAM: the layer's weights do not fit in shared memory (decoder block 2 conv1: 264 KB); the three kx tiles of
  // each A chunk travel WITH that chunk -- same ring slot, same full barrier (TMA complete_tx next to the gather
  // warps' arrivals) -- instead of staying resident.  Needs KCB == KC.
  static constexpr int A_PAD = (A_STAGE + 255) / 256 * 256;          // weight tiles start 256-byte aligned
  static constexpr int W_STAGE = STREAM ? 3 * U_TILE : 0;
  static constexpr int STAGE_BYTES = STREAM ? A_PAD + W_STAGE : A_STAGE;
  static constexpr int W_MAX = STREAM ? 0 : (kRowSmemMax - STG_TOTAL - A_STAGES * A_STAGE - MISC) / 1024 * 1024;
  static constexpr int SMEM_BYTES = W_MAX + STG_TOTAL + A_STAGES * STAGE_BYTES + MISC;
  static_assert(!STREAM || KCB == KC, "streamed weight tiles are one A chunk wide");
  static constexpr int CH = CO >= 32 ? 32 : 16;  // accumulator columns per tcgen05.ld
  static_assert(NSTG == 1 || NSTG == 2, "one or two staging buffers per epilogue group");
  static_assert(KCB % KC == 0 && (KCB == 64 || KCB == 32 || KCB == 16), "weight tile width");
  static_assert((PLANE_STRIDE / 16) % 2 == 1, "plane stride must be an odd number of 16-byte units");
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
  static_assert((STREAM || W_MAX > 0) && SMEM_BYTES <= kRowSmemMax, "shared memory budget");
  static_assert((R / 2) % RB == 0, "each epilogue group stores whole RB-row boxes");
  static_assert((kRowSeg * PLANES) % kRowGatherThreads == 0, "gather columns per thread");
};

__device__ __forceinline__ void rcp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void rcp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rcp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint64_t row_desc_planar(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 |
         (uint64_t)1 << 46;
}
__device__ __forceinline__ void row_warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}
struct RowTile {
  int n, y0, x0;
};
__device__ __forceinline__ RowTile row_decode(const ConvArgs& a, int tile, int rows_per_block) {
  RowTile t;
  t.x0 = (tile % a.tiles_x) * kRowSeg;
  const int m = tile / a.tiles_x;
  t.y0 = (m % a.tiles_y) * rows_per_block;
  t.n = m / a.tiles_y;
  return t;
}

template <int KC, int KCB, int CO, int R, int STAGES, int RB, int NSTG, bool STREAM, bool TMA_A = false>
__global__ void __launch_bounds__(kRowThreads, 1) conv_row_kernel(const __grid_constant__ ConvArgs a) {
  using Cfg = RowCfg<KC, KCB, CO, R, STAGES, RB, NSTG, STREAM, TMA_A>;
  const long long t_cta = (a.debug != nullptr && threadIdx.x == 0) ? clock64() : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  if ((raw & 1023u) != 0u) __trap();  // the swizzled regions rely on the 1024-byte alignment of the dynamic window
  const uint32_t w_base = raw;
  const uint32_t stg_base = w_base + Cfg::W_MAX;
  const uint32_t a_base = stg_base + Cfg::STG_TOTAL;
  const uint32_t bias_base = a_base + Cfg::A_STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = bias_base + 256;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::A_STAGES + s); };
  const uint32_t b_full = bar_base + 16u * Cfg::A_STAGES;
  auto acc_full = [&](int b) { return b_full + 8u + 8u * b; };
  auto acc_empty = [&](int b) { return b_full + 24u + 8u * b; };
  const uint32_t tmem_slot = b_full + 40u;
  // staging hand-off between an epilogue group and its store thread: full (4 warp arrivals) / empty (store thread)
  auto stg_full = [&](int g, int b) { return b_full + 48u + 8u * (g * 2 + b); };
  auto stg_empty = [&](int g, int b) { return b_full + 80u + 8u * (g * 2 + b); };
  auto res_full = [&](int g) { return b_full + 112u + 8u * g; };  // residual row landed in group g's staging buffer
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_base - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // The residual (ResNet shortcut) is added either by the tensor core -- one more K segment against an identity tile,
  // gathered like any other source -- or, when `res_tma` is set, by the epilogue: the store thread TMA-loads the
  // residual row into the output staging buffer (same box, same swizzle as the store), the epilogue adds it to the
  // accumulator in fp32 and overwrites it in place.  The second form halves the gather volume and drops 4*R MMAs per
  // block of the Cout-64 layers (layer1's conv2: 172 -> conv1's ~120 us per 74-slice pass).
  const bool res_tma = a.residual != nullptr && a.res_tma != 0;
  const bool has_res = !TMA_A && a.residual != nullptr && !res_tma;
  constexpr int RES_CHUNKS = CO / KC;    // A stages of the identity (residual) segment
  constexpr int RES_BTILES = CO / KCB;   // its weight tiles
  constexpr int CPB = KCB / KC;          // A chunks per weight tile

  if (threadIdx.x < CO) bias_s[threadIdx.x] = a.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.bmapf);
    if (has_res) tma_prefetch_desc(&a.bmapi);
    if (a.mode == kEpiBf16) tma_prefetch_desc(&a.omap);
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      // gather warps (+ the weight TMA's expect_tx), or the A TMA's expect_tx alone
      mbar_init(a_full(s), TMA_A ? 1 : kRowGatherThreads / 32 + (STREAM ? 1 : 0));
      mbar_init(a_empty(s), 1);
    }
    mbar_init(b_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 8);
    }
    for (int g = 0; g < 2; ++g)
      for (int b = 0; b < 2; ++b) {
        mbar_init(stg_full(g, b), 4);
        mbar_init(stg_empty(g, b), 1);
      }
    for (int g = 0; g < 2; ++g) mbar_init(res_full(g), 1);
    if (res_tma) tma_prefetch_desc(&a.rmap);
    if (TMA_A) tma_prefetch_desc(&a.rowmap);
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

  if (warp == 0 && STREAM) {
    // ------------------------------------------------------------ weights streamed with their A chunk
    const int up0 = a.seg[0].up;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      int kbase = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int cin = a.seg[s].cin;
        const bool upseg = s == 0 && up0;
        const uint32_t tb = upseg ? Cfg::U_TILE : Cfg::B_TILE;
        for (int cc = 0; cc < cin / KC; ++cc, ++it) {
          const int st = it % Cfg::A_STAGES;
          row_warp_wait(a_empty(st), ((it / Cfg::A_STAGES) & 1) ^ 1u, lane);
          if (elect_one()) {
            const uint32_t wdst = a_base + st * Cfg::STAGE_BYTES + Cfg::A_PAD;
            mbar_arrive_expect_tx(a_full(st), 3 * tb);
            for (int kx = 0; kx < 3; ++kx) {
              if (upseg) tma_load_2d(wdst + kx * tb, &a.bmapu, a_full(st), kx * cin + cc * KC, 0);
              else tma_load_2d(wdst + kx * tb, &a.bmapf, a_full(st), kbase + kx * cin + cc * KC, 0);
            }
          }
          __syncwarp();
        }
        kbase += 3 * cin;
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------ weights: every tile once, resident for the CTA's life
    if (elect_one()) {
      // segment 0 may be an upsampled source (four-slot tiles from `bmapu`); the others use the three-slot tiles
      const int up0 = a.seg[0].up;
      uint32_t wbytes = 0;
      for (int s = 0; s < a.nseg; ++s) wbytes += 3 * (a.seg[s].cin / KCB) * ((s == 0 && up0) ? Cfg::U_TILE : Cfg::B_TILE);
      mbar_arrive_expect_tx(b_full, wbytes + (has_res ? RES_BTILES * Cfg::I_TILE : 0));
      int kbase = 0;
      uint32_t off = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int cin = a.seg[s].cin;
        const bool upseg = s == 0 && up0;
        for (int cb = 0; cb < cin / KCB; ++cb)
          for (int kx = 0; kx < 3; ++kx) {
            if (upseg) tma_load_2d(w_base + off, &a.bmapu, b_full, kx * cin + cb * KCB, 0);
            else tma_load_2d(w_base + off, &a.bmapf, b_full, kbase + kx * cin + cb * KCB, 0);
            off += upseg ? Cfg::U_TILE : Cfg::B_TILE;
          }
        kbase += 3 * cin;
      }
      if (has_res)
        for (int cb = 0; cb < RES_BTILES; ++cb)
          tma_load_2d(w_base + off + cb * Cfg::I_TILE, &a.bmapi, b_full, cb * KCB, 0);
      if constexpr (TMA_A) {
        // ---------------------------------------------------------- A ring by TMA: three input rows per stage
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
          const RowTile tc = row_decode(a, tile, R);
          for (int hs = 0; hs < Cfg::TSUB; ++hs, ++it) {
            const int st = it % Cfg::A_STAGES;
            mbar_wait(a_empty(st), ((it / Cfg::A_STAGES) & 1) ^ 1u);
            mbar_arrive_expect_tx(a_full(st), Cfg::T_STAGE);
            // pixels x0-1 .. x0+128, rows y0-1+3hs .. +2: whatever lies outside the image arrives as zeros
            tma_load_4d(a_base + st * Cfg::STAGE_BYTES, &a.rowmap, a_full(st), 0, tc.x0 - 1,
                        tc.y0 - 1 + Cfg::TROWS * hs, tc.n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const uint32_t idesc1 = umma_idesc_f16(kTileM, CO, a.fp16);
    const uint32_t idesc2 = umma_idesc_f16(kTileM, 2 * CO, a.fp16);
    const uint32_t idesc3 = umma_idesc_f16(kTileM, 3 * CO, a.fp16);
    const uint64_t bdesc_base = umma_smem_desc<Cfg::SWB>(w_base);
    const uint32_t b_lo_base = (uint32_t)bdesc_base, b_hi = (uint32_t)(bdesc_base >> 32);
    const uint32_t idesc4 = umma_idesc_f16(kTileM, 4 * CO, a.fp16);
    const int up0 = a.seg[0].up;
    uint32_t wbytes = 0;  // bytes of all conv weight tiles = offset of the identity tiles
    for (int s = 0; s < a.nseg; ++s) wbytes += 3 * (a.seg[s].cin / KCB) * ((s == 0 && up0) ? Cfg::U_TILE : Cfg::B_TILE);
    const bool dbg = a.debug != nullptr;
    long long w_acc = 0, w_a = 0, w_b = 0, t_begin = dbg ? clock64() : 0, t0 = 0;
    if constexpr (!STREAM) {
      row_warp_wait(b_full, 0, lane);
      if (dbg) w_b = clock64() - t_begin;
      operand_ready_fence();
    }
    uint32_t ita = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1u;
      if (dbg) t0 = clock64();
      row_warp_wait(acc_empty(buf), ((tcount >> 1) & 1u) ^ 1u, lane);
      if (dbg) w_acc += clock64() - t0;
      tc_fence_after();
      const uint32_t dbase = tmem_base + buf * Cfg::ACC_COLS;
      // input row j (0 .. R+1) of the block against the kx tile at b_lo: the slots of [W2|W1|W0] whose output row
      // (j - 2 + slot) lies inside the block
      auto issue_row = [&](int j, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t accumulate) {
        const int lo_slot = j >= 2 ? 0 : 2 - j;
        const int hi_slot = j <= R - 1 ? 2 : R + 1 - j;
        const int nslots = hi_slot - lo_slot + 1;
        const uint32_t idesc = nslots == 3 ? idesc3 : (nslots == 2 ? idesc2 : idesc1);
        umma_f16_lohi(dbase + (uint32_t)((j - 2 + lo_slot) * CO), a_lo + (uint32_t)((j * kRowPitch) >> 4), a_hi,
                      b_lo + (uint32_t)((lo_slot * CO * Cfg::SWB) >> 4), b_hi, idesc, accumulate);
      };
      // source row js (0 .. R/2+1) of an upsampled segment against its four-slot tile: output rows 2js-3 .. 2js
      auto issue_up = [&](int js, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t accumulate) {
        const int r_lo = 2 * js - 3;
        const int lo_slot = r_lo < 0 ? -r_lo : 0;
        const int hi_slot = r_lo + 3 > R - 1 ? R - 1 - r_lo : 3;
        const int nslots = hi_slot - lo_slot + 1;
        const uint32_t idesc = nslots == 4 ? idesc4 : (nslots == 3 ? idesc3 : (nslots == 2 ? idesc2 : idesc1));
        umma_f16_lohi(dbase + (uint32_t)((r_lo + lo_slot) * CO), a_lo + (uint32_t)((js * kRowPitch) >> 4), a_hi,
                      b_lo + (uint32_t)((lo_slot * CO * Cfg::SWB) >> 4), b_hi, idesc, accumulate);
      };
      if constexpr (TMA_A) {
        // one 64-channel chunk per block, in TSUB stages of three input rows.  Stage hs holds block rows 3hs .. 3hs+2;
        // its last row is the first to touch the output rows no earlier stage has written (3hs .. 3hs+2 minus the
        // block's edges), so it goes first and overwrites; everything after accumulates.
        for (int hs = 0; hs < Cfg::TSUB; ++hs, ++ita) {
          const int sta = ita % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          row_warp_wait(a_full(sta), (ita / Cfg::A_STAGES) & 1, lane);
          if (dbg) w_a += clock64() - t0;
          operand_ready_fence();
          const uint32_t a_addr = a_base + sta * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_smem_desc<128>(a_addr);
          const uint32_t a_lo = (uint32_t)adesc, a_hi0 = (uint32_t)(adesc >> 32);
          if (elect_one()) {
            auto issue_t = [&](int jj, uint32_t kx, uint32_t kk, uint32_t accumulate) {
              const int j = Cfg::TROWS * hs + jj;
              const int lo_slot = j >= 2 ? 0 : 2 - j;
              const int hi_slot = j <= R - 1 ? 2 : R + 1 - j;
              const int nslots = hi_slot - lo_slot + 1;
              const uint32_t idesc = nslots == 3 ? idesc3 : (nslots == 2 ? idesc2 : idesc1);
              const uint32_t off = (uint32_t)(jj * Cfg::T_ROW) + kx * 128u + kk * 32u;  // bytes into the stage
              // development switch (row_tma == 2): matrix base offset = 128-byte row phase of the start address
              const uint32_t a_hi = a.row_tma == 2 ? (a_hi0 | ((((a_addr + off) >> 7) & 7u) << 17)) : a_hi0;
              umma_f16_lohi(dbase + (uint32_t)((j - 2 + lo_slot) * CO), a_lo + (off >> 4), a_hi,
                            b_lo_base + kx * (Cfg::B_TILE >> 4) + 2u * kk + (uint32_t)((lo_slot * CO * Cfg::SWB) >> 4), b_hi,
                            idesc, accumulate);
            };
#pragma unroll
            for (uint32_t kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (uint32_t kk = 0; kk < KC / 16; ++kk) {
                issue_t(Cfg::TROWS - 1, kx, kk, (kx | kk) ? 1u : 0u);
#pragma unroll
                for (int jj = 0; jj < Cfg::TROWS - 1; ++jj) issue_t(jj, kx, kk, 1u);
              }
            }
            umma_commit(a_empty(sta));
          }
          __syncwarp();
        }
      }
      uint32_t chunk = 0;
      uint32_t tile_off = 0;  // byte offset of the current segment's first weight tile
      for (int s = 0; s < (TMA_A ? 0 : a.nseg); ++s) {
        const bool upseg = s == 0 && up0;
        const uint32_t tile_bytes = upseg ? Cfg::U_TILE : Cfg::B_TILE;
        for (int cc = 0; cc < a.seg[s].cin / KC; ++cc, ++ita, ++chunk) {
          const int sta = ita % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          row_warp_wait(a_full(sta), (ita / Cfg::A_STAGES) & 1, lane);
          if (dbg) w_a += clock64() - t0;
          operand_ready_fence();
          const uint64_t adesc = row_desc_planar(a_base + sta * Cfg::STAGE_BYTES, Cfg::PLANE_STRIDE, 128);
          const uint32_t a_lo = (uint32_t)adesc, a_hi = (uint32_t)(adesc >> 32);
          // weight tile of this chunk: tiles are KCB wide, chunk `chunk` sits (chunk % CPB) * KC channels into its tile
          const uint32_t b_chunk =
              STREAM ? (uint32_t)umma_smem_desc<Cfg::SWB>(a_base + sta * Cfg::STAGE_BYTES + Cfg::A_PAD)
                     : b_lo_base + ((tile_off + (uint32_t)(cc / CPB) * 3u * tile_bytes) >> 4) + (uint32_t)(cc % CPB) * (KC / 8u);
          if (upseg) {
            if (elect_one()) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_k = a_lo + (uint32_t)kx + (uint32_t)(2 * kk) * (Cfg::PLANE_STRIDE >> 4);
                  const uint32_t b_k = b_chunk + (uint32_t)kx * (Cfg::U_TILE >> 4) + 2u * kk;
                  if (kx == 0 && kk == 0 && chunk == 0) {
                    // first touch: the odd source rows cover every output row exactly once (windows of four rows, two apart)
#pragma unroll
                    for (int js = 1; js < Cfg::ROWS_UP; js += 2) issue_up(js, a_k, a_hi, b_k, 0u);
#pragma unroll
                    for (int js = 0; js < Cfg::ROWS_UP; js += 2) issue_up(js, a_k, a_hi, b_k, 1u);
                  } else {
#pragma unroll
                    for (int js = 0; js < Cfg::ROWS_UP; ++js) issue_up(js, a_k, a_hi, b_k, 1u);
                  }
                }
              }
              umma_commit(a_empty(sta));
            }
            __syncwarp();
            continue;
          }
          if (elect_one()) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint32_t a_k = a_lo + (uint32_t)kx + (uint32_t)(2 * kk) * (Cfg::PLANE_STRIDE >> 4);
                const uint32_t b_k = b_chunk + (uint32_t)kx * (Cfg::B_TILE >> 4) + 2u * kk;
                if (kx == 0 && kk == 0 && chunk == 0) {
                  // first touch of this accumulator buffer: the input rows j = 2, 5, 8, ... cover every output row
                  // exactly once, so they overwrite (accumulate = 0); everything after accumulates
#pragma unroll
                  for (int j = 2; j < R + 2; j += 3) issue_row(j, a_k, a_hi, b_k, 0u);
#pragma unroll
                  for (int j = 0; j < R + 2; ++j)
                    if (j % 3 != 2) issue_row(j, a_k, a_hi, b_k, 1u);
                } else {
#pragma unroll
                  for (int j = 0; j < R + 2; ++j) issue_row(j, a_k, a_hi, b_k, 1u);
                }
              }
            }
            umma_commit(a_empty(sta));
          }
          __syncwarp();
        }
        tile_off += 3u * (uint32_t)(a.seg[s].cin / KCB) * tile_bytes;
      }
      if (has_res) {
        // residual: out row r += I * residual row r (gathered like any other source: block row r+1, centre column)
        for (int cc = 0; cc < RES_CHUNKS; ++cc, ++ita) {
          const int sta = ita % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          row_warp_wait(a_full(sta), (ita / Cfg::A_STAGES) & 1, lane);
          if (dbg) w_a += clock64() - t0;
          operand_ready_fence();
          const uint64_t adesc = row_desc_planar(a_base + sta * Cfg::STAGE_BYTES, Cfg::PLANE_STRIDE, 128);
          const uint32_t a_lo = (uint32_t)adesc, a_hi = (uint32_t)(adesc >> 32);
          const uint32_t b_i = b_lo_base + (uint32_t)((wbytes + (cc / CPB) * Cfg::I_TILE) >> 4) +
                               (uint32_t)(cc % CPB) * (KC / 8u);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk) {
#pragma unroll
              for (int r = 0; r < R; ++r)
                umma_f16_lohi(dbase + (uint32_t)(r * CO),
                              a_lo + (uint32_t)(((r + 1) * kRowPitch) >> 4) + 1u +
                                  (uint32_t)(2 * kk) * (Cfg::PLANE_STRIDE >> 4),
                              a_hi, b_i + 2u * kk, b_hi, idesc1, 1u);
            }
            umma_commit(a_empty(sta));
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(acc_full(buf));
      __syncwarp();
    }
    if (dbg && lane == 0) {
      atomicAdd(a.debug + 0, (unsigned long long)w_acc);
      atomicAdd(a.debug + 1, (unsigned long long)w_a);
      atomicAdd(a.debug + 2, (unsigned long long)w_b);
      atomicAdd(a.debug + 3, (unsigned long long)(clock64() - t_begin));
      atomicAdd(a.debug + 10, 1ull);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue: group g owns rows [g*R/2, (g+1)*R/2)
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;           // TMEM lanes [32q, 32q+32) = pixels [32q, 32q+32) of the row segment
    const int px = quarter * 32 + lane;
    const uint32_t stg_g = stg_base + g * NSTG * Cfg::STG;  // this group's staging buffer(s)
    uint32_t nstore = 0;
    // 16-byte chunk c of pixel px lives at chunk (c ^ swz) of its row: the TMA swizzle of a CO*2-byte wide box
    const uint32_t swz = CO == 64 ? (uint32_t)(px & 7) : (CO == 32 ? (uint32_t)((px >> 1) & 3) : (uint32_t)((px >> 2) & 1));
    const uint32_t px_off = (uint32_t)px * (CO * 2);
    uint32_t tcount = 0;
    const bool dbg = a.debug != nullptr && warp == 2 && lane == 0;
    long long w_full = 0, t_body = 0, t0 = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1u;
      const RowTile tc = row_decode(a, tile, R);
      if (dbg) t0 = clock64();
      row_warp_wait(acc_full(buf), (tcount >> 1) & 1u, lane);
      if (dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * Cfg::ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      if (a.mode == kEpiBf16) {
#pragma unroll 1
        for (int rb = 0; rb < R / 2; rb += RB) {
          const int r0 = g * (R / 2) + rb;             // first row of this store box
          if (tc.y0 + r0 >= a.out_h) break;            // ragged last block (uniform over the group)
          if (res_tma) {
            // residual in the epilogue: the store thread has TMA-loaded the residual row(s) of this box into the
            // staging buffer (which also says the buffer is free); add in fp32, pack, overwrite in place
            const uint32_t stg_px = stg_g + px_off;
            row_warp_wait(res_full(g), nstore & 1u, lane);
            ++nstore;
#pragma unroll
            for (int i = 0; i < RB; ++i) {
              const uint32_t taddr = tacc + (uint32_t)((r0 + i) * CO);
#pragma unroll
              for (int ch = 0; ch < CO / Cfg::CH; ++ch) {
                uint32_t acc[Cfg::CH];
                if constexpr (Cfg::CH == 32) tmem_ld_32x32(taddr + ch * 32, reinterpret_cast<uint32_t(&)[32]>(acc));
                else tmem_ld_32x16(taddr + ch * 16, reinterpret_cast<uint32_t(&)[16]>(acc));
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < Cfg::CH; j += 8) {
                  const uint32_t c = (uint32_t)((ch * Cfg::CH + j) / 8);
                  const uint32_t addr = stg_px + (uint32_t)i * (kRowSeg * CO * 2) + ((c ^ swz) << 4);
                  uint32_t rv[4], o[4];
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv[0]), "=r"(rv[1]), "=r"(rv[2]), "=r"(rv[3]) : "r"(addr));
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 b2 = *reinterpret_cast<const float2*>(bias_s + ch * Cfg::CH + j + 2 * k);
                    const float2 r2 = unpack16(rv[k], a.fp16);
                    o[k] = pack16_sat(__uint_as_float(acc[j + 2 * k]) + b2.x + r2.x,
                                      __uint_as_float(acc[j + 2 * k + 1]) + b2.y + r2.y, a.fp16);
                    if (a.relu) o[k] = relu16x2(o[k], a.fp16);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                               : "memory");
                }
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(stg_full(g, 0));
            continue;
          }
          // 1. TMEM -> registers -> bias / ReLU / 16-bit pack for the whole box, BEFORE touching the staging buffer:
          //    this part overlaps with the previous box's TMA store still draining
          uint4 pk[RB][CO / 8];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const uint32_t taddr = tacc + (uint32_t)((r0 + i) * CO);
#pragma unroll
            for (int ch = 0; ch < CO / Cfg::CH; ++ch) {
              uint32_t acc[Cfg::CH];
              if constexpr (Cfg::CH == 32) tmem_ld_32x32(taddr + ch * 32, reinterpret_cast<uint32_t(&)[32]>(acc));
              else tmem_ld_32x16(taddr + ch * 16, reinterpret_cast<uint32_t(&)[16]>(acc));
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < Cfg::CH; j += 8) {
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 b2 = *reinterpret_cast<const float2*>(bias_s + ch * Cfg::CH + j + 2 * k);
                  o[k] = pack16_sat(__uint_as_float(acc[j + 2 * k]) + b2.x, __uint_as_float(acc[j + 2 * k + 1]) + b2.y,
                                    a.fp16);
                  if (a.relu) o[k] = relu16x2(o[k], a.fp16);
                }
                pk[i][(ch * Cfg::CH + j) / 8] = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          // 2. wait until the store thread has drained this staging buffer  3. swizzled writes + proxy fence
          // 4. hand the box to the store thread (warp 13 / 17): the TMA issue, its fence and the wait for the bulk
          //    read never sit on the epilogue warps' critical path, and no CTA-level barrier is involved
          const uint32_t sb = NSTG == 2 ? (nstore & 1u) : 0u;
          const uint32_t use = nstore / NSTG;
          const uint32_t stg_px = stg_g + sb * Cfg::STG + px_off;
          ++nstore;
          row_warp_wait(stg_empty(g, sb), (use & 1u) ^ 1u, lane);
#pragma unroll
          for (int i = 0; i < RB; ++i) {
#pragma unroll
            for (int c = 0; c < CO / 8; ++c) {
              const uint32_t dst = stg_px + (uint32_t)i * (kRowSeg * CO * 2) + (((uint32_t)c ^ swz) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[i][c].x), "r"(pk[i][c].y),
                           "r"(pk[i][c].z), "r"(pk[i][c].w)
                           : "memory");
            }
          }
          fence_proxy_async();  // this thread's staging writes -> visible to the async proxy (the TMA store)
          __syncwarp();
          if (lane == 0) mbar_arrive(stg_full(g, sb));
        }
      } else if (a.num_classes <= 4) {
        // head (2-4 classes): the group's R/2 rows are loaded from TMEM together and their softmaxes run interleaved,
        // so the TMEM-load, exp and division latencies of one row hide under the others
        constexpr int HR = R / 2;
        uint32_t acc[HR][4];
#pragma unroll
        for (int i = 0; i < HR; ++i) tmem_ld_32x4(tacc + (uint32_t)((g * HR + i) * CO), acc[i]);
        tmem_ld_wait();
        const int x = tc.x0 + px;
        const bool xok = x < a.out_w;
#pragma unroll
        for (int i = 0; i < HR; ++i) {
          const int y = tc.y0 + g * HR + i;
          if (xok && y < a.out_h) {
            if (a.num_classes == 2) softmax_store<2>(a, bias_s, acc[i], tc.n, y, x);
            else if (a.num_classes == 4) softmax_store<4>(a, bias_s, acc[i], tc.n, y, x);
            else if (a.num_classes == 3) softmax_store<3>(a, bias_s, acc[i], tc.n, y, x);
            else softmax_store<1>(a, bias_s, acc[i], tc.n, y, x);
          }
        }
      } else {
#pragma unroll 1
        for (int r = g * (R / 2); r < (g + 1) * (R / 2); ++r) {
          const int y = tc.y0 + r;
          if (y >= a.out_h) break;
          uint4 unused[EpiCfg<16>::RV];
          epilogue_pixel<16>(a, bias_s, 0, tacc + (uint32_t)(r * CO), tc.n, y, tc.x0 + px, tc.x0 + px < a.out_w, unused);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      if (dbg) t_body += clock64() - t0;
    }
    if (dbg) {
      atomicAdd(a.debug + 8, (unsigned long long)w_full);
      atomicAdd(a.debug + 9, (unsigned long long)t_body);
    }
  } else if ((warp & 3) == 1) {
    // ------------------------------------------------------------ store threads: warp 13 serves epilogue group 0, warp 17 group 1
    if (a.mode == kEpiBf16 && lane == 0) {
      const int g = warp == 13 ? 0 : 1;
      const uint32_t stg_g = stg_base + g * NSTG * Cfg::STG;
      uint32_t nstore = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const RowTile tc = row_decode(a, tile, R);
        for (int rb = 0; rb < R / 2; rb += RB) {
          const int r0 = g * (R / 2) + rb;
          if (tc.y0 + r0 >= a.out_h) break;
          if (res_tma) {
            // the buffer is free (the previous box's store has been read out below): fetch this box's residual rows into
            // it, then store the finished box from the same place
            mbar_arrive_expect_tx(res_full(g), Cfg::STG);
            tma_load_4d(stg_g, &a.rmap, res_full(g), 0, tc.x0, tc.y0 + r0, tc.n);
            if (a.res_tma > 1) {
              // this group's NEXT box (same block, or the first one of the CTA's next block) into L2 now: its load
              // sits between two stores of a one-buffer group and would otherwise pay the DRAM latency in full
              if (rb + RB < R / 2 && tc.y0 + r0 + RB < a.out_h) {
                tma_prefetch_4d(&a.rmap, 0, tc.x0, tc.y0 + r0 + RB, tc.n);
              } else if (tile + (int)gridDim.x < a.total_tiles) {
                const RowTile tn = row_decode(a, tile + (int)gridDim.x, R);
                tma_prefetch_4d(&a.rmap, 0, tn.x0, tn.y0 + g * (R / 2), tn.n);
              }
            }
            mbar_wait(stg_full(g, 0), nstore & 1u);
            tma_store_4d(&a.omap, stg_g, 0, tc.x0, tc.y0 + r0, tc.n);
            bulk_commit();
            bulk_wait_read_0();
            ++nstore;
            continue;
          }
          const uint32_t sb = NSTG == 2 ? (nstore & 1u) : 0u;
          const uint32_t use = nstore / NSTG;
          mbar_wait(stg_full(g, sb), use & 1u);
          tma_store_4d(&a.omap, stg_g + sb * Cfg::STG, 0, tc.x0, tc.y0 + r0, tc.n);  // rows beyond the image are clipped
          bulk_commit();
          if constexpr (NSTG == 2) {
            // keep one store in flight: the PREVIOUS box has been read out of its buffer -> hand that buffer back
            if (nstore > 0) {
              bulk_wait_read_1();
              mbar_arrive(stg_empty(g, sb ^ 1u));
            }
          } else {
            bulk_wait_read_0();
            mbar_arrive(stg_empty(g, 0));
          }
          ++nstore;
        }
      }
      bulk_wait_all();
    }
  } else if (!TMA_A) {
    // ------------------------------------------------------------ gather: (R+2) rows x 130 pixels x KC channels per stage
    // Thread t owns fixed (plane, pixel) columns of the stage and walks the rows: per copy one bounds test and one
    // pointer add.  Lanes run over the planes of consecutive pixels, so a warp reads contiguous global memory.
    // The 128 interior pixels take (128 * PLANES) / 256 columns per
    // thread; the two halo pixels (x0-1, x0+128) are 2 * PLANES * ROWS single copies spread over the first threads.
    constexpr int P = Cfg::PLANES;
    constexpr int ROWS = Cfg::ROWS;
    constexpr int COLS_PER_THREAD = (kRowSeg * P) / kRowGatherThreads;  // 2 (KC = 32) or 1 (KC = 16)
    constexpr int PX_STEP = kRowGatherThreads / P;
    constexpr int NHALO = 2 * P * ROWS;
    constexpr int DEPTH = Cfg::A_STAGES - 1;  // cp.async groups kept in flight
    const int gw = warp - 10;
    const int t = (gw - (gw + 1) / 4) * 32 + lane;            // rank among the gather warps * 32 + lane
    const int plane = t % P;
    const int pxm = t / P;                                    // interior pixel 1 + pxm (+ k * PX_STEP)
    const uint32_t dst_col = plane * Cfg::PLANE_STRIDE + (1 + pxm) * 16;
    const bool halo_thread = t < NHALO;
    const int h_plane = t % P, h_row = (t / P) % ROWS, h_side = t / (P * ROWS);  // side 0: x0-1, side 1: x0+128
    const uint32_t dst_halo = h_plane * Cfg::PLANE_STRIDE + (h_row * kRowHaloPx + h_side * (kRowHaloPx - 1)) * 16;
    const int nsrc = a.nseg + (has_res ? 1 : 0);
    uint32_t it = 0;
    const bool dbg = a.debug != nullptr && t == 0;
    long long g_empty = 0, g_issue = 0, g_land = 0, g_rest = 0, t0 = 0, t_begin = dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const RowTile tc = row_decode(a, tile, R);
      for (int s = 0; s < nsrc; ++s) {
        const bool is_res = s >= a.nseg;
        const int cin = is_res ? CO : a.seg[s].cin;
        const int up = is_res ? 0 : a.seg[s].up;
        const int src_w = a.out_w >> up;
        const size_t row_stride = (size_t)src_w * cin;
        const __nv_bfloat16* src = is_res ? a.residual : a.src_ptr[s];
        const __nv_bfloat16* img = src + (size_t)tc.n * (a.out_h >> up) * row_stride;
        for (int cc = 0; cc < cin / KC; ++cc, ++it) {
          const int st = it % Cfg::A_STAGES;
          if (dbg) t0 = clock64();
          row_warp_wait(a_empty(st), ((it / Cfg::A_STAGES) & 1) ^ 1u, lane);
          if (dbg) { const long long t1 = clock64(); g_empty += t1 - t0; t0 = t1; }
          const uint32_t stage = a_base + st * Cfg::STAGE_BYTES;
          // the residual (identity) segment only reads block rows 1..R and the interior pixels; an upsampled
          // segment gathers its R/2+2 SOURCE rows (row js = source row y0/2 - 1 + js) and stays upsampled along x
          const int j_lo = is_res ? 1 : 0, j_hi = is_res ? ROWS - 1 : (up ? Cfg::ROWS_UP : ROWS);
          const int src_y0 = up ? (tc.y0 >> 1) - 1 : tc.y0 - 1;
          const int src_h = a.out_h >> up;
#pragma unroll
          for (int k = 0; k < COLS_PER_THREAD; ++k) {
            const int gx = tc.x0 + pxm + k * PX_STEP;
            const bool xok = gx < a.out_w;  // only the last segment of a row can be ragged (width % 128 != 0)
            const __nv_bfloat16* col = img + (size_t)(gx >> up) * cin + cc * KC + plane * 8;
            const uint32_t dst = stage + dst_col + k * (PX_STEP * 16);
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
              if (j >= j_lo && j < j_hi) {
                const int sy = src_y0 + j;
                const bool ok = xok && (unsigned)sy < (unsigned)src_h;
                rcp_async_16(dst + j * kRowPitch, ok ? col + (size_t)sy * row_stride : src, ok ? 16u : 0u);
              }
            }
          }
          if (halo_thread && !is_res && h_row < j_hi) {
            const int sy = src_y0 + h_row, gx = tc.x0 - 1 + h_side * (kRowSeg + 1);
            const bool ok = ((unsigned)sy < (unsigned)src_h) && ((unsigned)gx < (unsigned)a.out_w);
            const __nv_bfloat16* gp = ok ? img + (size_t)sy * row_stride + (size_t)(gx >> up) * cin + cc * KC + h_plane * 8 : src;
            rcp_async_16(stage + dst_halo, gp, ok ? 16u : 0u);
          }
          rcp_async_commit();
          if (dbg) { const long long t1 = clock64(); g_issue += t1 - t0; t0 = t1; }
          if (DEPTH == 1 || it >= (uint32_t)(DEPTH - 1)) {
            rcp_async_wait<DEPTH - 1>();
            if (dbg) { const long long t1 = clock64(); g_land += t1 - t0; t0 = t1; }
            __syncwarp();
            if (lane == 0) {
              fence_proxy_async();  // one per warp, after the sync that orders the lanes' writes before it
              mbar_arrive(a_full((it - (DEPTH - 1)) % Cfg::A_STAGES));
            }
            if (dbg) g_rest += clock64() - t0;
          }
        }
      }
    }
    if (dbg) {
      atomicAdd(a.debug + 4, (unsigned long long)g_empty);
      atomicAdd(a.debug + 5, (unsigned long long)g_issue);
      atomicAdd(a.debug + 6, (unsigned long long)g_land);
      atomicAdd(a.debug + 7, (unsigned long long)(clock64() - t_begin));
      atomicAdd(a.debug + 12, (unsigned long long)g_rest);
    }
    rcp_async_wait<0>();
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();
      const uint32_t first = (DEPTH == 1 || it >= (uint32_t)(DEPTH - 1)) ? it - (DEPTH - 1) : 0u;
      for (uint32_t k = first; k < it; ++k) mbar_arrive(a_full(k % Cfg::A_STAGES));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (a.debug != nullptr && threadIdx.x == 0) atomicAdd(a.debug + 11, (unsigned long long)(clock64() - t_cta));
}

// ------------------------------------------------------------------------------------------------- host side
int conv_row_mode(const ConvArgs& a);
// The instantiations: <KC, KCB, CO, R, STAGES, RB, NSTG, STREAM>
//   Cout 64 (layer1, decoder block 2 conv2): 32-channel A stages x 2, 128-byte weight tiles, 4-row blocks, one
//                                            staging buffer per group (72-80 KB of resident weights leave no more)
//   Cout 64, streamed weights (decoder block 2 conv1, 264 KB of weights): 16-channel A stages x 3 that also carry the
//                                            chunk's three weight tiles (measured slower than the resident
//                                            configuration where the weights do fit: 208k vs 179k cycles on layer1)
//   Cout 32 (decoder block 3):               16-channel A stages x 3, 8-row blocks, one staging buffer per group
//                                            (conv1's four-slot + three-slot weight tiles take 84 KB)
//   Cout 16 (decoder block 4, head):         16-channel A stages x 4, 8-row blocks, 2 rows per TMA store
#define IU_ROW_CFG64 32, 64, 64, 4, 2, 1, 1, false
//   Cout 64 from ONE identity 64-channel source (layer1, decoder block 2 conv2), A ring filled by TMA: 64-channel
//                                            stages of three input rows x 2, 128B-swizzled
#define IU_ROW_CFG64T 64, 64, 64, 4, 2, 1, 1, false, true
#define IU_ROW_CFG64S 16, 16, 64, 4, 3, 1, 2, true
#define IU_ROW_CFG32 16, 16, 32, 8, 3, 1, 1, false
#define IU_ROW_CFG16 16, 16, 16, 8, 4, 2, 2, false

template <int KC, int KCB, int CO, int R, int STAGES, int RB, int NSTG, bool STREAM>
static bool row_fits(const ConvArgs& a) {
  using Cfg = RowCfg<KC, KCB, CO, R, STAGES, RB, NSTG, STREAM>;
  int bytes = 0;
  for (int s = 0; s < a.nseg; ++s) {
    if (a.seg[s].cin % KCB) return false;
    if (a.seg[s].up && (s != 0 || (a.out_h & 1))) return false;  // only the first segment may be upsampled
    bytes += 3 * (a.seg[s].cin / KCB) * (a.seg[s].up ? Cfg::U_TILE : Cfg::B_TILE);
  }
  if (a.residual != nullptr) bytes += (CO / KCB) * Cfg::I_TILE;
  if (STREAM) return a.residual == nullptr;  // nothing is resident; the identity segment is not wired for streaming
  return bytes <= Cfg::W_MAX;
}

int conv_row_kc(int cout_pad, int mode) { return (cout_pad == 64 && mode != 2) ? 64 : 16; }
int conv_row_store_rows(int cout_pad) { return cout_pad == 16 ? 2 : 1; }

bool conv_row_applicable(const ConvArgs& a) { return conv_row_mode(a) != 0; }

// 0: not for this kernel; 1: resident weights; 2: streamed weights (Cout 64 layers whose weights exceed shared memory)
int conv_row_mode(const ConvArgs& a) {
  if (a.nseg < 1 || a.nseg > 2 || a.up2x) return 0;
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].ksize != 3 || a.seg[s].stride != 1 || a.seg[s].pad != 1 || a.src_ptr[s] == nullptr) return 0;
  // any width >= 128: a ragged last segment is zero-filled by the gather and clipped by the TMA store; below 60 %
  // column utilisation the 16x16-block kernels are the better fit
  const int segs = (a.out_w + kRowSeg - 1) / kRowSeg;
  if (a.out_w < kRowSeg || a.out_w * 5 < 3 * kRowSeg * segs || a.out_h < 1) return 0;
  if (a.mode == kEpiBf16) {
    if (a.cout == 64) {
      // with resident weights the 16x16-block kernel is nearly as fast, so it takes over below 75 % utilisation
      // (width 160: two segments for 1.25); the streamed-weights layer stays here down to 60 %
      if (row_fits<IU_ROW_CFG64>(a)) return a.out_w * 4 >= 3 * kRowSeg * segs ? 1 : 0;
      return row_fits<IU_ROW_CFG64S>(a) ? 2 : 0;
    }
    if (a.cout == 32) return row_fits<IU_ROW_CFG32>(a) ? 1 : 0;
    if (a.cout == 16) return row_fits<IU_ROW_CFG16>(a) ? 1 : 0;
    return 0;
  }
  // softmax head: the logits occupy the first num_classes of 16 padded output channels
  return (a.residual == nullptr && a.num_classes <= 16 && row_fits<IU_ROW_CFG16>(a)) ? 1 : 0;
}

// The TMA-filled variant takes the Cout-64 layers whose only source is an identity 64-channel tensor (the engine
// encodes `rowmap` and sets `row_tma` for exactly those); a shortcut must then come through the epilogue (`res_tma`).
bool conv_row_tma_applicable(const ConvArgs& a) {
  return a.mode == kEpiBf16 && a.cout == 64 && a.nseg == 1 && a.seg[0].cin == 64 && !a.seg[0].up &&
         (a.residual == nullptr || a.res_tma != 0) && conv_row_mode(a) == 1;
}

template <int KC, int KCB, int CO, int R, int STAGES, int RB, int NSTG, bool STREAM, bool TMA_A = false>
static cudaError_t launch_row_one(const ConvArgs& args_in, cudaStream_t stream) {
  using Cfg = RowCfg<KC, KCB, CO, R, STAGES, RB, NSTG, STREAM, TMA_A>;
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_row_kernel<KC, KCB, CO, R, STAGES, RB, NSTG, STREAM, TMA_A>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured_dev = dev;
  }
  ConvArgs args = args_in;
  args.tiles_x = (args.out_w + kRowSeg - 1) / kRowSeg;
  args.tiles_y = (args.out_h + R - 1) / R;
  args.ntiles_n = 1;
  args.total_tiles = args.tiles_x * args.tiles_y * args.batch;
  const int grid = args.total_tiles < num_sms ? args.total_tiles : num_sms;
  conv_row_kernel<KC, KCB, CO, R, STAGES, RB, NSTG, STREAM, TMA_A><<<grid, kRowThreads, Cfg::SMEM_BYTES, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launch_conv_row(const ConvArgs& args, cudaStream_t stream) {
  const int mode = conv_row_mode(args);
  if (mode == 0) return cudaErrorInvalidValue;
  const int co = args.mode == kEpiBf16 ? args.cout : 16;
  if (co == 64 && mode == 2) return launch_row_one<IU_ROW_CFG64S>(args, stream);
  if (co == 64 && args.row_tma && conv_row_tma_applicable(args)) return launch_row_one<IU_ROW_CFG64T>(args, stream);
  if (co == 64) return launch_row_one<IU_ROW_CFG64>(args, stream);
  if (co == 32) return launch_row_one<IU_ROW_CFG32>(args, stream);
  return launch_row_one<IU_ROW_CFG16>(args, stream);
}

}  // namespace iu
