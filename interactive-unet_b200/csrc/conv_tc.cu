// Implicit-GEMM convolution on the sm_100a tensor cores (persistent, warp-specialised).
//
//   D[128 pixels, BN channels] = sum over (segment, tap r,q, channel chunk)  A_tap[128, KC] * W_tap[BN, KC]^T
//
// * activations are NHWC fp16/bf16; the A tile of one filter tap is ONE 4-D TMA box (KC channels x tw cols x
//   th rows x nb images) fetched at the tap's offset: out-of-bounds coordinates are zero-filled by the
//   TMA unit, which is the convolution's zero padding; stride-2 convs use the map's element strides.
// * weights are pre-packed [Cout_pad][K] (K ordered exactly like the loop above) and fetched as 2-D TMA
//   boxes; both operands land in the 128B/64B/32B-swizzled K-major layout tcgen05.mma reads.
// * persistent CTAs (2 per SM) walk the tile list; three pipelines overlap inside a CTA:
//     TMA producer warp  --(smem stage ring, full/empty mbarriers)-->  MMA thread
//     MMA thread         --(2 TMEM accumulators, full/empty mbarriers)-->  2 epilogue warp groups
//   so the loads of tile i+1 and the math of tile i+1 run under the epilogue of tile i.
//   Small-K layers (KC <= 32) put the 3 taps of a filter row into one stage to cut barrier traffic.
// * the epilogue reads the accumulator with tcgen05.ld and fuses folded-BN bias, residual add, ReLU,
//   16-bit pack, optional nearest-2x upsampled store, or (head conv) softmax + probability store.
//
// Replaces the cuDNN conv2d / batch_norm / relu / add / cat / upsample_nearest2d / softmax launches
// issued by `smp.Unet.forward` under `/root/reference/interactive_unet/unet.py:67`.
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

constexpr int kSmemBudget = 100 * 1024;   // per CTA, so that two CTAs fit in the 227 KB of an SM
constexpr int kSmemBudgetWide = 200 * 1024;  // BN = 256: one CTA per SM (its two accumulators fill TMEM)

// BM: M tiles (of 128 pixels) per CTA tile.  BM = 2 fetches ONE weight box for two pixel tiles: with BN = 128 that is
// 48 KB of operands per eight MMA-equivalents (instead of 64), with BN = 256 it is 64 KB per sixteen (instead of 96).
template <int KC, int BN, int BM = 1>
struct ConvCfg {
  static constexpr int SW = KC * 2;  // bytes per operand row == TMA/UMMA swizzle span
  static constexpr int A_BYTES = kTileM * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int B_ALLOC = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int TAP_BYTES = BM * A_BYTES + B_ALLOC;
  static constexpr int TPS = KC <= 32 ? 3 : 1;  // taps per pipeline stage (3x3 layers only)
  static constexpr int STAGE_BYTES = TPS * TAP_BYTES;
  // BN = 256 (Cout >= 256 layers): one A box serves a 256-wide weight tile, i.e. 48 KB of operands per eight
  // MMA-equivalents instead of 64 KB with two BN = 128 tiles -- the kernel is L2-bandwidth bound (~17 TB/s measured)
  static constexpr int CTAS_PER_SM = (BN == 256 || BM == 2) ? 1 : 2;
  static constexpr int STAGES_RAW = (CTAS_PER_SM == 1 ? kSmemBudgetWide : kSmemBudget) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int TILE_COLS = BN < 32 ? 32 : BN;   // TMEM columns of one M tile's accumulator
  static constexpr int ACC_COLS = BM * TILE_COLS;       // ... of one CTA tile
  static constexpr int NBUF = 2 * ACC_COLS <= 512 ? 2 : 1;  // double-buffered unless one CTA tile fills TMEM (BN 256, BM 2)
  static constexpr int TMEM_COLS = NBUF * ACC_COLS;
  static_assert(BM == 1 || (BM == 2 && KC == 64), "two-M-tile CTA tiles are instantiated for 64-channel chunks only");
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + 1) * 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
  static constexpr int CHUNK = BN >= 32 ? 32 : 16;                            // accumulator columns per tcgen05.ld
};

// CTA tile `tile` -> Cout tile and its first M tile; M tile m -> pixel coordinates.  With BM == 2 a CTA tile is the
// pair of M tiles (2p, 2p + 1); an odd tail pairs with a tile beyond the batch, whose TMA boxes are zero-filled and
// whose epilogue rows are masked (n >= batch).
__device__ __forceinline__ TileCoord mtile_coord(const ConvArgs& a, int m, int ntile) {
  TileCoord t;
  t.ntile = ntile;
  t.x0 = (m % a.tiles_x) * a.tw;
  t.y0 = ((m / a.tiles_x) % a.tiles_y) * a.th;
  t.n0 = (m / (a.tiles_x * a.tiles_y)) * a.nb;
  return t;
}

// warp 0: TMA producer, warp 1: TMEM alloc + MMA issue, warps 2-5 / 6-9: epilogue groups 0 / 1
constexpr int kThreads = 320;

// CL = 2: the kernel runs as clusters of two CTAs that work on neighbouring pixel tiles of the SAME Cout tile; each CTA
// fetches half of every weight box and multicasts it into both (the kernel is L2-bandwidth bound and the weight box is
// a third to two thirds of its operand bytes).  A stage may be refilled once BOTH CTAs' MMAs have read it, so the
// MMA issuer's commit arrives on the stage's empty barrier in both CTAs (count 2).  Launched with ntiles_n == 1 only.
template <int KC, int BN, int BM, int CL>
__global__ void __launch_bounds__(kThreads, ConvCfg<KC, BN, BM>::CTAS_PER_SM) conv_tc_kernel(const __grid_constant__ ConvArgs a) {
  using Cfg = ConvCfg<KC, BN, BM>;
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
  // both CTAs of a cluster must run the same number of tiles: an odd total is rounded up with a tile beyond the batch
  // (zero-filled operands, masked epilogue)
  const int tiles_end = CL == 2 ? ((a.total_tiles + 1) & ~1) : a.total_tiles;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto acc_full_bar = [&](int b) { return bar_base + 16u * Cfg::STAGES + 8u * b; };
  auto acc_empty_bar = [&](int b) { return bar_base + 16u * Cfg::STAGES + 16u + 8u * b; };
  const uint32_t tmem_slot = bar_base + 16u * Cfg::STAGES + 32u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // IU_CONV_DEBUG=1: per-role cycle counters, same slots as conv_halo.cu / conv_row.cu (0 MMA wait for a drained
  // accumulator, 1 MMA wait for operands, 3 MMA thread total, 4 producer wait for a free stage, 7 producer total,
  // 8 epilogue wait for a full accumulator, 9 epilogue body, 10 CTAs, 11 CTA life)
  // (only in the one-CTA-per-SM shapes: the two-CTA shapes have no registers to spare for the counters)
  const bool dbg_on = Cfg::CTAS_PER_SM == 1 && a.debug != nullptr;
  const long long t_cta = (dbg_on && threadIdx.x == 0) ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.amap[0]);
    if (a.nseg > 1) tma_prefetch_desc(&a.amap[1]);
    tma_prefetch_desc(BN == 256 ? &a.bmap256 : &a.bmap);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), BM == 2 ? 8 : 4);  // one arrival per epilogue warp that drains this buffer
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      long long p_empty = 0, p_t0 = 0;
      const long long p_begin = dbg_on ? clock64() : 0;
      for (int tile = blockIdx.x; tile < tiles_end; tile += gridDim.x) {
        const int ntile = tile % a.ntiles_n, m0 = (tile / a.ntiles_n) * BM;
        const TileCoord tc = mtile_coord(a, m0, ntile);
        const TileCoord tc1 = mtile_coord(a, m0 + BM - 1, ntile);  // second M tile (BM == 2)
        int kbase = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const ConvSegment sg = a.seg[s];
          const int chunks = sg.cin / KC;
          const int tps = sg.ksize == 3 ? Cfg::TPS : 1;
          for (int r = 0; r < sg.ksize; ++r) {
            for (int q0 = 0; q0 < sg.ksize; q0 += tps) {
              for (int cc = 0; cc < chunks; ++cc, ++it) {
                const int st = it % Cfg::STAGES;
                const uint32_t ph = (it / Cfg::STAGES) & 1;
                if (dbg_on) p_t0 = clock64();
                if constexpr (CL == 2) mbar_wait_peer(empty_bar(st), ph ^ 1u);
                else mbar_wait(empty_bar(st), ph ^ 1u);
                if (dbg_on) p_empty += clock64() - p_t0;
                mbar_arrive_expect_tx(full_bar(st), tps * (BM * Cfg::A_BYTES + Cfg::B_BYTES));
                for (int j = 0; j < tps; ++j) {
                  const int q = q0 + j;
                  const uint32_t sa = base + st * Cfg::STAGE_BYTES + j * Cfg::TAP_BYTES;
                  tma_load_4d(sa, &a.amap[s], full_bar(st), cc * KC, tc.x0 * sg.stride - sg.pad + q,
                              tc.y0 * sg.stride - sg.pad + r, tc.n0);
                  if constexpr (BM == 2)
                    tma_load_4d(sa + Cfg::A_BYTES, &a.amap[s], full_bar(st), cc * KC, tc1.x0 * sg.stride - sg.pad + q,
                                tc1.y0 * sg.stride - sg.pad + r, tc1.n0);
                  if constexpr (CL == 2)
                    tma_load_2d_multicast(sa + BM * Cfg::A_BYTES + crank * (Cfg::B_BYTES / 2),
                                          BN == 256 ? &a.bmap : &a.bmap2, full_bar(st),
                                          kbase + (r * sg.ksize + q) * sg.cin + cc * KC,
                                          tc.ntile * BN + (int)crank * (BN / 2), (uint16_t)3);
                  else
                    tma_load_2d(sa + BM * Cfg::A_BYTES, BN == 256 ? &a.bmap256 : &a.bmap, full_bar(st),
                                kbase + (r * sg.ksize + q) * sg.cin + cc * KC, tc.ntile * BN);
                }
              }
            }
          }
          kbase += sg.ksize * sg.ksize * sg.cin;
        }
      }
      if (dbg_on) {
        atomicAdd(a.debug + 4, (unsigned long long)p_empty);
        atomicAdd(a.debug + 7, (unsigned long long)(clock64() - p_begin));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kTileM, BN, a.fp16);
      int iters_per_tile = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int tps = a.seg[s].ksize == 3 ? Cfg::TPS : 1;
        iters_per_tile += a.seg[s].ksize * (a.seg[s].ksize / tps) * (a.seg[s].cin / KC);
      }
      uint32_t it = 0, tcount = 0;
      long long w_acc = 0, w_op = 0, m_t0 = 0;
      const long long m_begin = dbg_on ? clock64() : 0;
      for (int tile = blockIdx.x; tile < tiles_end; tile += gridDim.x, ++tcount) {
        const uint32_t buf = Cfg::NBUF == 2 ? (tcount & 1u) : 0u;
        const uint32_t use = Cfg::NBUF == 2 ? (tcount >> 1) : tcount;  // how often this buffer has been used before
        if (dbg_on) m_t0 = clock64();
        mbar_wait(acc_empty_bar(buf), (use & 1u) ^ 1u);  // epilogue has drained this accumulator
        if (dbg_on) w_acc += clock64() - m_t0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * Cfg::ACC_COLS;
        uint32_t first = 1;
        int local = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const ConvSegment sg = a.seg[s];
          const int tps = sg.ksize == 3 ? Cfg::TPS : 1;
          const int n_it = sg.ksize * (sg.ksize / tps) * (sg.cin / KC);
          for (int i = 0; i < n_it; ++i, ++it, ++local) {
            const int st = it % Cfg::STAGES;
            const uint32_t ph = (it / Cfg::STAGES) & 1;
            if (dbg_on) m_t0 = clock64();
            mbar_wait(full_bar(st), ph);
            if (dbg_on) w_op += clock64() - m_t0;
            operand_ready_fence();
            for (int j = 0; j < tps; ++j) {
              const uint32_t sa = base + st * Cfg::STAGE_BYTES + j * Cfg::TAP_BYTES;
              const uint64_t adesc = umma_smem_desc<Cfg::SW>(sa);
              const uint64_t bdesc = umma_smem_desc<Cfg::SW>(sa + BM * Cfg::A_BYTES);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
                // advancing K by 16 elements = 32 bytes inside the swizzle span = +2 in the (addr >> 4) field
                umma_f16(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, first ? 0u : 1u);
                if constexpr (BM == 2) {
                  const uint64_t adesc1 = umma_smem_desc<Cfg::SW>(sa + Cfg::A_BYTES);
                  umma_f16(tmem_d + Cfg::TILE_COLS, adesc1 + 2u * k, bdesc + 2u * k, idesc, first ? 0u : 1u);
                }
                first = 0;
              }
            }
            if constexpr (CL == 2) umma_commit_multicast(empty_bar(st), (uint16_t)3);  // ... in both CTAs
            else umma_commit(empty_bar(st));  // stage reusable once these MMAs have read it
          }
        }
        (void)iters_per_tile;
        (void)local;
        umma_commit(acc_full_bar(buf));  // accumulator complete
      }
      if (dbg_on) {
        atomicAdd(a.debug + 0, (unsigned long long)w_acc);
        atomicAdd(a.debug + 1, (unsigned long long)w_op);
        atomicAdd(a.debug + 3, (unsigned long long)(clock64() - m_begin));
        atomicAdd(a.debug + 10, 1ull);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue groups (warps 2-5 and 6-9)
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;  // TMEM lanes [32q, 32q+32) are the only ones this warp may read
    const int row = quarter * 32 + lane;
    uint32_t tcount = 0;
    const bool edbg = dbg_on && warp == 2 && lane == 0;
    long long e_wait = 0, e_body = 0, e_t0 = 0;
    for (int tile = blockIdx.x; tile < tiles_end; tile += gridDim.x, ++tcount) {
      const uint32_t buf = Cfg::NBUF == 2 ? (tcount & 1u) : 0u;
      const uint32_t use = Cfg::NBUF == 2 ? (tcount >> 1) : tcount;
      // BM == 1: group g drains the tiles of buffer g; BM == 2: group g drains M tile g of every CTA tile
      if (BM == 1 && (int)buf != group) continue;
      const TileCoord tc = mtile_coord(a, (tile / a.ntiles_n) * BM + (BM == 2 ? group : 0), tile % a.ntiles_n);
      const int per_img = a.th * a.tw;
      const int n = tc.n0 + row / per_img;
      const int y = tc.y0 + (row % per_img) / a.tw;
      const int x = tc.x0 + row % a.tw;
      const bool valid = (n < a.batch) && (y < a.out_h) && (x < a.out_w);
      uint4 res[EpiCfg<BN>::RV];
      residual_prefetch<BN>(a, tc.ntile, n, y, x, valid, res);
      if (edbg) e_t0 = clock64();
      mbar_wait(acc_full_bar(buf), use & 1u);
      if (edbg) { const long long t1 = clock64(); e_wait += t1 - e_t0; e_t0 = t1; }
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * Cfg::ACC_COLS + (BM == 2 ? group * Cfg::TILE_COLS : 0) +
                             ((uint32_t)(quarter * 32) << 16);
      epilogue_pixel<BN>(a, a.bias, tc.ntile, taddr, n, y, x, valid, res);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty_bar(buf));
      if (edbg) e_body += clock64() - e_t0;
    }
    if (edbg) {
      atomicAdd(a.debug + 8, (unsigned long long)e_wait);
      atomicAdd(a.debug + 9, (unsigned long long)e_body);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();  // no CTA leaves while its peer can still multicast into it
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (dbg_on && threadIdx.x == 0) atomicAdd(a.debug + 11, (unsigned long long)(clock64() - t_cta));
}

template <int KC, int BN, int BM = 1, int CL = 1>
static cudaError_t launch_one(const ConvArgs& args_in, cudaStream_t stream) {
  using Cfg = ConvCfg<KC, BN, BM>;
  static int configured_dev = -1;  // per kernel instantiation; the attribute is per device
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<KC, BN, BM, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured_dev = dev;
  }
  for (int s = 0; s < args_in.nseg; ++s)
    if (Cfg::TPS != 1 && args_in.seg[s].ksize != 3 && args_in.seg[s].ksize != 1) return cudaErrorInvalidValue;
  ConvArgs args = args_in;
  const int mtiles = args.tiles_x * args.tiles_y * ((args.batch + args.nb - 1) / args.nb);
  const int cout_pad = (args.mode == kEpiBf16) ? args.cout : BN;
  args.ntiles_n = cout_pad / BN;
  args.total_tiles = ((mtiles + BM - 1) / BM) * args.ntiles_n;
  const int slots = Cfg::CTAS_PER_SM * num_sms;
  int grid = args.total_tiles < slots ? args.total_tiles : slots;
  if constexpr (CL == 2) {
    if (args.ntiles_n != 1) return cudaErrorInvalidValue;  // the two CTAs of a cluster must share their weights
    grid = (grid + 1) & ~1;
    if (grid > slots) grid = slots & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv_tc_kernel<KC, BN, BM, CL>, args);
  }
  conv_tc_kernel<KC, BN, BM, CL><<<grid, kThreads, Cfg::SMEM_BYTES, stream>>>(args);
  return cudaGetLastError();
}

#define IU_CONV_DISPATCH(KC_, BN_) \
  if (kc == KC_ && bn == BN_) return launch_one<KC_, BN_>(args, stream);

cudaError_t launch_conv_tc(const ConvArgs& args, int kc, int bn, cudaStream_t stream, int bm, int cluster) {
  if (cluster == 2 && kc == 64 && bn == 256 && bm == 1) return launch_one<64, 256, 1, 2>(args, stream);
  if (cluster == 2 && kc == 64 && bn == 128 && bm == 2) return launch_one<64, 128, 2, 2>(args, stream);
  if (bm == 2 && kc == 64 && bn == 256) return launch_one<64, 256, 2>(args, stream);
  if (bm == 2 && kc == 64 && bn == 128) return launch_one<64, 128, 2>(args, stream);
  IU_CONV_DISPATCH(64, 256)
  IU_CONV_DISPATCH(64, 128)
  IU_CONV_DISPATCH(64, 64)
  IU_CONV_DISPATCH(64, 32)
  IU_CONV_DISPATCH(32, 32)
  IU_CONV_DISPATCH(32, 16)
  IU_CONV_DISPATCH(16, 16)
  return cudaErrorInvalidValue;
}

int conv_tc_smem_bytes(int kc, int bn) {
#define IU_CONV_SMEM(KC_, BN_) \
  if (kc == KC_ && bn == BN_) return ConvCfg<KC_, BN_>::SMEM_BYTES;
  IU_CONV_SMEM(64, 256)
  IU_CONV_SMEM(64, 128)
  IU_CONV_SMEM(64, 64)
  IU_CONV_SMEM(64, 32)
  IU_CONV_SMEM(32, 32)
  IU_CONV_SMEM(32, 16)
  IU_CONV_SMEM(16, 16)
  return -1;
}

}  // namespace iu
