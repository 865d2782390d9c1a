// Per-tap TMA implicit-GEMM convolution on CTA PAIRS (tcgen05 cta_group::2) for the wide layers (Cout >= 128,
// 64-channel chunks): the Cout >= 128 layers of the encoder / decoder that conv_tc.cu runs one CTA at a time.
//
// Why: on those layers the single-CTA kernel sits on the SHARED-MEMORY bandwidth of an SM, not on the tensor pipe
// (profiles/r01_findings.md, findings 13-14): an M128 x N256 x K16 MMA reads 4 KB of A + 8 KB of B per 128 cycles
// (96 B/clk) while TMA writes another 48 KB per 512 cycles into the ring (94 B/clk) -- 190 B/clk against the SM's
// 128 B/clk, and the measured launch times equal that bound to within 10 %.  Two CTAs of a cluster (the two SMs of a
// TPC) running ONE tcgen05.mma.cta_group::2 stream (M = 256) each hold only HALF of every weight tile: the tensor
// cores exchange the B halves, so per SM an MMA reads 4 KB of A + 4 KB of B (64 B/clk) and TMA fills 32 KB per 512
// cycles (64 B/clk): 128 B/clk in total, and the weight traffic L2 -> SM is halved as well.
//   BN = 256, BM = 1: Cout >= 256 layers, one pixel tile per CTA;   BN = 128, BM = 2: Cout = 128 layers, two pixel
//   tiles per CTA on each weight half (both accumulators double-buffered: all 512 TMEM columns).
// Everything else is conv_tc.cu: one 4-D TMA box per filter tap with OOB zero fill = the conv's padding, element
// strides for stride 2, the fused 1x1/s2 downsample as a second K segment, persistent CTAs, double-buffered TMEM
// accumulators, conv_epilogue.cuh.  Barrier protocol (as in conv_halo.cu's pair kernel):
//   full / acc_empty : waited by the leader's MMA thread; arrivals come from both CTAs (TMA complete_tx addressed
//                      to the leader's barrier, remote mbarrier.arrive from the peer's epilogue warps);
//   empty / acc_full : signalled in both CTAs at once by the leader's multicast tcgen05.commit.
// Replaces the same cuDNN launches as conv_tc.cu (`smp.Unet.forward`, `/root/reference/interactive_unet/unet.py:67`).
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int BN, int BM>
struct Pair2Cfg {
  static constexpr int KC = 64;
  static constexpr int A_BYTES = kTileM * KC * 2;            // one 128-pixel tile of one tap: 16 KB
  static constexpr int B_HALF = (BN / 2) * KC * 2;           // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = BM * A_BYTES + B_HALF;  // 32 KB (BN 256) / 40 KB (BN 128, BM 2)
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int ACC_COLS = BM * BN;                   // per CTA and buffer
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (2 * STAGES + 5) * 8 + 1024;
  static_assert(TMEM_COLS == 512, "both shapes use the whole TMEM, double buffered");
  static_assert(SMEM_BYTES <= 227 * 1024, "pair kernel exceeds the shared memory of an SM");
};

constexpr int kPair2Threads = 320;  // warp 0 TMA producer, warp 1 MMA issue (leader), warps 2-5 / 6-9 epilogue groups

template <int BN, int BM>
__global__ void __launch_bounds__(kPair2Threads, 1) conv_tc2_kernel(const __grid_constant__ ConvArgs a) {
  using Cfg = Pair2Cfg<BN, BM>;
  constexpr int KC = Cfg::KC;
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
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // super tile s -> (Cout tile s % ntiles_n, 2*BM consecutive pixel tiles); this CTA's first pixel tile:
  const int mtiles = a.tiles_x * a.tiles_y * ((a.batch + a.nb - 1) / a.nb);
  const int n_super = ((mtiles + 2 * BM - 1) / (2 * BM)) * a.ntiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.amap[0]);
    if (a.nseg > 1) tma_prefetch_desc(&a.amap[1]);
    tma_prefetch_desc(BN == 256 ? &a.bmap : &a.bmap2);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), BM == 2 ? 16 : 8);  // the epilogue warps of BOTH CTAs that drain this buffer
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
    // ------------------------------------------------------------ TMA producer: this CTA's pixel tile(s) and weight half
    if (lane == 0) {
      uint32_t it = 0;
      for (int st_ = pair_id; st_ < n_super; st_ += n_pairs) {
        const int ntile = st_ % a.ntiles_n;
        const int m0 = (st_ / a.ntiles_n) * 2 * BM + (int)rank * BM;
        int kbase = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const ConvSegment sg = a.seg[s];
          for (int r = 0; r < sg.ksize; ++r) {
            for (int q = 0; q < sg.ksize; ++q) {
              for (int cc = 0; cc < sg.cin / KC; ++cc, ++it) {
                const int st = it % Cfg::STAGES;
                mbar_wait_peer(empty_bar(st), ((it / Cfg::STAGES) & 1) ^ 1u);
                const uint32_t leader_full = mapa_shared(full_bar(st), 0);
                if (rank == 0) mbar_arrive_expect_tx(full_bar(st), 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes
                const uint32_t sa = base + st * Cfg::STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < BM; ++j) {
                  const int m = m0 + j;   // a tile beyond the batch is zero-filled by TMA (image index out of bounds)
                  const int x0 = (m % a.tiles_x) * a.tw, y0 = ((m / a.tiles_x) % a.tiles_y) * a.th;
                  const int n0 = (m / (a.tiles_x * a.tiles_y)) * a.nb;
                  tma_load_4d_pair(sa + j * Cfg::A_BYTES, &a.amap[s], leader_full, cc * KC, x0 * sg.stride - sg.pad + q,
                                   y0 * sg.stride - sg.pad + r, n0);
                }
                tma_load_2d_pair(sa + BM * Cfg::A_BYTES, BN == 256 ? &a.bmap : &a.bmap2, leader_full,
                                 kbase + (r * sg.ksize + q) * sg.cin + cc * KC, ntile * BN + (int)rank * (BN / 2));
              }
            }
          }
          kbase += sg.ksize * sg.ksize * sg.cin;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(2 * kTileM, BN, a.fp16);
      uint32_t it = 0, tcount = 0;
      for (int st_ = pair_id; st_ < n_super; st_ += n_pairs, ++tcount) {
        const uint32_t buf = tcount & 1u;
        mbar_wait_cluster(acc_empty_bar(buf), ((tcount >> 1) & 1u) ^ 1u);  // both CTAs' epilogues drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * Cfg::ACC_COLS;
        uint32_t accumulate = 0;
        for (int s = 0; s < a.nseg; ++s) {
          const int n_it = a.seg[s].ksize * a.seg[s].ksize * (a.seg[s].cin / KC);
          for (int i = 0; i < n_it; ++i, ++it) {
            const int st = it % Cfg::STAGES;
            mbar_wait_cluster(full_bar(st), (it / Cfg::STAGES) & 1);
            operand_ready_fence();
            const uint32_t sa = base + st * Cfg::STAGE_BYTES;
            const uint64_t bdesc = umma_smem_desc<128>(sa + BM * Cfg::A_BYTES);
            const uint32_t b_lo = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
              for (int j = 0; j < BM; ++j) {
                const uint64_t adesc = umma_smem_desc<128>(sa + j * Cfg::A_BYTES);
                umma_f16_pair_lohi(tmem_d + j * BN, (uint32_t)adesc + 2u * k, (uint32_t)(adesc >> 32), b_lo + 2u * k, b_hi,
                                   idesc, accumulate);
              }
              accumulate = 1;
            }
            umma_commit_pair(empty_bar(st));  // the stage is reusable in both CTAs once these MMAs have read it
          }
        }
        umma_commit_pair(acc_full_bar(buf));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue groups (warps 2-5 and 6-9) of this CTA's rows
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t acc_empty_leader = mapa_shared(acc_empty_bar(0), 0);
    uint32_t tcount = 0;
    for (int st_ = pair_id; st_ < n_super; st_ += n_pairs, ++tcount) {
      const uint32_t buf = tcount & 1u;
      // BM == 1: group g drains buffer g; BM == 2: group g drains this CTA's pixel tile g of every super tile
      if (BM == 1 && (int)buf != group) continue;
      const int ntile = st_ % a.ntiles_n;
      const int m = (st_ / a.ntiles_n) * 2 * BM + (int)rank * BM + (BM == 2 ? group : 0);
      const int x0 = (m % a.tiles_x) * a.tw, y0 = ((m / a.tiles_x) % a.tiles_y) * a.th;
      const int n0 = (m / (a.tiles_x * a.tiles_y)) * a.nb;
      const int per_img = a.th * a.tw;
      const int n = n0 + row / per_img;
      const int y = y0 + (row % per_img) / a.tw;
      const int x = x0 + row % a.tw;
      const bool valid = (n < a.batch) && (y < a.out_h) && (x < a.out_w);
      uint4 res[EpiCfg<BN>::RV];
      residual_prefetch<BN>(a, ntile, n, y, x, valid, res);
      if (lane == 0) mbar_wait_peer(acc_full_bar(buf), (tcount >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * Cfg::ACC_COLS + (BM == 2 ? group * BN : 0) + ((uint32_t)(quarter * 32) << 16);
      epilogue_pixel<BN>(a, a.bias, ntile, taddr, n, y, x, valid, res);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8u * buf);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still signal this CTA's barriers / read its weight half until here
  if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int BM>
static cudaError_t launch_pair2(const ConvArgs& args_in, cudaStream_t stream) {
  using Cfg = Pair2Cfg<BN, BM>;
  static int configured_dev = -1;
  static int num_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured_dev = dev;
  }
  ConvArgs args = args_in;
  const int mtiles = args.tiles_x * args.tiles_y * ((args.batch + args.nb - 1) / args.nb);
  args.ntiles_n = args.cout / BN;
  const int n_super = ((mtiles + 2 * BM - 1) / (2 * BM)) * args.ntiles_n;
  args.total_tiles = n_super;
  const int max_pairs = num_sms / 2;
  const int pairs = n_super < max_pairs ? n_super : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kPair2Threads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BN, BM>, args);
}

// 16-bit outputs only, 64-channel chunks, Cout a multiple of the tile width, sources read at their own resolution
// (an upsampled segment needs the halo / row kernels' gather).
bool conv_tc2_applicable(const ConvArgs& a, int bn) {
  if (a.mode != kEpiBf16 || a.nseg < 1 || a.nseg > 2 || (bn != 128 && bn != 256) || a.cout % bn) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].up || a.seg[s].cin % 64 || (a.seg[s].ksize != 1 && a.seg[s].ksize != 3)) return false;
  return true;
}

cudaError_t launch_conv_tc2(const ConvArgs& args, int bn, cudaStream_t stream) {
  if (!conv_tc2_applicable(args, bn)) return cudaErrorInvalidValue;
  return bn == 256 ? launch_pair2<256, 1>(args, stream) : launch_pair2<128, 2>(args, stream);
}

}  // namespace iu
