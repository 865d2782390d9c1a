// Pipe-rate probes behind the kernel design decisions in DESIGN.md section 4 (development aid, B200):
//   1. tcgen05.mma cycles per instruction vs N (including the "taps folded into N" shapes 48 / 96 / 192) and vs the
//      number of MMAs per tcgen05.commit, with the issuing thread never waiting (deep window);
//   2. the same with two CTAs per SM (the <= 32-channel layers run two CTAs per SM);
//   3. tcgen05.ld throughput with 4 / 8 / 16 epilogue warps (bytes per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I interactive-unet_b200/csrc tools/pipe_probe.cu -o /tmp/pipe_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace iu;

__device__ __forceinline__ uint64_t desc_planar(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46;
}

// One issuing thread per CTA: `total` MMAs (M=128, K=16, N=n) with a tcgen05.commit after every `per`; the thread
// only waits for the LAST commit.  A = planar halo layout (as conv_halo), B = 128B-swizzled K-major.
__global__ void __launch_bounds__(128) mma_rate(int n, int per, int total, int tmem_cols, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_f16(128, n, 1);
    const uint64_t bdesc = umma_smem_desc<128>(base + 48 * 1024);
    const uint64_t a0 = desc_planar(base, 5200, 288);
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
    long long t0 = clock64();
    if (elect_one()) {
      int since = 0;
      for (int i = 0; i < total; ++i) {
        const uint32_t off = (uint32_t)((i % 3) * 18 + (i % 4) * 2 * 325);
        umma_f16_lohi(tmem, a_lo + off, a_hi, b_lo + 2u * (i & 3), b_hi, idesc, 1u);
        if (++since == per) {
          umma_commit(bar0 + 8);  // nobody waits on this one: phases just flip
          since = 0;
        }
      }
      umma_commit(bar0);
    }
    __syncwarp();
    if (threadIdx.x == 0) mbar_wait(bar0, 0);
    __syncwarp();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, tmem_cols);
}

// `warps` warps, each reading `reps` x (32 lanes x 32 columns) from its TMEM lane quarter.
__global__ void __launch_bounds__(512) tmem_ld_rate(int reps, unsigned long long* out) {
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32 % 512);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + (uint32_t)((r * 32) & 255), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

static double avg(unsigned long long* d_out, int n) {
  std::vector<unsigned long long> h(n);
  cudaMemcpy(h.data(), d_out, n * 8, cudaMemcpyDeviceToHost);
  double a = 0;
  for (auto v : h) a += (double)v;
  return a / n;
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 2048 * 8);
  cudaMemset(d_out, 0, 2048 * 8);
  const int total = 2880;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    const int smem = ctas == 1 ? 200 * 1024 : 100 * 1024;
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
      for (int per : {4, 8, 16, 36, 72, 2880}) {
        mma_rate<<<148 * ctas, 128, smem>>>(n, per, total, 256, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mma_rate n=%d per=%d: %s\n", n, per, cudaGetErrorString(e));
          return 1;
        }
        const double cyc = avg(d_out, 148 * ctas) / total;
        printf("mma  ctas/SM=%d N=%3d per-commit=%4d : %6.1f cycles/MMA per CTA  (%6.1f per SM; math floor %3d)\n", ctas, n,
               per, cyc, cyc / ctas, n / 2);
      }
    }
  }
  for (int warps : {4, 8, 16}) {
    const int reps = 4096;
    tmem_ld_rate<<<148, warps * 32>>>(reps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("tmem_ld_rate warps=%d: %s\n", warps, cudaGetErrorString(e));
      return 1;
    }
    const double cyc = avg(d_out, 148);
    printf("tmem ld 32x32b.x32: %2d warps : %7.1f cycles per ld per warp, %7.1f B/clk/SM\n", warps, cyc / reps,
           (double)warps * reps * 4096.0 / cyc);
  }
  return 0;
}
