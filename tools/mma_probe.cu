// tcgen05.mma issue-rate probe (development aid): cycles per MMA for the operand layouts the conv kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I interactive-unet_b200/csrc tools/mma_probe.cu -o /tmp/mma_probe
// One CTA per SM (or two), one thread issues `batches` x `per` MMAs (M=128, K=16) on garbage shared memory,
// committing each batch to an mbarrier and staying two batches ahead, like the conv kernels' stage rings.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace iu;

__device__ __forceinline__ uint64_t desc_planar(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46;
}

// mode 0: A,B 128B-swizzled K-major (KC=64 tiles)   mode 1: A planar no-swizzle, 128B-aligned start, B SW128
// mode 2: A planar, start shifted by 16 B * tap     mode 3: as 2 but halo-row SBO (288 B) like conv_halo
// mode 4: A,B 32B-swizzle (KC=16)                   mode 5: A from TMEM (.ts), B SW128
template <int mode>
__global__ void __launch_bounds__(128) probe(int n, int per, int batches, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const uint32_t idesc = umma_idesc_f16(128, n, 1);
    const uint32_t a_addr = base, b_addr = base + 96 * 1024;
    const uint64_t bdesc = mode == 4 ? umma_smem_desc<32>(b_addr) : umma_smem_desc<128>(b_addr);
    // per-MMA descriptors are base.lo + compile-time constants (fully unrolled), as in the conv kernels
    const uint64_t a0 = mode == 0 ? umma_smem_desc<128>(a_addr)
                      : mode == 4 ? umma_smem_desc<32>(a_addr)
                      : mode == 3 ? desc_planar(a_addr, 5200, 288) : desc_planar(a_addr, 5248, 128);
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
    long long t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      if (b >= 2) {
        if (lane == 0) mbar_wait(bar0 + 8 * (b & 1), ((b - 2) >> 1) & 1);
        __syncwarp();
      }
      if (elect_one()) {
      for (int rep = 0; rep < per / 18; ++rep) {
#pragma unroll
        for (int i = 0; i < 18; ++i) {
          const int tap = i % 9, kk = i / 9;
          uint32_t off;
          if (mode == 0) off = (tap & 3) * 1024 + 2 * kk;
          else if (mode == 1) off = kk * 2 * 328;
          else if (mode == 2) off = kk * 2 * 328 + tap;
          else if (mode == 3) off = kk * 2 * 325 + (tap / 3) * 18 + tap % 3;
          else off = tap * 256;
          if (mode == 5) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem + 256),
                "r"(tmem + kk * 8), "r"(b_lo + 2u * kk), "r"(b_hi), "r"(idesc), "r"(1u)
                : "memory");
          } else {
            umma_f16_lohi(tmem + (i & 1) * 256, a_lo + off, a_hi, b_lo + (mode == 4 ? 0u : 2u * kk), b_hi, idesc, 1u);
          }
        }
      }
      umma_commit(bar0 + 8 * (b & 1));
      }
      __syncwarp();
    }
    for (int b = batches - 2; b < batches; ++b)
      if (b >= 0 && lane == 0) mbar_wait(bar0 + 8 * (b & 1), (b >> 1) & 1);
    __syncwarp();
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 1024 * 8);
  const int smem = 200 * 1024;
  const char* names[] = {"SW128 A+B", "planar A aligned", "planar A +16B*tap", "planar A halo rows", "SW32 A+B", "A in TMEM"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int n : {16, 32, 64, 128, 256}) {
      for (int per : {18, 72}) {
        const int batches = 2000 / per * 4;
#define RUN(M_) case M_: cudaFuncSetAttribute(probe<M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); probe<M_><<<148, 128, smem>>>(n, per, batches, d_out); break;
        switch (mode) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e));
          return 1;
        }
        std::vector<unsigned long long> h(148);
        cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= 148.0 * batches * per;
        printf("%-20s N=%3d per-commit=%2d : %6.1f cycles/MMA (math floor %3d, smem model %3d)\n", names[mode], n, per,
               avg, n / 2, mode == 5 ? n / 4 : (128 + n) / 4);
      }
    }
  }
  return 0;
}
