// What slows tcgen05.mma down inside the conv kernels?  (development aid, B200)
// One thread issues M=128 MMAs (planar A, swizzled B, as conv_row.cu) while 8 other warps of the same CTA generate one
// kind of background traffic.  Prints cycles per MMA for each background kind.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I interactive-unet_b200/csrc tools/contention_probe.cu -o /tmp/cp
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace iu;

__device__ __forceinline__ uint64_t desc_planar(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46;
}
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(dst), "l"(src) : "memory");
}

// bg: 0 none, 1 cp.async 16 B (planar scatter like the gather), 2 tcgen05.ld, 3 st.shared.v4, 4 ld.global only (no smem),
//     5 st.global.v4 strided 128 B (the old epilogue pattern), 6 cp.async.bulk (TMA engine) 4 KB chunks
__global__ void __launch_bounds__(320) probe(int n, int total, int bg, const uint4* gsrc, uint4* gdst,
                                             unsigned long long* out, volatile int* stop_flag) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
    done = 0;
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(128, n, 1);
    const uint64_t bdesc = umma_smem_desc<32>(base + 100 * 1024);
    const uint64_t a0 = desc_planar(base, 20816, 128);
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
    long long t0 = clock64();
    if (elect_one()) {
      // lean issue stream like the unrolled kernels: compile-time descriptor offsets, ~2 instructions per MMA
      for (int i = 0; i < total; i += 30) {
#pragma unroll
        for (int k = 0; k < 30; ++k)
          umma_f16_lohi(tmem + (uint32_t)((k % 4) * 32), a_lo + (uint32_t)((k % 10) * 130 + (k / 10)), a_hi, b_lo, b_hi, idesc, 1u);
        umma_commit(bar0 + 8);
      }
      umma_commit(bar0);
    }
    __syncwarp();
    if (lane == 0) mbar_wait(bar0, 0);
    __syncwarp();
    long long t1 = clock64();
    if (lane == 0) {
      out[blockIdx.x] = (unsigned long long)(t1 - t0);
      done = 1;
    }
  } else if (warp >= 2) {
    const int t = threadIdx.x - 64;  // 0..255
    const uint32_t dst0 = base + (t & 1) * 20816 + (t >> 1) * 16;  // two planes, like KC = 16 stages
    const uint4* src = gsrc + (size_t)blockIdx.x * 65536 + t;
    uint32_t acc = 0;
    unsigned long long n_ops = 0;
    while (!done) {
      if (bg == 1) {
#pragma unroll
        for (int k = 0; k < 10; ++k) cpa16(dst0 + k * 2080, src + k * 256);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      } else if (bg == 2) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      } else if (bg == 3) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + 120 * 1024 + t * 128 + ((k ^ (t & 7)) << 4)), "r"(acc) : "memory");
      } else if (bg == 4) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const uint4 v = __ldg(src + ((k * 256 + (n_ops & 63) * 4096) & 65535));
          acc ^= v.x;
        }
      } else if (bg == 5) {
#pragma unroll
        for (int k = 0; k < 8; ++k) gdst[(size_t)blockIdx.x * 65536 + ((t * 8 + k + (n_ops & 7) * 2048) & 65535)] = make_uint4(acc, acc, acc, acc);
      } else if (bg == 6) {
        if (t == 0) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 4096, [%2];"
                       ::"r"(base + 150 * 1024), "l"(src), "r"(bar0 + 8) : "memory");
        }
        __nanosleep(200);
      } else if (bg == 7) {
        // pure ALU pressure on every scheduler (dependent FMA chains, 4 per thread)
        float x0 = __uint_as_float(acc | 0x3f800000u), x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          x0 = fmaf(x0, 1.0001f, 0.5f); x1 = fmaf(x1, 1.0001f, 0.5f); x2 = fmaf(x2, 1.0001f, 0.5f); x3 = fmaf(x3, 1.0001f, 0.5f);
        }
        acc ^= __float_as_uint(x0 + x1 + x2 + x3);
      } else if (bg == 8) {
        // integer / address-style work with warp-uniform operands (lands on the uniform datapath)
        uint32_t u = (uint32_t)n_ops;
#pragma unroll
        for (int k = 0; k < 64; ++k) u = u * 2654435761u + (uint32_t)total + (u >> 7);
        acc ^= u;
      } else {
        __nanosleep(500);
      }
      ++n_ops;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acc == 0x1234567u) stop_flag[0] = (int)acc;
    if (t == 0) out[256 + blockIdx.x] = n_ops;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d_out;
  uint4 *gsrc, *gdst;
  int* flag;
  cudaMalloc(&d_out, 1024 * 8);
  cudaMalloc(&gsrc, (size_t)148 * 65536 * 16 + 65536 * 16);
  cudaMalloc(&gdst, (size_t)148 * 65536 * 16 + 65536 * 16);
  cudaMalloc(&flag, 4);
  cudaMemset(gsrc, 0, (size_t)148 * 65536 * 16);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[9] = {"none", "cp.async 16B planar scatter", "tcgen05.ld", "st.shared.v4 (swizzled rows)", "ld.global only",
                         "st.global.v4 (128 B lane stride)", "cp.async.bulk 4 KB", "FMA spin (ALU pressure)", "uniform integer spin"};
  const int total = 4080;
  for (int n : {48, 96, 192}) {
    for (int bg = 0; bg < 9; ++bg) {
      cudaMemset(d_out, 0, 1024 * 8);
      probe<<<148, 320, smem>>>(n, total, bg, gsrc, gdst, d_out, flag);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("n=%d bg=%d: %s\n", n, bg, cudaGetErrorString(e));
        return 1;
      }
      std::vector<unsigned long long> h(512);
      cudaMemcpy(h.data(), d_out, 512 * 8, cudaMemcpyDeviceToHost);
      double cyc = 0, ops = 0;
      for (int i = 0; i < 148; ++i) {
        cyc += (double)h[i];
        ops += (double)h[256 + i];
      }
      cyc /= 148.0;
      ops /= 148.0;
      printf("N=%3d  background %-34s: %6.1f cycles/MMA   (background iterations per thread per kcycle: %.2f)\n", n, names[bg],
             cyc / total, ops / (cyc / 1000.0));
    }
  }
  return 0;
}
