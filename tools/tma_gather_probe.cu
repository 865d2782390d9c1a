// Can TMA produce the planar (8-channel plane) halo layout straight from an NHWC tensor?  (B200 probe)
// A 5-D view (c8 = 8 channels, x, y, n, plane) of NHWC has a 16-byte innermost run; the box (8, 130, ROWS, 1, PLANES)
// lands in shared memory exactly as conv_row.cu's A stage.  This measures the bytes per clock per SM it sustains,
// next to the 16-byte cp.async gather the kernels use today.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I interactive-unet_b200/csrc tools/tma_gather_probe.cu -o /tmp/tg -lcuda
#include <cuda.h>
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace iu;

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// every CTA streams `iters` boxes (rows x 130 px x planes) from its own region of a big NHWC tensor, 3 in flight
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(dst), "l"(src) : "memory");
}

// `with_cpasync`: 8 more warps run a 16-byte cp.async gather (same volume per box) concurrently -- do the two paths add up?
__global__ void __launch_bounds__(384) tma_gather(const __grid_constant__ CUtensorMap map, int rows, int planes, int iters, int h, int w,
                                                  unsigned long long* out, const uint4* gsrc, int with_cpasync, int with_tma) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  __shared__ uint64_t bars[4];
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t bytes = (uint32_t)rows * 130u * planes * 16u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  __shared__ volatile int tma_done;
  if (threadIdx.x == 0) tma_done = 0;
  __syncthreads();
  if (threadIdx.x >= 128) {
    if (with_cpasync) {
      const int t = threadIdx.x - 128;
      const uint32_t dst0 = base + 150 * 1024 + t * 16;
      const uint4* src = gsrc + (size_t)blockIdx.x * 65536 + t;
      const int per_box = rows * 130 * planes / 256 + 1;  // copies per thread per box
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        for (int k = 0; k < per_box; ++k) cpa16(dst0 + (k & 7) * 4096, src + ((i * per_box + k) * 256 & 65535));
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (t == 0) out[256 + blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    return;
  }
  if (threadIdx.x == 0 && with_tma) {
    long long t0 = clock64();
    const int blocks_y = h / 8, blocks_x = w / 128;
    for (int i = 0; i < iters + 3; ++i) {
      if (i >= 3) mbar_wait(bar0 + 8 * ((i - 3) & 3), ((i - 3) >> 2) & 1);
      if (i < iters) {
        const int t = blockIdx.x * iters + i;
        const int x0 = (t % blocks_x) * 128, y0 = ((t / blocks_x) % blocks_y) * 8, n = t / (blocks_x * blocks_y);
        mbar_arrive_expect_tx(bar0 + 8 * (i & 3), bytes);
        tma_load_5d(base + (i & 3) * 49152, &map, bar0 + 8 * (i & 3), 0, x0 - 1, y0 - 1, n, 0);
      }
    }
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn encode = (EncodeFn)fn;
  unsigned long long* d_out;
  cudaMalloc(&d_out, 1024 * 8);
  void* gsrc;
  cudaMalloc(&gsrc, (size_t)149 * 65536 * 16);
  cudaMemset(gsrc, 0, (size_t)149 * 65536 * 16);
  const int n = 74, h = 512, w = 512;
  for (int c : {16, 32, 64}) {
    void* src;
    const size_t bytes = (size_t)n * h * w * c * 2;
    cudaMalloc(&src, bytes);
    cudaMemset(src, 0, bytes);
    for (int rows : {6, 10}) {
      const int planes = c >= 32 ? 4 : 2;  // KC = 32 or 16 channels per stage
      CUtensorMap map;
      cuuint64_t dims[5] = {8, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, (cuuint64_t)(c / 8)};
      cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, 16};
      cuuint32_t box[5] = {8, 130, (cuuint32_t)rows, 1, (cuuint32_t)planes};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        printf("C=%d rows=%d: encode failed (%d)\n", c, rows, (int)r);
        continue;
      }
      const int iters = 64;
      cudaFuncSetAttribute(tma_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      const double box_bytes = (double)rows * 130 * planes * 16;
      if (box_bytes > 48 * 1024) continue;  // the probe's ring slots are 48 KB
      for (int mode = 0; mode < 3; ++mode) {  // 0: TMA only, 1: cp.async only, 2: both at once
        cudaMemset(d_out, 0, 1024 * 8);
        tma_gather<<<148, 384, 200 * 1024>>>(map, rows, planes, iters, h, w, d_out, (const uint4*)gsrc, mode >= 1, mode != 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("C=%d rows=%d: %s\n", c, rows, cudaGetErrorString(e));
          return 1;
        }
        std::vector<unsigned long long> hcyc(512);
        cudaMemcpy(hcyc.data(), d_out, 512 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0, cyc2 = 0;
        for (int i = 0; i < 148; ++i) { cyc += (double)hcyc[i]; cyc2 += (double)hcyc[256 + i]; }
        cyc /= 148.0; cyc2 /= 148.0;
        const double cp_bytes = (double)((int)(rows * 130 * planes / 256) + 1) * 256 * 16;
        printf("C=%2d  box %2d rows x %d planes (%5.1f KB)  %-13s: TMA %5.1f B/clk/SM   cp.async %5.1f B/clk/SM\n", c, rows, planes,
               box_bytes / 1024.0, mode == 0 ? "TMA only" : (mode == 1 ? "cp.async only" : "both"),
               cyc > 0 ? box_bytes * iters / cyc : 0.0, cyc2 > 0 ? cp_bytes * iters / cyc2 : 0.0);
      }
    }
    cudaFree(src);
  }
  return 0;
}
