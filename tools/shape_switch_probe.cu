// Does switching the MMA shape (instruction descriptor N) between consecutive tcgen05.mma cost anything?  (B200 probe)
// The row-folded kernel issues, per filter column, MMAs of N = CO, 2CO, 3CO, ..., 3CO, 2CO, CO.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I interactive-unet_b200/csrc tools/shape_switch_probe.cu -o /tmp/ip
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace iu;

__device__ __forceinline__ uint64_t desc_planar(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46;
}

// mode 0: 30 MMAs all N = 3*co          mode 1: row-fold order (co,2co,3co x6,2co,co) x 3, D and B offsets as in conv_row
// mode 2: same 30 MMAs grouped by shape (6 x co, 6 x 2co, 18 x 3co)
// mode 3: row-fold order but every MMA uses the SAME idesc (N = 3co; wrong maths, timing only)
// mode 4: row-fold order, same D address for all
template <int CO>
__global__ void __launch_bounds__(128) probe(int mode, int reps, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
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
    const uint32_t id1 = umma_idesc_f16(128, CO, 1), id2 = umma_idesc_f16(128, 2 * CO, 1), id3 = umma_idesc_f16(128, 3 * CO, 1);
    const uint64_t bdesc = umma_smem_desc<32>(base + 100 * 1024);
    const uint64_t a0 = desc_planar(base, 20816, 128);
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)bdesc, b_hi = (uint32_t)(bdesc >> 32);
    constexpr int R = 8;
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) {
#pragma unroll
          for (int k = 0; k < 30; ++k)
            umma_f16_lohi(tmem + (uint32_t)((k % 6) * CO), a_lo + (uint32_t)((k % 10) * 130 + k / 10), a_hi, b_lo, b_hi, id3, 1u);
        } else if (mode == 2) {
#pragma unroll
          for (int cls = 1; cls <= 3; ++cls) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int j = 0; j < R + 2; ++j) {
                const int lo = j >= 2 ? 0 : 2 - j, hi = j <= R - 1 ? 2 : R + 1 - j, ns = hi - lo + 1;
                if (ns != cls) continue;
                umma_f16_lohi(tmem + (uint32_t)((j - 2 + lo) * CO), a_lo + (uint32_t)(j * 130 + kx), a_hi,
                              b_lo + (uint32_t)((lo * CO * 32) >> 4) + (uint32_t)kx * ((3 * CO * 32) >> 4), b_hi,
                              ns == 3 ? id3 : (ns == 2 ? id2 : id1), 1u);
              }
            }
          }
        } else {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int j = 0; j < R + 2; ++j) {
              const int lo = j >= 2 ? 0 : 2 - j, hi = j <= R - 1 ? 2 : R + 1 - j, ns = hi - lo + 1;
              const uint32_t id = mode == 3 ? id3 : (ns == 3 ? id3 : (ns == 2 ? id2 : id1));
              const uint32_t d = mode == 4 ? tmem : tmem + (uint32_t)((j - 2 + lo) * CO);
              umma_f16_lohi(d, a_lo + (uint32_t)(j * 130 + kx), a_hi,
                            b_lo + (uint32_t)((lo * CO * 32) >> 4) + (uint32_t)kx * ((3 * CO * 32) >> 4), b_hi, id, 1u);
            }
          }
        }
        umma_commit(bar0 + 8);
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
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int CO>
void run(unsigned long long* d_out) {
  const int smem = 200 * 1024, reps = 128;
  cudaFuncSetAttribute(probe<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"30 x N=3co, one shape", "row-fold order (co,2co,3co..,2co,co)", "same MMAs grouped by shape",
                         "row-fold order, one idesc (3co)", "row-fold order, one D address"};
  for (int mode = 0; mode < 5; ++mode) {
    probe<CO><<<148, 128, smem>>>(mode, reps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("co=%d mode %d: %s\n", CO, mode, cudaGetErrorString(e));
      exit(1);
    }
    std::vector<unsigned long long> h(148);
    cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (auto v : h) cyc += (double)v;
    printf("CO=%2d  %-40s: %7.1f cycles per 30-MMA group (%5.1f per MMA)\n", CO, names[mode], cyc / 148.0 / reps,
           cyc / 148.0 / reps / 30.0);
  }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 1024 * 8);
  run<16>(d_out);
  run<32>(d_out);
  run<64>(d_out);
  return 0;
}
