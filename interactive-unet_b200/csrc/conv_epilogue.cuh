// Fused epilogue shared by the tensor-core conv kernels: one thread owns one output pixel (one TMEM lane)
// and walks the accumulator columns: + folded-BN bias (+ residual) (+ ReLU) -> 16-bit NHWC store
// (optionally to the 2x2 nearest-upsampled positions), or softmax over the class logits -> fp32 store.
#pragma once
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace iu {

// 16-bit storage helpers: `fp16` selects IEEE half (clamped to the finite range) or bfloat16.
__device__ __forceinline__ uint32_t pack16(float lo, float hi, int fp16) {
  if (fp16) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -65504.0f), 65504.0f), fminf(fmaxf(hi, -65504.0f), 65504.0f));
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack16(uint32_t v, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<__half2*>(&v));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// `taddr`: TMEM address of this warp's 32 lanes at the accumulator's first column.
// Must be called by all 32 lanes of the warp (tcgen05.ld is warp-collective); `valid` masks the stores.
template <int BN>
__device__ __forceinline__ void epilogue_pixel(const ConvArgs& a, const float* bias, int ntile, uint32_t taddr, int n, int y, int x,
                                               bool valid) {
  constexpr int CHUNK = BN >= 32 ? 32 : 16;  // accumulator columns per tcgen05.ld
  if (a.mode == kEpiBf16) {
    const size_t pix = ((size_t)n * a.out_h + y) * a.out_w + x;
    const int col0 = ntile * BN;
#pragma unroll 1
    for (int ch = 0; ch < BN / CHUNK; ++ch) {
      uint32_t acc[CHUNK];
      if constexpr (CHUNK == 32) tmem_ld_32x32(taddr + ch * CHUNK, reinterpret_cast<uint32_t(&)[32]>(acc));
      else tmem_ld_32x16(taddr + ch * CHUNK, reinterpret_cast<uint32_t(&)[16]>(acc));
      tmem_ld_wait();
      if (valid) {
        const int c = col0 + ch * CHUNK;
        float v[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(bias + c + j);
          v[j] = __uint_as_float(acc[j]) + b.x;
          v[j + 1] = __uint_as_float(acc[j + 1]) + b.y;
          v[j + 2] = __uint_as_float(acc[j + 2]) + b.z;
          v[j + 3] = __uint_as_float(acc[j + 3]) + b.w;
        }
        if (a.residual != nullptr) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.residual + pix * a.cout + c);
#pragma unroll
          for (int j = 0; j < CHUNK / 8; ++j) {
            const uint4 rv = __ldg(rp + j);
            const float2 r0 = unpack16(rv.x, a.fp16), r1 = unpack16(rv.y, a.fp16);
            const float2 r2 = unpack16(rv.z, a.fp16), r3 = unpack16(rv.w, a.fp16);
            v[8 * j + 0] += r0.x; v[8 * j + 1] += r0.y;
            v[8 * j + 2] += r1.x; v[8 * j + 3] += r1.y;
            v[8 * j + 4] += r2.x; v[8 * j + 5] += r2.y;
            v[8 * j + 6] += r3.x; v[8 * j + 7] += r3.y;
          }
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        uint4 pk[CHUNK / 8];
#pragma unroll
        for (int j = 0; j < CHUNK / 8; ++j) {
          pk[j].x = pack16(v[8 * j + 0], v[8 * j + 1], a.fp16);
          pk[j].y = pack16(v[8 * j + 2], v[8 * j + 3], a.fp16);
          pk[j].z = pack16(v[8 * j + 4], v[8 * j + 5], a.fp16);
          pk[j].w = pack16(v[8 * j + 6], v[8 * j + 7], a.fp16);
        }
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out);
        if (!a.up2x) {
          uint4* dst = reinterpret_cast<uint4*>(outp + pix * a.cout + c);
#pragma unroll
          for (int j = 0; j < CHUNK / 8; ++j) dst[j] = pk[j];
        } else {
          const int oh = 2 * a.out_h, ow = 2 * a.out_w;
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const size_t up = ((size_t)n * oh + 2 * y + (d >> 1)) * ow + 2 * x + (d & 1);
            uint4* dst = reinterpret_cast<uint4*>(outp + up * a.cout + c);
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j) dst[j] = pk[j];
          }
        }
      }
    }
  } else {
    // head: logits live in the first num_classes accumulator columns
    uint32_t acc[16];
    tmem_ld_32x16(taddr, acc);
    tmem_ld_wait();
    if (valid) {
      const int nc = a.num_classes;
      float l[16];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        l[j] = (j < nc) ? __uint_as_float(acc[j]) + bias[j] : -INFINITY;
        mx = fmaxf(mx, l[j]);
      }
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        l[j] = (j < nc) ? expf(l[j] - mx) : 0.0f;
        sum += l[j];
      }
      float* outp = reinterpret_cast<float*>(a.out);
      if (a.mode == kEpiSoftmaxNHWC) {
        const size_t rowoff =
            ((size_t)(y / a.row_block) * a.slice_count + a.slice0 + n) * a.row_block + (y % a.row_block);
        float* dst = outp + (rowoff * a.out_w + x) * nc;
        if (nc == 2) {
          *reinterpret_cast<float2*>(dst) = make_float2(__fdiv_rn(l[0], sum), __fdiv_rn(l[1], sum));
        } else if (nc == 4) {
          *reinterpret_cast<float4*>(dst) =
              make_float4(__fdiv_rn(l[0], sum), __fdiv_rn(l[1], sum), __fdiv_rn(l[2], sum), __fdiv_rn(l[3], sum));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nc) dst[j] = __fdiv_rn(l[j], sum);
        }
      } else {
        const size_t plane = (size_t)a.out_h * a.out_w;
        float* dst = outp + (size_t)n * nc * plane + (size_t)y * a.out_w + x;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nc) dst[j * plane] = __fdiv_rn(l[j], sum);
      }
    }
  }
}

struct TileCoord {
  int x0, y0, n0, ntile;
};
// tile index -> (Cout tile, column block, row block, image group); Cout tiles vary fastest so that CTAs
// working on neighbouring tile indices share the same activation tile in L2.
__device__ __forceinline__ TileCoord decode_tile(const ConvArgs& a, int tile) {
  TileCoord t;
  t.ntile = tile % a.ntiles_n;
  const int m = tile / a.ntiles_n;
  t.x0 = (m % a.tiles_x) * a.tw;
  t.y0 = ((m / a.tiles_x) % a.tiles_y) * a.th;
  t.n0 = (m / (a.tiles_x * a.tiles_y)) * a.nb;
  return t;
}

}  // namespace iu
