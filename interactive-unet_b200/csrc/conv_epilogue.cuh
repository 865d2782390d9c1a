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

// 16-bit pack with saturation to the finite range (one cvt per pair) and ReLU on the packed pair.
__device__ __forceinline__ uint32_t pack16_sat(float lo, float hi, int fp16) {
  uint32_t r;
  if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t relu16x2(uint32_t v, int fp16) {
  uint32_t r;
  const uint32_t zero = 0u;
  if (fp16) asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(zero));
  else asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(zero));
  return r;
}

// Address of class 0 of pixel (n, y, x) in the head's output and the stride between classes.
__device__ __forceinline__ void softmax_dst(const ConvArgs& a, int n, int y, int x, float** dst, size_t* cstride) {
  float* outp = reinterpret_cast<float*>(a.out);
  if (a.mode == kEpiSoftmaxNHWC) {
    // single-GPU layout (row_block == image height): no division; destination-major layout for the all-to-all otherwise
    const size_t rowoff = a.row_block == a.out_h
                              ? (size_t)(a.slice0 + n) * a.out_h + y
                              : ((size_t)(y / a.row_block) * a.slice_count + a.slice0 + n) * a.row_block + (y % a.row_block);
    *dst = outp + (rowoff * a.out_w + x) * a.num_classes;
    *cstride = 1;
  } else {
    const size_t plane = (size_t)a.out_h * a.out_w;
    *dst = outp + (size_t)n * a.num_classes * plane + (size_t)y * a.out_w + x;
    *cstride = plane;
  }
}

// softmax over NC logits (`unet.py:63,67`): exp(l - max) / sum, IEEE division; vector store when NHWC.
template <int NC>
__device__ __forceinline__ void softmax_store(const ConvArgs& a, const float* bias, const uint32_t (&acc)[4], int n,
                                              int y, int x) {
  float l[NC];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    l[j] = __uint_as_float(acc[j]) + bias[j];
    mx = fmaxf(mx, l[j]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    l[j] = expf(l[j] - mx);
    sum += l[j];
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) l[j] = __fdiv_rn(l[j], sum);
  float* dst;
  size_t cstride;
  softmax_dst(a, n, y, x, &dst, &cstride);
  if (a.mode == kEpiSoftmaxNHWC && NC == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(l[0], l[NC > 1 ? 1 : 0]);
  } else if (a.mode == kEpiSoftmaxNHWC && NC == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(l[0], l[NC > 1 ? 1 : 0], l[NC > 2 ? 2 : 0], l[NC > 3 ? 3 : 0]);
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) dst[j * cstride] = l[j];
  }
}

// Residual operand of one pixel: CHUNK channels (CHUNK / 8 16-byte vectors) starting at channel `c`.
template <int BN>
struct EpiCfg {
  static constexpr int CHUNK = BN >= 32 ? 32 : 16;  // accumulator columns per tcgen05.ld
  static constexpr int RV = CHUNK / 8;               // 16-byte vectors per chunk of 16-bit channels
};
template <int BN>
__device__ __forceinline__ void residual_load(const ConvArgs& a, size_t pix, int c, uint4 (&r)[EpiCfg<BN>::RV]) {
  const uint4* rp = reinterpret_cast<const uint4*>(a.residual + pix * a.cout + c);
#pragma unroll
  for (int j = 0; j < EpiCfg<BN>::RV; ++j) r[j] = __ldg(rp + j);
}
// Issued BEFORE waiting for the accumulator so that the residual's memory latency hides under the MMAs.
template <int BN>
__device__ __forceinline__ void residual_prefetch(const ConvArgs& a, int ntile, int n, int y, int x, bool valid,
                                                  uint4 (&r)[EpiCfg<BN>::RV]) {
  if (a.mode == kEpiBf16 && a.residual != nullptr && valid)
    residual_load<BN>(a, ((size_t)n * a.out_h + y) * a.out_w + x, ntile * BN, r);
}

// `taddr`: TMEM address of this warp's 32 lanes at the accumulator's first column; `res`: residual of the first
// chunk (residual_prefetch).  Must be called by all 32 lanes of the warp (tcgen05.ld is warp-collective);
// `valid` masks the loads / stores.
template <int BN>
__device__ __forceinline__ void epilogue_pixel(const ConvArgs& a, const float* bias, int ntile, uint32_t taddr, int n, int y, int x,
                                               bool valid, uint4 (&res)[EpiCfg<BN>::RV]) {
  constexpr int CHUNK = EpiCfg<BN>::CHUNK;
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
#pragma unroll
          for (int j = 0; j < CHUNK / 8; ++j) {
            const uint4 rv = res[j];
            const float2 r0 = unpack16(rv.x, a.fp16), r1 = unpack16(rv.y, a.fp16);
            const float2 r2 = unpack16(rv.z, a.fp16), r3 = unpack16(rv.w, a.fp16);
            v[8 * j + 0] += r0.x; v[8 * j + 1] += r0.y;
            v[8 * j + 2] += r1.x; v[8 * j + 3] += r1.y;
            v[8 * j + 4] += r2.x; v[8 * j + 5] += r2.y;
            v[8 * j + 6] += r3.x; v[8 * j + 7] += r3.y;
          }
          if (ch + 1 < BN / CHUNK) residual_load<BN>(a, pix, c + CHUNK, res);  // in flight during this chunk's stores
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
    // head: logits live in the first num_classes accumulator columns.  The class count picks a fixed-size
    // path (2 and 4 are the reference's common settings) so that only num_classes exponentials are evaluated.
    const int nc = a.num_classes;
    if (nc <= 4) {
      uint32_t acc[4];
      tmem_ld_32x4(taddr, acc);
      tmem_ld_wait();
      if (valid) {
        if (nc == 2) softmax_store<2>(a, bias, acc, n, y, x);
        else if (nc == 4) softmax_store<4>(a, bias, acc, n, y, x);
        else if (nc == 3) softmax_store<3>(a, bias, acc, n, y, x);
        else softmax_store<1>(a, bias, acc, n, y, x);
      }
    } else {
      uint32_t acc[16];
      tmem_ld_32x16(taddr, acc);
      tmem_ld_wait();
      if (valid) {
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
          if (j < nc) {
            l[j] = expf(l[j] - mx);
            sum += l[j];
          }
        }
        float* dst;
        size_t cstride;
        softmax_dst(a, n, y, x, &dst, &cstride);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nc) dst[j * cstride] = __fdiv_rn(l[j], sum);
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
