// HBM-bound kernels around the conv stack: slice gather (K1), max-pool, and the fused
// cross-axis reduce / quantise / argmax (K4).  All arithmetic that the reference does on the host
// in fp32 (`predict.py:110,237,244-245,255`) is reproduced with round-to-nearest intrinsics so that
// no FMA contraction or reciprocal substitution can change a bit.
#include "aux_kernels.cuh"

#include <cstdlib>

namespace iu {

// =========================================================================== K1 gather
__device__ __forceinline__ float norm_val(uint8_t v) { return __fdiv_rn((float)v, 255.0f); }
__device__ __forceinline__ float norm_val(float v) { return v; }

// A slice source is described by strides (in elements): element (slice b, row r, col c) lives at
//   base[b * s_slice + r * s_row + c * s_col].
// Cubic volume [z][y][x] of edge n: axis 0 -> (n*n, n, 1), axis 1 -> (n, n*n, 1), axis 2 -> (1, n*n, n); the z-slab /
// strip buffers of the multi-GPU path (distributed.py) use the same three shapes with other extents.
struct SliceSrc {
  const void* base;
  long long s_slice, s_row, s_col;
  int count, h, w;
};

// rows contiguous in the source (s_col == 1): 16-element vector copy per thread
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const SliceSrc s, float* __restrict__ out) {
  const T* __restrict__ vol = static_cast<const T*>(s.base);
  const int groups_per_row = s.w / 16;
  const size_t total = (size_t)s.count * s.h * groups_per_row;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const int cg = (int)(g % groups_per_row);
    const int r = (int)((g / groups_per_row) % s.h);
    const int b = (int)(g / ((size_t)groups_per_row * s.h));
    const T* src = vol + (size_t)b * s.s_slice + (size_t)r * s.s_row + cg * 16;
    float* dst = out + ((size_t)b * s.h + r) * s.w + cg * 16;
    float v[16];
    if constexpr (sizeof(T) == 1) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src));
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = norm_val((uint8_t)((w[j >> 2] >> (8 * (j & 3))) & 0xff));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(src) + j);
        v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}

// slices contiguous in the source (s_slice == 1): transpose a [64 cols][32 slices] tile in smem
template <typename T>
__global__ void __launch_bounds__(256) gather_cols_kernel(const SliceSrc s, float* __restrict__ out) {
  __shared__ float tile[64][33];
  const T* __restrict__ vol = static_cast<const T*>(s.base);
  const int c0 = blockIdx.x * 64;
  const int r = blockIdx.y;
  for (int b0 = 0; b0 < s.count; b0 += 32) {
    const int bl = threadIdx.x & 31;
    for (int cl = threadIdx.x >> 5; cl < 64; cl += 8) {
      if (b0 + bl < s.count && c0 + cl < s.w)
        tile[cl][bl] = norm_val(vol[(size_t)r * s.s_row + (size_t)(c0 + cl) * s.s_col + b0 + bl]);
    }
    __syncthreads();
    const int cl = threadIdx.x & 63;
    for (int b = threadIdx.x >> 6; b < 32; b += 4) {
      if (b0 + b < s.count && c0 + cl < s.w) out[((size_t)(b0 + b) * s.h + r) * s.w + c0 + cl] = tile[cl][b];
    }
    __syncthreads();
  }
}

// any other stride pattern: one element per thread (correct for every source, fast for none)
template <typename T>
__global__ void __launch_bounds__(256) gather_any_kernel(const SliceSrc s, float* __restrict__ out) {
  const T* __restrict__ vol = static_cast<const T*>(s.base);
  const size_t total = (size_t)s.count * s.h * s.w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % s.w);
    const int r = (int)((i / s.w) % s.h);
    const size_t b = i / ((size_t)s.w * s.h);
    out[i] = norm_val(vol[b * s.s_slice + (size_t)r * s.s_row + (size_t)c * s.s_col]);
  }
}

cudaError_t launch_gather_strided(const void* base, int is_f32, int count, int h, int w, long long s_slice,
                                  long long s_row, long long s_col, float* out, cudaStream_t stream) {
  if (count <= 0 || h <= 0 || w <= 0 || s_slice < 0 || s_row < 0 || s_col < 0) return cudaErrorInvalidValue;
  SliceSrc s{base, s_slice, s_row, s_col, count, h, w};
  const size_t esz = is_f32 ? 4 : 1;
  const bool vec_ok = s_col == 1 && w % 16 == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0 &&
                      ((size_t)s_slice * esz) % 16 == 0 && ((size_t)s_row * esz) % 16 == 0;
  if (vec_ok) {
    const size_t groups = (size_t)count * h * (w / 16);
    const int blocks = (int)((groups + 255) / 256 < 148 * 16 ? (groups + 255) / 256 : 148 * 16);
    if (is_f32) gather_rows_kernel<float><<<blocks, 256, 0, stream>>>(s, out);
    else gather_rows_kernel<uint8_t><<<blocks, 256, 0, stream>>>(s, out);
  } else if (s_slice == 1 && h <= 65535) {
    dim3 grid((w + 63) / 64, h);
    if (is_f32) gather_cols_kernel<float><<<grid, 256, 0, stream>>>(s, out);
    else gather_cols_kernel<uint8_t><<<grid, 256, 0, stream>>>(s, out);
  } else {
    const size_t total = (size_t)count * h * w;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    if (is_f32) gather_any_kernel<float><<<blocks, 256, 0, stream>>>(s, out);
    else gather_any_kernel<uint8_t><<<blocks, 256, 0, stream>>>(s, out);
  }
  return cudaGetLastError();
}

cudaError_t launch_gather_slices(const void* vol, int vol_is_f32, int n, int axis, int start, int count, float* out,
                                 cudaStream_t stream) {
  if (n % 16 != 0 || axis < 0 || axis > 2 || count <= 0) return cudaErrorInvalidValue;
  const long long nn = (long long)n * n;
  const long long ss = axis == 0 ? nn : (axis == 1 ? n : 1), sr = axis == 0 ? n : nn, sc = axis == 2 ? n : 1;
  const char* base = static_cast<const char*>(vol) + (size_t)start * ss * (vol_is_f32 ? 4 : 1);
  return launch_gather_strided(base, vol_is_f32, count, n, n, ss, sr, sc, out, stream);
}

// =========================================================================== max-pool 3x3/s2/p1 (NHWC 16-bit, values >= 0)
// The generic form: one output pixel x 8 channels per thread, nine 16-byte loads.  Used where the channel count is not 64.
__global__ void __launch_bounds__(256) maxpool_kernel(const __nv_bfloat16* __restrict__ in, int batch, int h, int w,
                                                      int c, __nv_bfloat16* __restrict__ out) {
  const int oh = h / 2, ow = w / 2, groups = c / 8;
  const size_t total = (size_t)batch * oh * ow * groups;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int ox = (int)((i / groups) % ow);
    const int oy = (int)((i / ((size_t)groups * ow)) % oh);
    const int n = (int)(i / ((size_t)groups * ow * oh));
    uint32_t m[4] = {0u, 0u, 0u, 0u};   // +0.0 in both formats: identity of max over non-negative values
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = 2 * oy - 1 + r;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = 2 * ox - 1 + s;
        if (ix < 0 || ix >= w) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * h + iy) * w + ix) * c + g * 8));
        m[0] = __vmaxu2(m[0], v.x);
        m[1] = __vmaxu2(m[1], v.y);
        m[2] = __vmaxu2(m[2], v.z);
        m[3] = __vmaxu2(m[3], v.w);
      }
    }
    *reinterpret_cast<uint4*>(out + (((size_t)n * oh + oy) * ow + ox) * c + g * 8) = make_uint4(m[0], m[1], m[2], m[3]);
  }
}

// The encoder's pool (64 channels): a thread owns a 2 x 2 patch of output pixels x 8 channels.  Its 5 x 5 input
// window is read once (25 loads for four outputs instead of 36), each row's two horizontal 3-maxima feed the upper
// and / or lower output row, and the block -- 4 x 8 patches x 8 channel groups = 8 x 16 output pixels of one image --
// is addressed by blockIdx alone: no 64-bit divisions (the generic form spends more issue slots on index arithmetic
// than on the pooling: 59 % busy at 48 % of the DRAM peak in ncu).
__device__ __forceinline__ uint4 vmax4(const uint4& a, const uint4& b) {
  return make_uint4(__vmaxu2(a.x, b.x), __vmaxu2(a.y, b.y), __vmaxu2(a.z, b.z), __vmaxu2(a.w, b.w));
}
__global__ void __launch_bounds__(256) maxpool64_kernel(const __nv_bfloat16* __restrict__ in, int h, int w,
                                                        __nv_bfloat16* __restrict__ out) {
  const int oh = h >> 1, ow = w >> 1;
  const int g = threadIdx.x & 7, px = (threadIdx.x >> 3) & 7, py = threadIdx.x >> 6;
  const int n = blockIdx.z;
  const int oy = blockIdx.y * 8 + 2 * py, ox = blockIdx.x * 16 + 2 * px;  // top-left output pixel of the patch
  if (oy >= oh || ox >= ow) return;
  const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
  const __nv_bfloat16* img = in + (size_t)n * h * w * 64 + g * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);  // +0.0 in both formats: identity of max over non-negative values
  uint4 top_l = zero, top_r = zero, bot_l = zero, bot_r = zero;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int iy = iy0 + r;
    if ((unsigned)iy >= (unsigned)h) continue;
    const __nv_bfloat16* row = img + (size_t)iy * w * 64;
    uint4 v[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const int ix = ix0 + s;
      v[s] = (unsigned)ix < (unsigned)w ? __ldg(reinterpret_cast<const uint4*>(row + (size_t)ix * 64)) : zero;
    }
    const uint4 hl = vmax4(vmax4(v[0], v[1]), v[2]), hr = vmax4(vmax4(v[2], v[3]), v[4]);
    if (r <= 2) { top_l = vmax4(top_l, hl); top_r = vmax4(top_r, hr); }
    if (r >= 2) { bot_l = vmax4(bot_l, hl); bot_r = vmax4(bot_r, hr); }
  }
  __nv_bfloat16* o = out + (((size_t)n * oh + oy) * ow + ox) * 64 + g * 8;
  const bool right = ox + 1 < ow, below = oy + 1 < oh;
  *reinterpret_cast<uint4*>(o) = top_l;
  if (right) *reinterpret_cast<uint4*>(o + 64) = top_r;
  if (below) {
    *reinterpret_cast<uint4*>(o + (size_t)ow * 64) = bot_l;
    if (right) *reinterpret_cast<uint4*>(o + (size_t)ow * 64 + 64) = bot_r;
  }
}

cudaError_t launch_maxpool(const __nv_bfloat16* in, int batch, int h, int w, int c, __nv_bfloat16* out,
                           cudaStream_t stream) {
  if (c % 8 != 0) return cudaErrorInvalidValue;
  static const bool patch = [] { const char* v = getenv("IU_POOL_PATCH"); return v == nullptr || atoi(v) != 0; }();  // development switch
  if (patch && c == 64 && batch <= 65535 && h >= 2 && w >= 2) {
    const dim3 grid((w / 2 + 15) / 16, (h / 2 + 7) / 8, batch);
    maxpool64_kernel<<<grid, 256, 0, stream>>>(in, h, w, out);
    return cudaGetLastError();
  }
  const size_t total = (size_t)batch * (h / 2) * (w / 2) * (c / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  maxpool_kernel<<<blocks, 256, 0, stream>>>(in, batch, h, w, c, out);
  return cudaGetLastError();
}

// =========================================================================== K4 reduce + quantise + argmax
// Block = one z, a 32(y) x 32(x) tile.  The axis-2 operand is x-major in memory, so its tile is staged
// through shared memory (coalesced along y) and read back transposed; axes 0 and 1 are read directly.
// VEC (C = 2 / 4, 16-byte aligned buffers): one 8- / 16-byte load per voxel and operand, packed uint8 stores.
template <int C, bool VEC>
__device__ __forceinline__ void load_classes(const float* __restrict__ s, float (&p)[C]) {
  if constexpr (VEC && C == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(s));
    p[0] = v.x; p[1] = v.y;
  } else if constexpr (VEC && C == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(s));
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] = __ldg(s + c);
  }
}
template <int C, bool VEC>
__device__ __forceinline__ void store_classes(float* __restrict__ d, const float (&p)[C]) {
  if constexpr (VEC && C == 2) {
    *reinterpret_cast<float2*>(d) = make_float2(p[0], p[1]);
  } else if constexpr (VEC && C == 4) {
    *reinterpret_cast<float4*>(d) = make_float4(p[0], p[1], p[2], p[3]);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = p[c];
  }
}
template <int C, bool VEC>
__device__ __forceinline__ void store_bytes(uint8_t* __restrict__ d, const uint8_t (&q)[C]) {
  if constexpr (VEC && C == 2) {
    *reinterpret_cast<uchar2*>(d) = make_uchar2(q[0], q[1]);
  } else if constexpr (VEC && C == 4) {
    *reinterpret_cast<uchar4*>(d) = make_uchar4(q[0], q[1], q[2], q[3]);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = q[c];
  }
}

template <int C, bool VEC>
__global__ void __launch_bounds__(256) reduce_kernel(const ReduceArgs a) {
  extern __shared__ float tile[];  // [32 x][32*C + 1]
  constexpr int pitch = 32 * C + 1;
  const int z = blockIdx.z + a.zoff;
  const int y0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
  const int n = a.n, t = a.t;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;

  if (a.p[2] != nullptr) {
    for (int xl = wrp; xl < 32; xl += 8) {
      if (x0 + xl < n) {
        const float* src = a.p[2] + (((size_t)(x0 + xl) * t + z) * n + y0) * C;
        const int lim = min(32, n - y0) * C;
        if (VEC && lim == 32 * C) {
          // full tile row: 32*C floats as 16-byte loads (8*C lanes), scattered into the odd-pitch tile
          if (lane < 8 * C) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + lane);
            float* d = tile + xl * pitch + 4 * lane;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
          }
        } else {
          for (int j = lane; j < lim; j += 32) tile[xl * pitch + j] = __ldg(src + j);
        }
      }
    }
  }
  __syncthreads();

  const int x = x0 + lane;
  if (x >= n) return;
  // window = clip((g[z]*g[y])*g[x] / gmax, lo, 1) (predict.py:327-347); gmax is 1.0f for every window the host builds
  // (g is normalised to a maximum of exactly 1), and x / 1.0f == x bit for bit, so that division is skipped then
  const bool windowed = a.g1d != nullptr;
  const float gz = windowed ? __ldg(a.g1d + a.z0 + z) : 0.0f;
  const float gx = windowed ? __ldg(a.g1d + x) : 0.0f;
  const bool unit_gmax = a.gmax == 1.0f;
  const float axes_f = (float)a.n_axes;
#pragma unroll 2
  for (int yl = wrp; yl < 32; yl += 8) {
    const int y = y0 + yl;
    if (y >= n) break;
    float p[3][C];
    if (a.p[0] != nullptr) load_classes<C, VEC>(a.p[0] + (((size_t)z * n + y) * n + x) * C, p[0]);
    if (a.p[1] != nullptr) load_classes<C, VEC>(a.p[1] + (((size_t)y * t + z) * n + x) * C, p[1]);
    if (a.p[2] != nullptr) {
#pragma unroll
      for (int c = 0; c < C; ++c) p[2][c] = tile[lane * pitch + yl * C + c];
    }
    float m[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float acc = 0.0f;
      for (int i = 0; i < a.n_axes; ++i) {
        const int ax = a.order[i];
        const float v = ax == 0 ? p[0][c] : (ax == 1 ? p[1][c] : p[2][c]);
        acc = __fadd_rn(acc, v);                       // predict.py:101-106, in the caller's axis order
      }
      m[c] = __fdiv_rn(acc, axes_f);               // predict.py:110
    }
    const size_t vox = ((size_t)z * n + y) * n + x;
    if (a.out_mean != nullptr) store_classes<C, VEC>(a.out_mean + vox * C, m);
    if (a.out_labels != nullptr) {
      int best = 0;
#pragma unroll
      for (int c = 1; c < C; ++c)
        if (m[c] > m[best]) best = c;                  // first maximum wins (np.argmax, predict.py:38)
      a.out_labels[vox] = (uint8_t)best;
    }
    float wgt = 0.0f;
    if (windowed) {
      wgt = __fmul_rn(__fmul_rn(gz, __ldg(a.g1d + y)), gx);
      if (!unit_gmax) wgt = __fdiv_rn(wgt, a.gmax);
      wgt = fminf(fmaxf(wgt, a.lo), 1.0f);             // predict.py:345
    }
    if (a.blend_pred != nullptr) {
      if (z >= a.l0[0] && z < a.l1[0] && y >= a.l0[1] && y < a.l1[1] && x >= a.l0[2] && x < a.l1[2]) {
        const int gz = a.gd_ring > 0 ? (a.b0[0] + z) % a.gd_ring : a.b0[0] + z;
        const size_t g = ((size_t)gz * a.gh + (a.b0[1] + y)) * a.gw + (a.b0[2] + x);
        float acc[C];
        load_classes<C, false>(a.blend_pred + g * C, acc);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(m[c], wgt));   // predict.py:244
        store_classes<C, false>(a.blend_pred + g * C, acc);
        a.blend_weight[g] = __fadd_rn(a.blend_weight[g], wgt);                          // predict.py:245
      }
    }
    if (a.out_u8 != nullptr) {
      uint8_t q[C];
      if (windowed) {
        const float den = fmaxf(wgt, 1e-3f);           // predict.py:253,255
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float pred = __fmul_rn(m[c], wgt);     // predict.py:244
          q[c] = (uint8_t)(int)__fdiv_rn(__fmul_rn(255.0f, pred), den);
        }
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) q[c] = (uint8_t)(int)__fmul_rn(255.0f, m[c]);
      }
      store_bytes<C, VEC>(a.out_u8 + vox * C, q);
    }
  }
}

// Single-block mode (no blend accumulators), 2 or 4 classes, even n, 16-byte aligned buffers: a 32(y) x 64(x) tile per
// block and TWO x-adjacent voxels per thread -- 16 / 32-byte loads per operand, packed 4 / 8-byte uint8 stores, 2-byte
// label stores, half the address arithmetic per voxel.  Same per-voxel arithmetic (and bits) as reduce_kernel.
template <int C>
__global__ void __launch_bounds__(256) reduce_pair_kernel(const ReduceArgs a) {
  extern __shared__ float tile[];  // [64 x][32*C + 1]
  constexpr int pitch = 32 * C + 1;
  const int z = blockIdx.z + a.zoff;
  const int y0 = blockIdx.y * 32, x0 = blockIdx.x * 64;
  const int n = a.n, t = a.t;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (a.p[2] != nullptr) {
    const int ylim = min(32, n - y0);
    for (int xl = wrp; xl < 64; xl += 8) {
      if (x0 + xl < n) {
        const float* src = a.p[2] + (((size_t)(x0 + xl) * t + z) * n + y0) * C;
        if (ylim == 32) {
          if (lane < 8 * C) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + lane);
            float* d = tile + xl * pitch + 4 * lane;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
          }
        } else {
          for (int j = lane; j < ylim * C; j += 32) tile[xl * pitch + j] = __ldg(src + j);
        }
      }
    }
  }
  __syncthreads();
  const int xl = 2 * lane, x = x0 + xl;
  if (x >= n) return;                      // n is even: a pair is inside or outside as a whole
  const bool windowed = a.g1d != nullptr;
  const float gz = windowed ? __ldg(a.g1d + a.z0 + z) : 0.0f;
  const float gx[2] = {windowed ? __ldg(a.g1d + x) : 0.0f, windowed ? __ldg(a.g1d + x + 1) : 0.0f};
  const bool unit_gmax = a.gmax == 1.0f;
  const float axes_f = (float)a.n_axes;
#pragma unroll 2
  for (int yl = wrp; yl < 32; yl += 8) {
    const int y = y0 + yl;
    if (y >= n) break;
    float p[3][2][C];
    if (a.p[0] != nullptr) {
      const float4* s = reinterpret_cast<const float4*>(a.p[0] + (((size_t)z * n + y) * n + x) * C);
#pragma unroll
      for (int k = 0; k < C / 2; ++k) {
        const float4 v = __ldg(s + k);
        float* d = &p[0][0][0] + 4 * k;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    }
    if (a.p[1] != nullptr) {
      const float4* s = reinterpret_cast<const float4*>(a.p[1] + (((size_t)y * t + z) * n + x) * C);
#pragma unroll
      for (int k = 0; k < C / 2; ++k) {
        const float4 v = __ldg(s + k);
        float* d = &p[1][0][0] + 4 * k;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    }
    if (a.p[2] != nullptr) {
#pragma unroll
      for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int c = 0; c < C; ++c) p[2][v][c] = tile[(xl + v) * pitch + yl * C + c];
    }
    const float gzy = windowed ? __fmul_rn(gz, __ldg(a.g1d + y)) : 0.0f;
    float m[2][C];
    uint8_t q[2][C];
    uint8_t lab[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float acc = 0.0f;
        for (int i = 0; i < a.n_axes; ++i) {
          const int ax = a.order[i];
          const float val = ax == 0 ? p[0][v][c] : (ax == 1 ? p[1][v][c] : p[2][v][c]);
          acc = __fadd_rn(acc, val);                   // predict.py:101-106, in the caller's axis order
        }
        m[v][c] = __fdiv_rn(acc, axes_f);              // predict.py:110
      }
      int best = 0;
#pragma unroll
      for (int c = 1; c < C; ++c)
        if (m[v][c] > m[v][best]) best = c;            // first maximum wins (np.argmax, predict.py:38)
      lab[v] = (uint8_t)best;
      if (windowed) {
        float wgt = __fmul_rn(gzy, gx[v]);
        if (!unit_gmax) wgt = __fdiv_rn(wgt, a.gmax);
        wgt = fminf(fmaxf(wgt, a.lo), 1.0f);           // predict.py:345
        const float den = fmaxf(wgt, 1e-3f);           // predict.py:253,255
#pragma unroll
        for (int c = 0; c < C; ++c) q[v][c] = (uint8_t)(int)__fdiv_rn(__fmul_rn(255.0f, __fmul_rn(m[v][c], wgt)), den);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) q[v][c] = (uint8_t)(int)__fmul_rn(255.0f, m[v][c]);
      }
    }
    const size_t vox = ((size_t)z * n + y) * n + x;
    if (a.out_mean != nullptr) {
      float4* d = reinterpret_cast<float4*>(a.out_mean + vox * C);
#pragma unroll
      for (int k = 0; k < C / 2; ++k) {
        const float* sm = &m[0][0] + 4 * k;
        d[k] = make_float4(sm[0], sm[1], sm[2], sm[3]);
      }
    }
    if (a.out_labels != nullptr) *reinterpret_cast<uchar2*>(a.out_labels + vox) = make_uchar2(lab[0], lab[1]);
    if (a.out_u8 != nullptr) {
      if constexpr (C == 2) {
        *reinterpret_cast<uchar4*>(a.out_u8 + vox * C) = make_uchar4(q[0][0], q[0][1], q[1][0], q[1][1]);
      } else {
        uint2 w;
        w.x = q[0][0] | (q[0][1] << 8) | (q[0][2] << 16) | ((uint32_t)q[0][3] << 24);
        w.y = q[1][0] | (q[1][1] << 8) | (q[1][2] << 16) | ((uint32_t)q[1][3] << 24);
        *reinterpret_cast<uint2*>(a.out_u8 + vox * C) = w;
      }
    }
  }
}

// =========================================================================== tiled mode: block extraction / finalise
// numpy 'reflect' of index i against a run of length len (period 2*(len-1), the edge sample is not repeated)
__device__ __forceinline__ int reflect_index(int i, int len) {
  if (len <= 1) return 0;
  const int period = 2 * (len - 1);
  int m = i % period;
  if (m < 0) m += period;
  return m < len ? m : period - m;
}

__global__ void __launch_bounds__(256) extract_block_kernel(const uint8_t* __restrict__ vol, int d, int h, int w, int i0,
                                                            int j0, int k0, int s, uint8_t* __restrict__ out) {
  // clipped box (predict.py:300-303); reflection is relative to the CLIPPED block, as np.pad sees it (predict.py:313)
  const int ci0 = max(i0, 0), ci1 = min(i0 + s, d), cj0 = max(j0, 0), cj1 = min(j0 + s, h), ck0 = max(k0, 0), ck1 = min(k0 + s, w);
  const size_t total = (size_t)s * s * s;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(o % s), y = (int)((o / s) % s), z = (int)(o / ((size_t)s * s));
    const int gz = ci0 + reflect_index(i0 + z - ci0, ci1 - ci0);
    const int gy = cj0 + reflect_index(j0 + y - cj0, cj1 - cj0);
    const int gx = ck0 + reflect_index(k0 + x - ck0, ck1 - ck0);
    out[o] = __ldg(vol + ((size_t)gz * h + gy) * w + gx);
  }
}

cudaError_t launch_extract_block(const uint8_t* vol, int d, int h, int w, int i0, int j0, int k0, int s, uint8_t* out,
                                 cudaStream_t stream) {
  if (s < 1 || i0 >= d || j0 >= h || k0 >= w || i0 + s <= 0 || j0 + s <= 0 || k0 + s <= 0) return cudaErrorInvalidValue;
  const size_t total = (size_t)s * s * s;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  extract_block_kernel<<<blocks, 256, 0, stream>>>(vol, d, h, w, i0, j0, k0, s, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) finalise_kernel(const float* __restrict__ pred, const float* __restrict__ weight,
                                                       size_t voxels, int c, uint8_t* __restrict__ out_u8,
                                                       uint8_t* __restrict__ out_labels) {
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < voxels; v += (size_t)gridDim.x * blockDim.x) {
    const float den = fmaxf(__ldg(weight + v), 1e-3f);   // predict.py:253,255
    int best = 0;
    float best_p = 0.0f;
    for (int k = 0; k < c; ++k) {
      const float p = __ldg(pred + v * c + k);
      if (out_u8 != nullptr) out_u8[v * c + k] = (uint8_t)(int)__fdiv_rn(__fmul_rn(255.0f, p), den);
      if (k == 0 || p > best_p) {
        best = k == 0 ? 0 : k;
        best_p = p;
      }
    }
    if (out_labels != nullptr) out_labels[v] = (uint8_t)best;
  }
}

cudaError_t launch_finalise(const float* pred, const float* weight, size_t voxels, int num_classes, uint8_t* out_u8,
                            uint8_t* out_labels, cudaStream_t stream) {
  if (num_classes < 1) return cudaErrorInvalidValue;
  const int blocks = (int)((voxels + 255) / 256 < 148 * 16 ? (voxels + 255) / 256 : 148 * 16);
  finalise_kernel<<<blocks, 256, 0, stream>>>(pred, weight, voxels, num_classes, out_u8, out_labels);
  return cudaGetLastError();
}

cudaError_t launch_reduce(const ReduceArgs& args, cudaStream_t stream) {
  if (args.n_axes < 1 || args.n_axes > 3) return cudaErrorInvalidValue;
  if (args.zoff < 0 || args.zcount < 0 || args.zoff + args.zcount > args.t) return cudaErrorInvalidValue;
  dim3 grid((args.n + 31) / 32, (args.n + 31) / 32, args.zcount ? args.zcount : args.t);
  const int c = args.num_classes;
  const size_t smem = (size_t)32 * (32 * c + 1) * sizeof(float);
  // vector path: 2 / 4 classes with every buffer aligned to its widest access
  auto aligned = [](const void* p, size_t a) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % a == 0; };
  const bool vec = (c == 2 || c == 4) && aligned(args.p[0], 16) && aligned(args.p[1], 16) && aligned(args.p[2], 16) &&
                   aligned(args.out_mean, 16) && aligned(args.out_u8, 4);
  // two voxels per thread: single-block mode, even edge, every buffer aligned to its widest access
  if (vec && args.blend_pred == nullptr && args.n % 2 == 0 && aligned(args.out_u8, 8) && aligned(args.out_labels, 2)) {
    dim3 grid2((args.n + 63) / 64, (args.n + 31) / 32, args.zcount ? args.zcount : args.t);
    const size_t smem2 = (size_t)64 * (32 * c + 1) * sizeof(float);
    if (c == 2) {
      reduce_pair_kernel<2><<<grid2, 256, smem2, stream>>>(args);
    } else {
      cudaFuncSetAttribute(reduce_pair_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      reduce_pair_kernel<4><<<grid2, 256, smem2, stream>>>(args);
    }
    return cudaGetLastError();
  }
#define IU_REDUCE_LAUNCH(C_, V_)                                                                         \
  cudaFuncSetAttribute(reduce_kernel<C_, V_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  reduce_kernel<C_, V_><<<grid, 256, smem, stream>>>(args);
#define IU_REDUCE_CASE(C_) \
  case C_:                 \
    IU_REDUCE_LAUNCH(C_, false) break;
  switch (c) {
    IU_REDUCE_CASE(1)
    case 2:
      if (vec) { IU_REDUCE_LAUNCH(2, true) } else { IU_REDUCE_LAUNCH(2, false) }
      break;
    IU_REDUCE_CASE(3)
    case 4:
      if (vec) { IU_REDUCE_LAUNCH(4, true) } else { IU_REDUCE_LAUNCH(4, false) }
      break;
    IU_REDUCE_CASE(5)
    IU_REDUCE_CASE(6)
    IU_REDUCE_CASE(7)
    IU_REDUCE_CASE(8)
    IU_REDUCE_CASE(9)
    IU_REDUCE_CASE(10)
    default:
      return cudaErrorInvalidValue;
  }
#undef IU_REDUCE_CASE
#undef IU_REDUCE_LAUNCH
  return cudaGetLastError();
}

// =========================================================================== Zarr staging: chunk layout, pyramid
// `elem` = bytes per voxel in the volume, `celem` >= elem = bytes per voxel in a chunk (a chunk's class extent may
// exceed the array's: pyramid levels keep level 0's chunk shape while their class axis is halved, utils.py:66-71).
template <typename V>
__global__ void __launch_bounds__(256) chunk_layout_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           int d, int h, int w, int elem, int celem, int cz, int cy,
                                                           int cx, int gy, int gx, size_t total, bool to_chunks) {
  constexpr int VEC = (int)sizeof(V);
  constexpr int UNROLL = 4;  // independent loads in flight per thread before the first store
  const int row_units = cx * celem / VEC;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * UNROLL) {
    size_t vol_off[UNROLL];
    bool inside[UNROLL];
    V v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const size_t i = i0 + k * stride;
      const int u = (int)(i % row_units);
      size_t r = i / row_units;
      const int y = (int)(r % cy);
      r /= cy;
      const int z = (int)(r % cz);
      const size_t chunk = r / cz;
      const int gxi = (int)(chunk % gx), gyi = (int)((chunk / gx) % gy), gzi = (int)(chunk / ((size_t)gx * gy));
      const int Z = gzi * cz + z, Y = gyi * cy + y;
      // vector path: celem == elem, so a row of a chunk is a contiguous piece of a volume row
      const int xb = u * VEC;                                    // byte inside the chunk row
      const int X = gxi * cx + (VEC == 1 ? xb / celem : 0);      // byte path: voxel and byte inside the voxel
      const int b = VEC == 1 ? xb % celem : xb;
      const size_t xbyte = (size_t)X * elem + b;
      inside[k] = i < total && Z < d && Y < h && (VEC == 1 ? (X < w && b < elem) : xbyte < (size_t)w * elem);
      vol_off[k] = ((size_t)Z * h + Y) * w * elem + xbyte;
      v[k] = V();
      if (to_chunks) {
        if (inside[k]) v[k] = *reinterpret_cast<const V*>(src + vol_off[k]);
      } else if (inside[k]) {
        v[k] = *reinterpret_cast<const V*>(src + i * VEC);
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const size_t i = i0 + k * stride;
      if (to_chunks) {
        if (i < total) *reinterpret_cast<V*>(dst + i * VEC) = v[k];
      } else if (inside[k]) {
        *reinterpret_cast<V*>(dst + vol_off[k]) = v[k];
      }
    }
  }
}

cudaError_t launch_chunk_layout(const uint8_t* src, uint8_t* dst, int d, int h, int w, int elem, int celem, int cz, int cy,
                                int cx, bool to_chunks, cudaStream_t stream) {
  if (d < 1 || h < 1 || w < 1 || elem < 1 || celem < elem || cz < 1 || cy < 1 || cx < 1) return cudaErrorInvalidValue;
  const int gz = (d + cz - 1) / cz, gy = (h + cy - 1) / cy, gx = (w + cx - 1) / cx;
  const size_t staged_bytes = (size_t)gz * gy * gx * cz * cy * cx * celem;
  // 16-byte accesses when every row segment starts and ends on a 16-byte boundary in both layouts
  const bool vec = celem == elem && ((size_t)cx * elem) % 16 == 0 && ((size_t)w * elem) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  const size_t total = vec ? staged_bytes / 16 : staged_bytes;
  const size_t want = (total + 255) / 256;
  const int blocks = (int)(want < (size_t)148 * 8 ? want : (size_t)148 * 8);
  if (vec)
    chunk_layout_kernel<uint4><<<blocks, 256, 0, stream>>>(src, dst, d, h, w, elem, celem, cz, cy, cx, gy, gx, total,
                                                           to_chunks);
  else
    chunk_layout_kernel<uint8_t><<<blocks, 256, 0, stream>>>(src, dst, d, h, w, elem, celem, cz, cy, cx, gy, gx, total,
                                                             to_chunks);
  return cudaGetLastError();
}

// PER consecutive items of one output row (x, class) per thread, stored as one aligned pack.
template <typename T, int PER>
struct alignas(sizeof(T) * PER <= 16 ? sizeof(T) * PER : 16) ZoomPack {
  T v[PER];
};

template <typename T, int PER>
__global__ void __launch_bounds__(256) zoom_gather_kernel(const T* __restrict__ src, int sh, int sw, int sc,
                                                          T* __restrict__ dst, int dh, int dw, int dc, size_t total,
                                                          const int* __restrict__ tz, const int* __restrict__ ty,
                                                          const int* __restrict__ tx, const int* __restrict__ tc) {
  const int row = dw * dc / PER;  // packs per output row (the launcher picks PER so that it divides dw * dc)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e0 = (int)(i % row) * PER;
    size_t r = i / row;
    const int y = (int)(r % dh);
    const int z = (int)(r / dh);
    const int iz = __ldg(tz + z), iy = __ldg(ty + y);
    const T* line = src + ((size_t)(iz < 0 ? 0 : iz) * sh + (iy < 0 ? 0 : iy)) * sw * sc;
    ZoomPack<T, PER> p;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int e = e0 + k;
      const int ix = __ldg(tx + e / dc), ic = __ldg(tc + e % dc);
      p.v[k] = (iz | iy | ix | ic) >= 0 ? line[(size_t)ix * sc + ic] : T();
    }
    *reinterpret_cast<ZoomPack<T, PER>*>(dst + i * PER) = p;
  }
}

template <typename T>
static cudaError_t zoom_launch(const uint8_t* src, const int* sd, uint8_t* dst, const int* dd, const int* tz,
                               const int* ty, const int* tx, const int* tc, cudaStream_t stream) {
  const size_t items = (size_t)dd[0] * dd[1] * dd[2] * dd[3];
  if (items == 0) return cudaSuccess;
  const bool packed = ((size_t)dd[2] * dd[3]) % 4 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  const size_t total = packed ? items / 4 : items;
  const size_t want = (total + 255) / 256;
  const int blocks = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);
  if (packed)
    zoom_gather_kernel<T, 4><<<blocks, 256, 0, stream>>>(reinterpret_cast<const T*>(src), sd[1], sd[2], sd[3],
                                                         reinterpret_cast<T*>(dst), dd[1], dd[2], dd[3], total, tz, ty, tx,
                                                         tc);
  else
    zoom_gather_kernel<T, 1><<<blocks, 256, 0, stream>>>(reinterpret_cast<const T*>(src), sd[1], sd[2], sd[3],
                                                         reinterpret_cast<T*>(dst), dd[1], dd[2], dd[3], total, tz, ty, tx,
                                                         tc);
  return cudaGetLastError();
}

cudaError_t launch_zoom_gather(const uint8_t* src, const int* sdims, uint8_t* dst, const int* ddims, const int* tz,
                               const int* ty, const int* tx, const int* tc, int item, cudaStream_t stream) {
  switch (item) {
    case 1: return zoom_launch<uint8_t>(src, sdims, dst, ddims, tz, ty, tx, tc, stream);
    case 2: return zoom_launch<uint16_t>(src, sdims, dst, ddims, tz, ty, tx, tc, stream);
    case 4: return zoom_launch<uint32_t>(src, sdims, dst, ddims, tz, ty, tx, tc, stream);
    case 8: return zoom_launch<uint64_t>(src, sdims, dst, ddims, tz, ty, tx, tc, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace iu
