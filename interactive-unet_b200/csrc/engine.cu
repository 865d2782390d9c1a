// Host side of the engine behind include/iunet_b200.h: weight folding / packing, the layer program of
// smp.Unet('resnet34'), workspace + TMA descriptor planning, and the C ABI.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "conv_tc.cuh"
#include "iunet_b200.h"

using namespace iu;

namespace {

thread_local std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TensorSpec {
  int c;
  int hdiv;  // spatial size = input size / hdiv
};

struct ConvLayer {
  std::string name;
  int nseg = 1;
  ConvSegment seg[2];
  int src[2] = {-1, -1};
  int out = -1;       // tensor id (bf16 modes); -1 for the head
  int residual = -1;  // tensor id or -1
  int relu = 1, up2x = 0, mode = kEpiBf16;
  int cout = 0, cout_pad = 0, ktot = 0, kc = 0, bn = 0;
  int out_hdiv = 1;  // output geometry BEFORE the optional 2x upsample
  __nv_bfloat16* d_w = nullptr;
  float* d_b = nullptr;
  // row-folded packing (conv_row.cu) for the stride-1 3x3 layers with cout_pad <= 64; nullptr otherwise
  __nv_bfloat16* d_wf = nullptr;
  int ktot_f = 0;
  // four-slot packing of an upsampled first segment (conv_row.cu): [W2 | W1+W2 | W0+W1 | W0] per filter column
  __nv_bfloat16* d_wu = nullptr;
};

struct Plan {
  int batch = 0, batch_pad = 0, h = 0, w = 0;
  std::vector<__nv_bfloat16*> bufs;
  float* x_in = nullptr;
  std::vector<ConvArgs> args;
  ConvArgs stem_epi;
  CUtensorMap stem_omap;
  ChainArgs chain;         // fused decoder block 4 conv1 -> conv2 -> head (conv_chain.cu), valid when use_chain
  bool use_chain = false;
  size_t bytes = 0;
  uint64_t last_use = 0;  // LRU stamp while the plan sits in the cache
  // latency path (`predict_slice`, predict.py:16-47): the 46 launches of one single-batch forward captured as a CUDA
  // graph.  The graph reads plan.x_in and writes plan.fwd_out, so its kernel arguments never change.
  float* fwd_out = nullptr;  // [batch][C][h][w] fp32
  cudaGraphExec_t graph = nullptr;
  int fwd_calls = 0;
  int graph_launches = 0;  // kernels inside the graph
};

struct Scratch {
  void* ptr;
  size_t bytes;
  bool used;
};

struct ProfSpan {
  cudaEvent_t begin, end;
  int cls;
};

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
uint16_t f2h(float f) {
  f = std::max(-65504.0f, std::min(65504.0f, f));
  return __half_as_ushort(__float2half_rn(f));
}
uint16_t to16(float f, int fp16) { return fp16 ? f2h(f) : f2bf(f); }

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

bool pick_tile(int cin_min, int cout_pad, int* kc, int* bn) {
  *kc = cin_min >= 64 ? 64 : cin_min;
  *bn = cout_pad >= 128 ? 128 : cout_pad;
  return conv_tc_smem_bytes(*kc, *bn) > 0;
}

}  // namespace

struct iu_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // result copies that overlap the last axis (predict_volume), created on first use
  std::string err;
  EncodeTiledFn encode = nullptr;
  int num_classes = 0;
  bool loaded = false;
  int max_batch = 0;
  int fp16 = 1;  // 16-bit storage format of weights / activations: 1 = IEEE fp16 (default), 0 = bf16
  int auto_batch_override = 0;  // env IU_AUTO_BATCH: slices per internal batch instead of the automatic choice
  int conv_pair = 0;  // env IU_CONV_PAIR=1 routes the wide layers to the CTA-pair kernel (opt-in: not yet faster end to end)
  int conv_variant = 0;  // 0 = auto (halo kernel where applicable), 1 = per-tap TMA kernel only (env IU_CONV_VARIANT)
  int conv_row = 1;  // env IU_CONV_ROW=0 keeps the narrow layers off the row-folded kernel
  int conv_bn256 = 1;  // env IU_CONV_BN256=0: per-tap kernel with 128-wide Cout tiles only
  int conv_cluster = 0;  // env IU_CONV_CLUSTER=1: per-tap kernel in CTA pairs that multicast the weight boxes (single-Cout-tile layers)
  int conv_bm = 1;     // env IU_CONV_BM2: bit 0 = two-M-tile CTA tiles for BN 128, bit 1 = for BN 256 (per-tap kernel)
  __nv_bfloat16* d_ident = nullptr;  // [64][64] identity in the storage format: the residual segment of conv_row.cu
  unsigned long long* d_debug = nullptr;  // env IU_CONV_DEBUG=1: 16 cycle counters per conv layer (development aid)
  int64_t launches = 0;
  size_t weight_bytes = 0;

  __nv_bfloat16* d_stem_w = nullptr;  // [64][64] 16-bit, k = filter_row * 8 + filter_col (conv_stem.cu)
  float* d_stem_b = nullptr;
  CUtensorMap stem_bmap;
  std::vector<TensorSpec> tensors;
  std::vector<ConvLayer> convs;
  int t_f1 = -1, t_p1 = -1;
  float* d_window = nullptr;        // device copy of the last 1-D window factor given to iu_engine_reduce(_planes)
  std::vector<float> window_host;   // ... and its host values: re-uploaded only when they change
  Plan plan;                 // the current activation plan
  std::vector<Plan> cached;  // other (batch, h, w) plans kept allocated: the app alternates predict_slice / predict_volumes
  int plan_cache = 3;        // env IU_PLAN_CACHE: plans kept besides none in use (0 = re-plan on every shape change)
  uint64_t use_clock = 0;
  int use_graph = 1;         // env IU_GRAPH=0: never capture the single-batch forward
  int conv_chain = 1;        // env IU_CONV_CHAIN=0: decoder block 4 + head as three separate row-folded launches
  int halo_tma = 0;          // env IU_HALO_TMA=1: the Cout >= 128 stride-1 layers run on the halo kernel with TMA-filled
                             // tiles instead of the per-tap kernel (2: only the Cout-128 layers; 24: halo rows padded to 24 pixels).  Measured neutral:
                             // 60 % less L2 -> SM traffic and 6 % fewer cycles on the Cout-128 layers, 20 % more cycles on
                             // the Cout >= 256 ones (N = 128 instead of N = 256 MMAs), same step time under the power cap
  int row_tma = 1;           // env IU_ROW_TMA=0: the Cout-64 row layers with one identity source gather with cp.async again
  int row_res_tma = 1;       // env IU_ROW_RES_TMA=0: the Cout-64 row layers add their shortcut as an identity K segment
  int stem_pool = 0;         // env IU_STEM_POOL=1: max-pool inside the stem's epilogue (measured slower: 9.7 ms against
                             // 4.1 + 4.3 ms per 512^3 -- 40 % of a tile's pooled pixels go through red.max)
  int conv_small_bn = 1;     // env IU_CONV_SMALL_BN=0: keep the wide Cout tiles even when they leave most SMs idle
  int num_sms = 148;
  int conv_pair2 = 0;        // env IU_CONV_PAIR2: per-tap kernel on CTA pairs; bit 0: Cout >= 256 layers, bit 1: Cout 128
  std::vector<Scratch> scratch;
  size_t scratch_keep = ~(size_t)0;  // env IU_SCRATCH_KEEP_MB: idle scratch above this is returned to the driver when a
                                     // volume call returns (default: keep everything for the next call of the same
                                     // size; `iu_engine_release_workspace` is the explicit hand-back)

  // optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline numbers)
  bool prof = false;
  std::vector<ProfSpan> spans;
  size_t spans_used = 0;
  double prof_ms[IU_PROF_CLASSES] = {0, 0, 0, 0, 0};
  int64_t prof_n[IU_PROF_CLASSES] = {0, 0, 0, 0, 0};

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();
      return fail(IU_ERR_OOM, std::string("CUDA out of memory (") + what + ")");
    }
    return fail(IU_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
};

#define IU_CUDA(e_, call_)                                   \
  do {                                                       \
    cudaError_t _err = (call_);                              \
    if (_err != cudaSuccess) return (e_)->cuda_fail(_err, #call_); \
  } while (0)

namespace {

// ------------------------------------------------------------------ profiling spans
void prof_flush(iu_engine* e) {
  if (e->spans_used == 0) return;
  cudaStreamSynchronize(e->stream);
  for (size_t i = 0; i < e->spans_used; ++i) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, e->spans[i].begin, e->spans[i].end) == cudaSuccess) {
      e->prof_ms[e->spans[i].cls] += ms;
      e->prof_n[e->spans[i].cls] += 1;
    }
  }
  cudaGetLastError();
  e->spans_used = 0;
}
void prof_begin(iu_engine* e, int cls) {
  if (!e->prof) return;
  if (e->spans_used == e->spans.size()) {
    if (e->spans.size() >= 16384) {
      prof_flush(e);
    } else {
      ProfSpan s;
      cudaEventCreate(&s.begin);
      cudaEventCreate(&s.end);
      s.cls = cls;
      e->spans.push_back(s);
    }
  }
  e->spans[e->spans_used].cls = cls;
  cudaEventRecord(e->spans[e->spans_used].begin, e->stream);
}
void prof_end(iu_engine* e) {
  if (!e->prof) return;
  cudaEventRecord(e->spans[e->spans_used].end, e->stream);
  e->spans_used += 1;
}

// ------------------------------------------------------------------ scratch pool
int scratch_get(iu_engine* e, size_t bytes, void** out) {
  int best = -1;
  for (size_t i = 0; i < e->scratch.size(); ++i)
    if (!e->scratch[i].used && e->scratch[i].bytes >= bytes &&
        (best < 0 || e->scratch[i].bytes < e->scratch[best].bytes))
      best = (int)i;
  if (best >= 0) {
    e->scratch[best].used = true;
    *out = e->scratch[best].ptr;
    return IU_OK;
  }
  void* p = nullptr;
  cudaError_t ce = cudaMalloc(&p, bytes ? bytes : 16);
  if (ce != cudaSuccess) {
    // drop idle cached blocks and retry once
    for (auto& s : e->scratch)
      if (!s.used && s.ptr) {
        cudaFree(s.ptr);
        s.ptr = nullptr;
        s.bytes = 0;
      }
    cudaGetLastError();
    ce = cudaMalloc(&p, bytes ? bytes : 16);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaMalloc(scratch)");
  }
  e->scratch.push_back({p, bytes, true});
  *out = p;
  return IU_OK;
}
void scratch_put(iu_engine* e, void* p) {
  if (!p) return;
  for (auto& s : e->scratch)
    if (s.ptr == p) s.used = false;
}
// Idle scratch beyond `scratch_keep` bytes goes back to the driver (largest blocks first): after a tiled volume the
// fp32 accumulators would otherwise stay allocated for the engine's lifetime, outside torch's caching allocator, and
// the trainer / suggestor in the same process would see that much less free HBM.
void scratch_trim(iu_engine* e) {
  size_t idle = 0;
  for (auto& s : e->scratch)
    if (!s.used && s.ptr) idle += s.bytes;
  while (idle > e->scratch_keep) {
    int big = -1;
    for (size_t i = 0; i < e->scratch.size(); ++i)
      if (!e->scratch[i].used && e->scratch[i].ptr && (big < 0 || e->scratch[i].bytes > e->scratch[big].bytes)) big = (int)i;
    if (big < 0) break;
    cudaFree(e->scratch[big].ptr);
    idle -= e->scratch[big].bytes;
    e->scratch.erase(e->scratch.begin() + big);
  }
}
// Every ABI call runs on the engine's device and puts the caller's current device back on return (a multi-GPU
// process -- torch -- must not find its current device changed by a prediction call).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t enter(int device) {
    cudaError_t ce = cudaGetDevice(&prev);
    if (ce != cudaSuccess) return ce;
    if (prev != device) {
      ce = cudaSetDevice(device);
      switched = ce == cudaSuccess;
    }
    return ce;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------------ weights
struct HostTensors {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  const float* get(const std::string& k, int64_t numel, std::string* err) const {
    auto it = t.find(k);
    if (it == t.end()) {
      *err = "missing tensor '" + k + "'";
      return nullptr;
    }
    if (it->second.second != numel) {
      *err = "tensor '" + k + "' has " + std::to_string(it->second.second) + " elements, expected " +
             std::to_string(numel);
      return nullptr;
    }
    return it->second.first;
  }
};

// Eval-mode BatchNorm folded into the preceding bias-free conv:
//   y = (conv(x) - mean) * gamma / sqrt(var + eps) + beta  =  conv_{w * s}(x) + (beta - mean * s)
bool fold_bn(const HostTensors& ht, const std::string& conv_w, const std::string& bn, int cout, int64_t per_out,
             std::vector<float>* w, std::vector<float>* b, std::string* err) {
  const float* cw = ht.get(conv_w, (int64_t)cout * per_out, err);
  const float* g = ht.get(bn + ".weight", cout, err);
  const float* be = ht.get(bn + ".bias", cout, err);
  const float* mu = ht.get(bn + ".running_mean", cout, err);
  const float* var = ht.get(bn + ".running_var", cout, err);
  if (!cw || !g || !be || !mu || !var) return false;
  w->resize((size_t)cout * per_out);
  b->resize(cout);
  for (int co = 0; co < cout; ++co) {
    const double s = (double)g[co] / std::sqrt((double)var[co] + 1e-5);
    for (int64_t i = 0; i < per_out; ++i) (*w)[co * per_out + i] = (float)((double)cw[co * per_out + i] * s);
    (*b)[co] = (float)((double)be[co] - (double)mu[co] * s);
  }
  return true;
}

// K order of the implicit GEMM: segment -> tap row -> tap col -> channel (see conv_tc.cu producer loop).
void pack_segment(std::vector<uint16_t>& dst, int fp16, int ktot, int kbase, const float* w, int cout, int cin_total,
                  int coff, int cin_s, int ks) {
  for (int co = 0; co < cout; ++co)
    for (int r = 0; r < ks; ++r)
      for (int q = 0; q < ks; ++q)
        for (int c = 0; c < cin_s; ++c)
          dst[(size_t)co * ktot + kbase + (r * ks + q) * cin_s + c] =
              to16(w[(((size_t)co * cin_total + coff + c) * ks + r) * ks + q], fp16);
}

// Row-folded packing (conv_row.cu): rows = (2 - ky) * cout_pad + co, K = segment -> kx -> channel, so that the
// [3*cout_pad x KC] tile of one (chunk, kx) holds the three vertical taps side by side in N.
void pack_fold_segment(std::vector<uint16_t>& dst, int fp16, int ktot_f, int kbase, const float* w, int cout,
                       int cout_pad, int cin_total, int coff, int cin_s) {
  for (int co = 0; co < cout; ++co)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx)
        for (int c = 0; c < cin_s; ++c)
          dst[(size_t)((2 - ky) * cout_pad + co) * ktot_f + kbase + kx * cin_s + c] =
              to16(w[(((size_t)co * cin_total + coff + c) * 3 + ky) * 3 + kx], fp16);
}

// Upsampled segment (channels [0, cin0) of the conv): source row s feeds output rows 2s-1 .. 2s+2 through the
// vertically pre-summed filters; rows = slot * cout_pad + co, K = kx -> channel.
void pack_up_fold(std::vector<uint16_t>& dst, int fp16, const float* w, int cout, int cout_pad, int cin_total, int cin0) {
  const int k3 = 3 * cin0;
  for (int co = 0; co < cout; ++co)
    for (int kx = 0; kx < 3; ++kx)
      for (int c = 0; c < cin0; ++c) {
        const float* f = w + (((size_t)co * cin_total + c) * 3) * 3 + kx;  // f[ky * 3]
        const float w0 = f[0], w1 = f[3], w2 = f[6];
        const float slot[4] = {w2, w1 + w2, w0 + w1, w0};
        for (int t = 0; t < 4; ++t) dst[(size_t)(t * cout_pad + co) * k3 + kx * cin0 + c] = to16(slot[t], fp16);
      }
}

int upload_fold(iu_engine* e, ConvLayer& L, const float* w, int cout, int cin_total, int nsrc, const int* src_cin) {
  L.ktot_f = 3 * cin_total;
  std::vector<uint16_t> packed((size_t)3 * L.cout_pad * L.ktot_f, 0);
  int kbase = 0, coff = 0;
  for (int s = 0; s < nsrc; ++s) {
    pack_fold_segment(packed, e->fp16, L.ktot_f, kbase, w, cout, L.cout_pad, cin_total, coff, src_cin[s]);
    kbase += 3 * src_cin[s];
    coff += src_cin[s];
  }
  IU_CUDA(e, cudaMalloc(&L.d_wf, packed.size() * 2));
  IU_CUDA(e, cudaMemcpy(L.d_wf, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  e->weight_bytes += packed.size() * 2;
  if (L.seg[0].up) {
    std::vector<uint16_t> pu((size_t)4 * L.cout_pad * 3 * src_cin[0], 0);
    pack_up_fold(pu, e->fp16, w, cout, L.cout_pad, cin_total, src_cin[0]);
    IU_CUDA(e, cudaMalloc(&L.d_wu, pu.size() * 2));
    IU_CUDA(e, cudaMemcpy(L.d_wu, pu.data(), pu.size() * 2, cudaMemcpyHostToDevice));
    e->weight_bytes += pu.size() * 2;
  }
  return IU_OK;
}

int ensure_identity(iu_engine* e) {
  if (e->d_ident) return IU_OK;
  std::vector<uint16_t> id(64 * 64, 0);
  for (int i = 0; i < 64; ++i) id[i * 64 + i] = to16(1.0f, e->fp16);
  IU_CUDA(e, cudaMalloc(&e->d_ident, id.size() * 2));
  IU_CUDA(e, cudaMemcpy(e->d_ident, id.data(), id.size() * 2, cudaMemcpyHostToDevice));
  return IU_OK;
}

int upload_conv(iu_engine* e, ConvLayer& L, const std::vector<uint16_t>& packed, const std::vector<float>& bias) {
  IU_CUDA(e, cudaMalloc(&L.d_w, packed.size() * 2));
  IU_CUDA(e, cudaMalloc(&L.d_b, (size_t)L.cout_pad * 4));
  std::vector<float> bp(L.cout_pad, 0.0f);
  std::copy(bias.begin(), bias.end(), bp.begin());
  IU_CUDA(e, cudaMemcpy(L.d_w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  IU_CUDA(e, cudaMemcpy(L.d_b, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
  e->weight_bytes += packed.size() * 2 + bp.size() * 4;
  return IU_OK;
}

void free_weights(iu_engine* e) {
  for (auto& L : e->convs) {
    if (L.d_w) cudaFree(L.d_w);
    if (L.d_b) cudaFree(L.d_b);
    if (L.d_wf) cudaFree(L.d_wf);
    if (L.d_wu) cudaFree(L.d_wu);
  }
  if (e->d_ident) cudaFree(e->d_ident);
  e->d_ident = nullptr;
  e->convs.clear();
  e->tensors.clear();
  if (e->d_stem_w) cudaFree(e->d_stem_w);
  if (e->d_stem_b) cudaFree(e->d_stem_b);
  e->d_stem_w = nullptr;
  e->d_stem_b = nullptr;
  e->loaded = false;
  e->weight_bytes = 0;
}

void release_plan(Plan& p) {
  for (auto b : p.bufs)
    if (b) cudaFree(b);
  if (p.x_in) cudaFree(p.x_in);
  if (p.fwd_out) cudaFree(p.fwd_out);
  if (p.graph) cudaGraphExecDestroy(p.graph);
  p = Plan();
}
void free_plan(iu_engine* e) { release_plan(e->plan); }
void free_cached_plans(iu_engine* e) {
  for (auto& p : e->cached) release_plan(p);
  e->cached.clear();
}
void free_all_plans(iu_engine* e) {
  free_plan(e);
  free_cached_plans(e);
}

int new_tensor(iu_engine* e, int c, int hdiv) {
  e->tensors.push_back({c, hdiv});
  return (int)e->tensors.size() - 1;
}

int encode_weight_map(iu_engine* e, CUtensorMap* m, const void* base, int ktot, int cout_pad, int kc, int bn);

// Build one folded conv layer (+ optional fused 1x1/s2 downsample as a second K segment).
int add_conv(iu_engine* e, const HostTensors& ht, const std::string& name, const std::string& conv_key,
             const std::string& bn_key, int nsrc, const int* src, const int* src_cin, int ksize, int stride,
             int cout, int out_tensor, int out_hdiv, int residual, int relu, int up2x, const std::string& ds_conv,
             const std::string& ds_bn, int ds_src, int ds_cin, int src_up_mask = 0) {
  ConvLayer L;
  L.name = name;
  L.cout = cout;
  L.cout_pad = (cout + 15) / 16 * 16;
  L.out = out_tensor;
  L.out_hdiv = out_hdiv;
  L.residual = residual;
  L.relu = relu;
  L.up2x = up2x;
  int cin_total = 0, cin_min = 1 << 30;
  for (int s = 0; s < nsrc; ++s) {
    cin_total += src_cin[s];
    cin_min = std::min(cin_min, src_cin[s]);
  }
  std::vector<float> w, b;
  std::string err;
  if (!fold_bn(ht, conv_key, bn_key, cout, (int64_t)cin_total * ksize * ksize, &w, &b, &err))
    return e->fail(IU_ERR_INVALID, err);
  L.nseg = nsrc;
  int k = 0;
  for (int s = 0; s < nsrc; ++s) {
    L.seg[s] = {src_cin[s], ksize, stride, ksize / 2, (src_up_mask >> s) & 1};
    L.src[s] = src[s];
    k += ksize * ksize * src_cin[s];
  }
  std::vector<float> wd, bd;
  if (!ds_conv.empty()) {
    if (nsrc != 1) return e->fail(IU_ERR_INVALID, "downsample fusion needs a single main source");
    if (!fold_bn(ht, ds_conv, ds_bn, cout, ds_cin, &wd, &bd, &err)) return e->fail(IU_ERR_INVALID, err);
    L.nseg = 2;
    L.seg[1] = {ds_cin, 1, 2, 0, 0};
    L.src[1] = ds_src;
    k += ds_cin;
    cin_min = std::min(cin_min, ds_cin);
    for (int co = 0; co < cout; ++co) b[co] += bd[co];
  }
  L.ktot = k;
  if (!pick_tile(cin_min, L.cout_pad, &L.kc, &L.bn))
    return e->fail(IU_ERR_INVALID, "no kernel variant for layer " + name);
  for (int s = 0; s < L.nseg; ++s)
    if (L.seg[s].cin % L.kc) return e->fail(IU_ERR_INVALID, "channel count not a multiple of the K chunk: " + name);
  std::vector<uint16_t> packed((size_t)L.cout_pad * L.ktot, 0);
  int kbase = 0, coff = 0;
  for (int s = 0; s < nsrc; ++s) {
    pack_segment(packed, e->fp16, L.ktot, kbase, w.data(), cout, cin_total, coff, src_cin[s], ksize);
    kbase += ksize * ksize * src_cin[s];
    coff += src_cin[s];
  }
  if (!ds_conv.empty()) pack_segment(packed, e->fp16, L.ktot, kbase, wd.data(), cout, ds_cin, 0, ds_cin, 1);
  int rc = upload_conv(e, L, packed, b);
  if (rc != IU_OK) return rc;
  if (ds_conv.empty() && ksize == 3 && stride == 1 && !up2x && L.cout_pad <= 64 && L.cout_pad == cout) {
    rc = upload_fold(e, L, w.data(), cout, cin_total, nsrc, src_cin);
    if (rc != IU_OK) return rc;
  }
  e->convs.push_back(L);
  return IU_OK;
}

int build_network(iu_engine* e, const HostTensors& ht, int num_classes) {
  std::string err;
  {
    int rc = ensure_identity(e);
    if (rc != IU_OK) return rc;
  }
  // ---- stem: encoder.conv1 + bn1 (+ReLU) as a K=64 GEMM on the tensor cores (conv_stem.cu)
  {
    std::vector<float> w, b;
    if (!fold_bn(ht, "encoder.conv1.weight", "encoder.bn1", 64, 49, &w, &b, &err)) return e->fail(IU_ERR_INVALID, err);
    std::vector<uint16_t> wk(64 * 64, 0);
    for (int co = 0; co < 64; ++co)
      for (int r = 0; r < 7; ++r)
        for (int q = 0; q < 7; ++q) wk[co * 64 + r * 8 + q] = to16(w[co * 49 + r * 7 + q], e->fp16);
    IU_CUDA(e, cudaMalloc(&e->d_stem_w, wk.size() * 2));
    IU_CUDA(e, cudaMalloc(&e->d_stem_b, 64 * 4));
    IU_CUDA(e, cudaMemcpy(e->d_stem_w, wk.data(), wk.size() * 2, cudaMemcpyHostToDevice));
    IU_CUDA(e, cudaMemcpy(e->d_stem_b, b.data(), 64 * 4, cudaMemcpyHostToDevice));
    e->weight_bytes += wk.size() * 2 + 256;
    int rc = encode_weight_map(e, &e->stem_bmap, e->d_stem_w, 64, 64, 64, 64);
    if (rc != IU_OK) return rc;
  }
  e->t_f1 = new_tensor(e, 64, 2);
  e->t_p1 = new_tensor(e, 64, 4);

  // ---- encoder: torchvision BasicBlock ResNets -- resnet34 [3,4,6,3] (the reference's accelerated configuration)
  //      and resnet18 [2,2,2,2]; the block counts are read off the tensor names
  int nblocks[4] = {0, 0, 0, 0};
  for (int li = 0; li < 4; ++li) {
    const std::string layer = "encoder.layer" + std::to_string(li + 1) + ".";
    while (ht.t.count(layer + std::to_string(nblocks[li]) + ".conv1.weight")) ++nblocks[li];
    if (nblocks[li] < 1) return e->fail(IU_ERR_INVALID, "missing tensor '" + layer + "0.conv1.weight'");
    if (ht.t.count(layer + "0.conv3.weight"))
      return e->fail(IU_ERR_INVALID, "bottleneck encoders (resnet50 and up) are not supported: BasicBlock ResNets only");
  }
  const int chans[4] = {64, 128, 256, 512};
  int cur = e->t_p1, cur_c = 64, cur_hdiv = 4;
  int feat[6] = {-1, e->t_f1, -1, -1, -1, -1};
  for (int li = 0; li < 4; ++li) {
    const int cout = chans[li];
    for (int b = 0; b < nblocks[li]; ++b) {
      const std::string base = "encoder.layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const bool down = (li > 0 && b == 0);
      const int stride = down ? 2 : 1;
      const int hdiv = cur_hdiv * stride;
      const int t = new_tensor(e, cout, hdiv);
      int rc = add_conv(e, ht, base + ".conv1", base + ".conv1.weight", base + ".bn1", 1, &cur, &cur_c, 3, stride,
                        cout, t, hdiv, -1, 1, 0, "", "", -1, 0);
      if (rc != IU_OK) return rc;
      const int o = new_tensor(e, cout, hdiv);
      if (down)
        rc = add_conv(e, ht, base + ".conv2+downsample", base + ".conv2.weight", base + ".bn2", 1, &t, &cout, 3, 1,
                      cout, o, hdiv, -1, 1, 0, base + ".downsample.0.weight", base + ".downsample.1", cur, cur_c);
      else
        rc = add_conv(e, ht, base + ".conv2", base + ".conv2.weight", base + ".bn2", 1, &t, &cout, 3, 1, cout, o,
                      hdiv, cur, 1, 0, "", "", -1, 0);
      if (rc != IU_OK) return rc;
      cur = o;
      cur_c = cout;
      cur_hdiv = hdiv;
    }
    feat[li + 2] = cur;
  }
  // ---- decoder (smp UnetDecoder): x = cat([up2x(x), skip]) -> conv1 -> conv2; x stays at its own resolution in
  //      memory and conv1's operand gather reads it through the 2x nearest upsample (ConvSegment::up).
  const int dec_out[5] = {256, 128, 64, 32, 16};
  const int skip_t[5] = {feat[4], feat[3], feat[2], feat[1], -1};
  const int skip_c[5] = {256, 128, 64, 64, 0};
  int x = cur, x_c = 512;
  for (int i = 0; i < 5; ++i) {
    const int hdiv = 16 >> i;
    const std::string base = "decoder.blocks." + std::to_string(i);
    const int t = new_tensor(e, dec_out[i], hdiv);
    int srcs[2] = {x, skip_t[i]};
    int cins[2] = {x_c, skip_c[i]};
    int rc = add_conv(e, ht, base + ".conv1", base + ".conv1.0.weight", base + ".conv1.1", skip_t[i] >= 0 ? 2 : 1,
                      srcs, cins, 3, 1, dec_out[i], t, hdiv, -1, 1, 0, "", "", -1, 0, /*src_up_mask=*/1);
    if (rc != IU_OK) return rc;
    const int o = new_tensor(e, dec_out[i], hdiv);
    rc = add_conv(e, ht, base + ".conv2", base + ".conv2.0.weight", base + ".conv2.1", 1, &t, &dec_out[i], 3, 1,
                  dec_out[i], o, hdiv, -1, 1, 0, "", "", -1, 0);
    if (rc != IU_OK) return rc;
    x = o;
    x_c = dec_out[i];
  }
  // ---- segmentation head: Conv2d(16, C, 3, padding=1) with bias, then the reference's Softmax(dim=1)
  {
    ConvLayer L;
    L.name = "segmentation_head.0";
    L.cout = num_classes;
    L.cout_pad = 16;
    L.nseg = 1;
    L.seg[0] = {16, 3, 1, 1, 0};
    L.src[0] = x;
    L.out = -1;
    L.out_hdiv = 1;
    L.relu = 0;
    L.mode = kEpiSoftmaxNHWC;
    L.ktot = 9 * 16;
    L.kc = 16;
    L.bn = 16;
    const float* w = ht.get("segmentation_head.0.weight", (int64_t)num_classes * 16 * 9, &err);
    const float* b = ht.get("segmentation_head.0.bias", num_classes, &err);
    if (!w || !b) return e->fail(IU_ERR_INVALID, err);
    std::vector<uint16_t> packed((size_t)16 * L.ktot, 0);
    pack_segment(packed, e->fp16, L.ktot, 0, w, num_classes, 16, 0, 16, 3);
    int rc = upload_conv(e, L, packed, std::vector<float>(b, b + num_classes));
    if (rc != IU_OK) return rc;
    const int head_cin = 16;
    rc = upload_fold(e, L, w, num_classes, 16, 1, &head_cin);
    if (rc != IU_OK) return rc;
    e->convs.push_back(L);
  }
  return IU_OK;
}

// ------------------------------------------------------------------ planning
int encode_act_map(iu_engine* e, CUtensorMap* m, const void* base, int c, int w, int h, int n, int kc, int boxw,
                   int boxh, int boxn, int estride) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)boxw, (cuuint32_t)boxh, (cuuint32_t)boxn};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  const CUtensorMapSwizzle swz =
      kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = e->encode(m, e->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return e->fail(IU_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed with CUresult " + std::to_string((int)r));
  return IU_OK;
}
int encode_weight_map(iu_engine* e, CUtensorMap* m, const void* base, int ktot, int cout_pad, int kc, int bn) {
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle swz =
      kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = e->encode(m, e->fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return e->fail(IU_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed with CUresult " + std::to_string((int)r));
  return IU_OK;
}

// Fill the geometry / tile fields of a ConvArgs for an output of out_h x out_w.
// The 128-pixel tile is nb images x th rows x tw columns (powers of two).  Among tw = 16, 8, 4 pick the shape that
// wastes the fewest pixels on this image size (20x20: 16x8 tiles cover 52 %, 4x32 tiles 62 %); ties go to the wider tile.
void set_tiling(ConvArgs* a, int batch, int out_h, int out_w) {
  a->batch = batch;
  a->out_h = out_h;
  a->out_w = out_w;
  long best_cover = -1;
  for (int tw0 = 16; tw0 >= 4; tw0 >>= 1) {
    const int tw = std::min(tw0, pow2_ceil(out_w));
    const int th = std::min(kTileM / tw, pow2_ceil(out_h));
    const int tx = (out_w + tw - 1) / tw, ty = (out_h + th - 1) / th;
    const long cover = (long)tx * tw * ty * th;  // pixels computed per image (>= out_h * out_w)
    if (best_cover < 0 || cover < best_cover) {
      best_cover = cover;
      a->tw = tw;
      a->th = th;
      a->nb = kTileM / (tw * th);
      a->tiles_x = tx;
      a->tiles_y = ty;
    }
  }
}

size_t plan_bytes(const iu_engine* e, int batch_pad, int h, int w) {
  size_t total = (size_t)batch_pad * h * w * 4;
  for (const auto& t : e->tensors) total += (size_t)batch_pad * (h / t.hdiv) * (w / t.hdiv) * t.c * 2;
  return total;
}

// Cout tile width for layers whose pixel count gives only a few CTA tiles (the latency path): 0 = keep the default.
int plan_small_bn(const iu_engine* e, const ConvArgs& a, int kc, int cout_pad) {
  if (kc != 64 || a.mode != kEpiBf16 || cout_pad % 128) return 0;
  const long pixels = (long)a.batch * a.out_h * a.out_w;
  const int mtiles = (int)((pixels + kTileM - 1) / kTileM);
  if (mtiles * (cout_pad / 64) * 2 <= e->num_sms) return 32;   // even 64-wide tiles would fill under half the SMs
  if (mtiles * (cout_pad / 128) * 4 <= e->num_sms) return 64;
  return 0;
}

// Allocate and describe the workspace of (batch, h, w) in e->plan (which must be empty).
int build_plan(iu_engine* e, int batch, int h, int w) {
  Plan& p = e->plan;
  const int bp = (batch + 7) / 8 * 8;  // images are tiled in groups of up to 8: keep every TMA box inside the tensor
  p.bufs.assign(e->tensors.size(), nullptr);
  for (size_t i = 0; i < e->tensors.size(); ++i) {
    const TensorSpec& t = e->tensors[i];
    const size_t bytes = (size_t)bp * (h / t.hdiv) * (w / t.hdiv) * t.c * 2;
    cudaError_t ce = cudaMalloc(&p.bufs[i], bytes);
    if (ce != cudaSuccess) {
      free_plan(e);
      return e->cuda_fail(ce, "cudaMalloc(activations)");
    }
    cudaMemsetAsync(p.bufs[i], 0, bytes, e->stream);
    p.bytes += bytes;
  }
  {
    cudaError_t ce = cudaMalloc(&p.x_in, (size_t)bp * h * w * 4);
    if (ce != cudaSuccess) {
      free_plan(e);
      return e->cuda_fail(ce, "cudaMalloc(input slices)");
    }
    p.bytes += (size_t)bp * h * w * 4;
  }
  p.args.resize(e->convs.size());
  for (size_t i = 0; i < e->convs.size(); ++i) {
    const ConvLayer& L = e->convs[i];
    ConvArgs& a = p.args[i];
    memset(&a, 0, sizeof(a));
    set_tiling(&a, batch, h / L.out_hdiv, w / L.out_hdiv);
    a.nseg = L.nseg;
    for (int s = 0; s < L.nseg; ++s) {
      a.seg[s] = L.seg[s];
      a.src_ptr[s] = p.bufs[L.src[s]];
      const TensorSpec& ts = e->tensors[L.src[s]];
      const int st = L.seg[s].stride;
      if (L.seg[s].up) continue;  // read by the halo kernel's gather only (raw pointer)
      int rc = encode_act_map(e, &a.amap[s], p.bufs[L.src[s]], ts.c, w / ts.hdiv, h / ts.hdiv, bp, L.kc, a.tw * st,
                              a.th * st, a.nb, st);
      if (rc != IU_OK) {
        free_plan(e);
        return rc;
      }
    }
    int rc = encode_weight_map(e, &a.bmap, L.d_w, L.ktot, L.cout_pad, L.kc, L.bn);
    if (rc == IU_OK && L.kc == 64 && L.cout_pad % 128 == 0)
      rc = encode_weight_map(e, &a.bmap2, L.d_w, L.ktot, L.cout_pad, 64, 64);
    if (rc == IU_OK && L.kc == 64 && L.cout_pad % 256 == 0 && L.mode == kEpiBf16) {
      rc = encode_weight_map(e, &a.bmap256, L.d_w, L.ktot, L.cout_pad, 64, 256);
      a.use_bn256 = rc == IU_OK;
    }
    if (rc != IU_OK) {
      free_plan(e);
      return rc;
    }
    a.mode = L.mode;
    a.small_bn = plan_small_bn(e, a, L.kc, L.cout_pad);
    if (a.small_bn && encode_weight_map(e, &a.bmap_small, L.d_w, L.ktot, L.cout_pad, 64, a.small_bn) != IU_OK) a.small_bn = 0;
    a.cout = L.cout;
    a.bias = L.d_b;
    a.residual = L.residual >= 0 ? p.bufs[L.residual] : nullptr;
    a.out = L.out >= 0 ? p.bufs[L.out] : nullptr;
    a.relu = L.relu;
    a.fp16 = e->fp16;
    a.up2x = L.up2x;
    a.mode = L.mode;
    a.num_classes = e->num_classes;
    a.slice0 = 0;
    a.slice_count = batch;
    a.row_block = h;
    a.debug = (e->d_debug && i < 64) ? e->d_debug + 16 * i : nullptr;
    a.use_row = 0;
    a.halo_tma = 0;
    if (e->halo_tma && (e->halo_tma != 2 || L.cout_pad == 128) && L.kc == 64 && conv_halo_tma_applicable(a) &&
        a.out_h >= kHaloTile && a.out_w >= kHaloTile) {
      // halo tiles by TMA where the 16x16 blocks fill the GPU (a single small slice stays on the per-tap kernel)
      const long blocks = (long)((a.out_w + kHaloTile - 1) / kHaloTile) * ((a.out_h + kHaloTile - 1) / kHaloTile) * batch *
                          (L.cout_pad / 128);
      if (blocks >= e->num_sms) {
        for (int s = 0; s < L.nseg && rc == IU_OK; ++s) {
          const TensorSpec& ts = e->tensors[L.src[s]];
          rc = encode_act_map(e, &a.hmap[s], p.bufs[L.src[s]], ts.c, w / ts.hdiv, h / ts.hdiv, bp, 64,
                              e->halo_tma == 24 ? 24 : kHaloTile + 2, kHaloTile + 2, 1, 1);
        }
        if (rc != IU_OK) {
          free_plan(e);
          return rc;
        }
        a.halo_tma = e->halo_tma == 24 ? 24 : kHaloTile + 2;  // halo pixels per buffer row
      }
    }
    const int row_mode = (L.d_wf && e->conv_row) ? conv_row_mode(a) : 0;
    if (row_mode) {
      const int kcr = conv_row_kc(L.cout_pad, row_mode);
      rc = encode_weight_map(e, &a.bmapf, L.d_wf, L.ktot_f, 3 * L.cout_pad, kcr, 3 * L.cout_pad);
      if (rc == IU_OK && L.residual >= 0) rc = encode_weight_map(e, &a.bmapi, e->d_ident, 64, 64, kcr, L.cout_pad);
      if (rc == IU_OK && L.d_wu)
        rc = encode_weight_map(e, &a.bmapu, L.d_wu, 3 * L.seg[0].cin, 4 * L.cout_pad, kcr, 4 * L.cout_pad);
      if (rc == IU_OK && L.mode == kEpiBf16) {
        const TensorSpec& to = e->tensors[L.out];
        rc = encode_act_map(e, &a.omap, p.bufs[L.out], L.cout_pad, w / to.hdiv, h / to.hdiv, bp, L.cout_pad, 128,
                            conv_row_store_rows(L.cout_pad), 1, 1);
        // Cout-64 layers with a shortcut (layer1's conv2): the residual is added in the epilogue from a TMA-loaded
        // staging buffer (conv_row.cu) instead of travelling through the gather and the tensor core
        if (rc == IU_OK && L.residual >= 0 && L.cout_pad == 64 && row_mode == 1 && e->row_res_tma) {
          rc = encode_act_map(e, &a.rmap, p.bufs[L.residual], L.cout_pad, w / to.hdiv, h / to.hdiv, bp, L.cout_pad, 128,
                              conv_row_store_rows(L.cout_pad), 1, 1);
          a.res_tma = rc == IU_OK ? e->row_res_tma : 0;
        }
        // ... and those fed by ONE identity 64-channel tensor have their A ring filled by TMA (conv_row.cu, TMA_A)
        if (rc == IU_OK && e->row_tma && conv_row_tma_applicable(a)) {
          const TensorSpec& ts = e->tensors[L.src[0]];
          rc = encode_act_map(e, &a.rowmap, p.bufs[L.src[0]], ts.c, w / ts.hdiv, h / ts.hdiv, bp, 64, 130, 3, 1, 1);
          a.row_tma = rc == IU_OK ? e->row_tma : 0;
        }
      }
      if (rc == IU_OK) a.use_row = 1;
      else if (rc > 0) {
        free_plan(e);
        return rc;
      }
    }
  }
  {
    ConvArgs& a = p.stem_epi;
    memset(&a, 0, sizeof(a));
    a.batch = batch;
    a.out_h = h / 2;
    a.out_w = w / 2;
    a.cout = 64;
    a.bias = e->d_stem_b;
    a.out = p.bufs[e->t_f1];
    a.relu = 1;
    a.fp16 = e->fp16;
    a.mode = kEpiBf16;
    int rc = encode_act_map(e, &p.stem_omap, p.bufs[e->t_f1], 64, w / 2, h / 2, bp, 64, 8, 16, 1, 1);
    if (rc != IU_OK) {
      free_plan(e);
      return rc;
    }
  }
  // fused tail: the last three convs are decoder block 4 conv1 (upsampled 32 -> 16), conv2 (16 -> 16) and the head
  p.use_chain = false;
  const size_t nc = e->convs.size();
  // IU_CONV_CHAIN=2 (development): also where the strip overhead makes the fused tail the slower choice
  if (e->conv_chain && e->conv_row && e->conv_variant != 1 && nc >= 3 &&
      (conv_chain_applicable(h, w) || (e->conv_chain == 2 && h >= 16 && h % 8 == 0 && w >= 128))) {
    const ConvLayer &l1 = e->convs[nc - 3], &l2 = e->convs[nc - 2], &l3 = e->convs[nc - 1];
    const ConvArgs &a1 = p.args[nc - 3], &a2 = p.args[nc - 2], &a3 = p.args[nc - 1];
    const bool shapes = l1.nseg == 1 && l1.seg[0].up && l1.seg[0].cin == 32 && l1.cout == 16 && l1.d_wu && l1.relu &&
                        l1.residual < 0 && l1.out_hdiv == 1 && l2.nseg == 1 && !l2.seg[0].up && l2.seg[0].cin == 16 &&
                        l2.cout == 16 && l2.relu && l2.residual < 0 && l2.src[0] == l1.out && l3.mode != kEpiBf16 &&
                        l3.nseg == 1 && l3.seg[0].cin == 16 && l3.src[0] == l2.out && l3.cout_pad == 16;
    if (shapes && a1.use_row && a2.use_row && a3.use_row) {
      ChainArgs& c = p.chain;
      memset(&c, 0, sizeof(c));
      c.w1 = a1.bmapu;
      c.w2 = a2.bmapf;
      c.w3 = a3.bmapf;
      c.src = p.bufs[l1.src[0]];
      c.b1 = l1.d_b;
      c.b2 = l2.d_b;
      c.batch = batch;
      c.h = h;
      c.w = w;
      c.fp16 = e->fp16;
      c.head = a3;
      p.use_chain = true;
    }
  }
  p.batch = batch;
  p.batch_pad = bp;
  p.h = h;
  p.w = w;
  return IU_OK;
}

// Make (batch, h, w) the current plan.  Plans are cached (the app alternates `predict_slice` on one slice with
// `predict_volumes` on batches: re-planning would free and re-allocate gigabytes on every switch); the least recently
// used ones are dropped beyond `plan_cache` entries, and all of them when an allocation fails.
int ensure_plan(iu_engine* e, int batch, int h, int w) {
  e->use_clock += 1;
  if (e->plan.batch == batch && e->plan.h == h && e->plan.w == w) return IU_OK;
  if (e->plan.batch != 0) {
    if (e->plan_cache > 0) {
      e->plan.last_use = e->use_clock - 1;
      e->cached.push_back(std::move(e->plan));
      e->plan = Plan();
    } else {
      cudaStreamSynchronize(e->stream);
      free_plan(e);
    }
  }
  for (size_t i = 0; i < e->cached.size(); ++i)
    if (e->cached[i].batch == batch && e->cached[i].h == h && e->cached[i].w == w) {
      e->plan = std::move(e->cached[i]);
      e->cached.erase(e->cached.begin() + i);
      return IU_OK;
    }
  while ((int)e->cached.size() > std::max(0, e->plan_cache - 1)) {
    size_t lru = 0;
    for (size_t i = 1; i < e->cached.size(); ++i)
      if (e->cached[i].last_use < e->cached[lru].last_use) lru = i;
    cudaStreamSynchronize(e->stream);
    release_plan(e->cached[lru]);
    e->cached.erase(e->cached.begin() + lru);
  }
  int rc = build_plan(e, batch, h, w);
  if (rc == IU_ERR_OOM && (!e->cached.empty() || !e->scratch.empty())) {
    // give everything idle back to the driver and try once more
    cudaStreamSynchronize(e->stream);
    free_cached_plans(e);
    for (auto& sc : e->scratch)
      if (!sc.used && sc.ptr) {
        cudaFree(sc.ptr);
        sc.ptr = nullptr;
        sc.bytes = 0;
      }
    cudaGetLastError();
    rc = build_plan(e, batch, h, w);
  }
  return rc;
}

// Slices per internal batch.  Bigger batches amortise the 46 launches per pass and their prologues (weight tiles,
// TMEM allocation); measured on B200 at 512^2: 37 / 74 / 148 / 256 slices -> 141.4 / 136.5 / 135.4 / 133.9 ms per
// 512^3 volume (keeping producer -> consumer tensors inside L2 with small batches does NOT pay).  Target 256 slices
// of 512^2 (12 GB of 16-bit activations), scaled by the image area, and split the work into equal batches so that no
// small ragged batch is left over.
int auto_batch(const iu_engine* e, int h, int w, int want) {
  const double rel = ((double)h * w) / (512.0 * 512.0);
  int target = std::max(1, (int)std::floor(256.0 / rel + 0.5));
  if (e->auto_batch_override > 0) target = e->auto_batch_override;
  if (e->max_batch > 0) target = std::min(target, e->max_batch);
  target = std::max(1, std::min(target, want));
  const int nbatches = (want + target - 1) / target;
  return (want + nbatches - 1) / nbatches;
}

// Kernel choice per conv (measurements in profiles/r01_findings.md):
//   * row-folded kernel (conv_row.cu): stride-1 3x3 layers with Cout <= 64 on images at least 128 wide -- three
//     vertical taps per MMA; weights resident, or streamed with the A chunks when they exceed shared memory;
//   * halo kernel (conv_halo.cu): the other layers that read an upsampled source (decoder blocks 0-1 conv1) and the
//     narrow layers on small images -- every activation fetched once per channel chunk instead of once per tap;
//   * per-tap TMA kernel (conv_tc.cu): Cout >= 128 layers, strided / 1x1 segments; L2-bandwidth bound, so Cout >= 256
//     layers use 256-wide weight tiles (one A box per 32 KB of weights).
// IU_CONV_VARIANT: 0 = automatic, 1 = per-tap kernel wherever it can run, 2 = halo kernel wherever it applies;
// IU_CONV_ROW=0 / IU_CONV_BN256=0 / IU_CONV_PAIR=1 switch the row-folded kernel, the wide tiles, the CTA-pair kernel.
cudaError_t launch_conv(iu_engine* e, const ConvArgs& a, int kc, int bn) {
  bool has_up = false;
  for (int s = 0; s < a.nseg; ++s) has_up |= a.seg[s].up != 0;
  if (a.use_row && e->conv_row && e->conv_variant != 1) return launch_conv_row(a, e->stream);
  const bool applicable = conv_halo_applicable(a);
  if (has_up && !applicable) return cudaErrorInvalidValue;
  if (e->conv_pair && e->conv_variant != 1 && kc == 64 && bn == 128 && conv_pair_applicable(a) &&
      (e->conv_variant == 2 || (a.out_h >= kHaloTile && a.out_w >= kHaloTile)))
    return launch_conv_pair(a, e->stream);
  bool halo;
  if (e->conv_variant == 1) halo = has_up;
  else if (e->conv_variant == 2) halo = applicable;
  else halo = has_up || (applicable && (kc <= 32 || bn <= 64) && a.out_h >= kHaloTile && a.out_w >= kHaloTile);
  if (halo) {
    if (a.small_bn && e->conv_small_bn && a.small_bn <= bn) {
      ConvArgs n = a;
      n.bmap = a.bmap_small;
      return launch_conv_halo(n, kc, a.small_bn, e->stream);
    }
    return launch_conv_halo(a, kc, bn, e->stream);
  }
  // stride-1 layers with identity sources and Cout >= 128 on images of at least one 16x16 block, when there are enough
  // blocks to fill the GPU: halo tile by TMA, nine taps through shifted swizzled descriptors (each activation crosses
  // L2 -> SM once per 128 output channels instead of once per tap)
  if (a.halo_tma && e->halo_tma && e->conv_variant != 1 && kc == 64 && !(a.small_bn && e->conv_small_bn))
    return launch_conv_halo_tma(a, e->stream);
  if (e->conv_pair2 && kc == 64 && a.mode == kEpiBf16) {
    const int bn2 = (a.use_bn256 && e->conv_bn256) ? 256 : bn;
    if (((bn2 == 256 && (e->conv_pair2 & 1)) || (bn2 == 128 && (e->conv_pair2 & 2))) && conv_tc2_applicable(a, bn2))
      return launch_conv_tc2(a, bn2, e->stream);
  }
  // Latency path: when the wide-tile shapes would give a handful of CTA tiles (one 256^2 slice: 2 pixel tiles in
  // layer3, 1 in layer4), every CTA streams its whole share of the layer's weights through ONE SM's TMA ring and the
  // launch takes ~25 us whatever the math.  The plan picked 64- or 32-wide Cout tiles for such layers (plan_small_bn).
  if (a.small_bn && e->conv_small_bn) {
    ConvArgs n = a;
    n.bmap = a.bmap_small;
    n.use_bn256 = 0;
    return launch_conv_tc(n, kc, a.small_bn, e->stream, 1, 1);
  }
  const int bm = (kc == 64 && a.mode == kEpiBf16) ? e->conv_bm : 1;
  // weight multicast across CTA pairs: layers whose Cout is ONE tile (Cout 256 with 256-wide tiles, Cout 128 with
  // paired pixel tiles)
  if (a.use_bn256 && e->conv_bn256 && kc == 64) {
    const int bm256 = (bm & 2) ? 2 : 1;
    const int cl = (e->conv_cluster && bm256 == 1 && a.cout == 256) ? 2 : 1;
    return launch_conv_tc(a, kc, 256, e->stream, bm256, cl);
  }
  const int bm128 = (bn == 128 && (bm & 1)) ? 2 : 1;
  const int cl = (e->conv_cluster && bm128 == 2 && a.cout == 128 && a.mode == kEpiBf16) ? 2 : 1;
  return launch_conv_tc(a, kc, bn, e->stream, bm128, cl);
}

// Run the network on the `batch` slices already in plan.x_in; the head writes according to (mode, out, ...).
int run_network(iu_engine* e, int batch, int head_mode, float* head_out, int slice0, int slice_count, int row_block) {
  Plan& p = e->plan;
  if (e->stem_pool) {
    // fused max-pool: the stem's epilogue pools its own tile; tiles combine their shared border windows with red.max
    // into the zeroed pooled tensor
    const size_t pooled = (size_t)batch * (p.h / 4) * (p.w / 4) * 64 * 2;
    prof_begin(e, IU_PROF_STEM);
    IU_CUDA(e, cudaMemsetAsync(p.bufs[e->t_p1], 0, pooled, e->stream));
    IU_CUDA(e, launch_conv_stem(e->stem_bmap, p.stem_omap, p.x_in, batch, p.h, p.w, p.stem_epi, e->stream,
                                p.bufs[e->t_p1]));
    prof_end(e);
    e->launches -= 1;  // one kernel where there were two
  } else {
    prof_begin(e, IU_PROF_STEM);
    IU_CUDA(e, launch_conv_stem(e->stem_bmap, p.stem_omap, p.x_in, batch, p.h, p.w, p.stem_epi, e->stream));
    prof_end(e);
    prof_begin(e, IU_PROF_POOL);
    IU_CUDA(e, launch_maxpool(p.bufs[e->t_f1], batch, p.h / 2, p.w / 2, 64, p.bufs[e->t_p1], e->stream));
    prof_end(e);
  }
  e->launches += 2;
  for (size_t i = 0; i < e->convs.size(); ++i) {
    const ConvLayer& L = e->convs[i];
    if (p.use_chain && i + 3 == e->convs.size()) {
      ChainArgs c = p.chain;
      c.batch = batch;
      c.head.batch = batch;
      c.head.mode = head_mode;
      c.head.out = head_out;
      c.head.slice0 = slice0;
      c.head.slice_count = slice_count;
      c.head.row_block = row_block;
      prof_begin(e, IU_PROF_CONV);
      cudaError_t ce = launch_conv_chain(c, e->stream);
      prof_end(e);
      if (ce != cudaSuccess) return e->cuda_fail(ce, "launch decoder tail (conv_chain)");
      e->launches += 1;
      break;
    }
    ConvArgs a = p.args[i];
    a.batch = batch;
    if (L.mode != kEpiBf16) {
      a.mode = head_mode;
      a.out = head_out;
      a.slice0 = slice0;
      a.slice_count = slice_count;
      a.row_block = row_block;
    }
    prof_begin(e, IU_PROF_CONV);
    cudaError_t ce = launch_conv(e, a, L.kc, L.bn);
    prof_end(e);
    if (ce != cudaSuccess) return e->cuda_fail(ce, ("launch " + L.name).c_str());
    e->launches += 1;
  }
  return IU_OK;
}

int finish(iu_engine* e, unsigned flags) {
  if (flags & IU_FLAG_ASYNC) return IU_OK;
  IU_CUDA(e, cudaStreamSynchronize(e->stream));
  return IU_OK;
}

bool check_engine(iu_engine* e, bool need_weights, int* rc, DeviceGuard* guard) {
  if (!e) {
    *rc = IU_ERR_INVALID;
    return false;
  }
  cudaError_t ce = guard->enter(e->device);
  if (ce != cudaSuccess) {
    *rc = e->cuda_fail(ce, "cudaSetDevice");
    return false;
  }
  if (need_weights && !e->loaded) {
    *rc = e->fail(IU_ERR_STATE, "weights not loaded (call iu_engine_load_weights first)");
    return false;
  }
  return true;
}

}  // namespace

// =========================================================================== C ABI
extern "C" {

int iu_abi_version(void) { return IU_ABI_VERSION; }

const char* iu_last_error(const iu_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int iu_engine_create(int device, iu_engine** out) {
  if (!out) return IU_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(ce) +
                     "); iunet_b200 has no CPU fallback";
    cudaGetLastError();
    return IU_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return IU_ERR_INVALID;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; iunet_b200 kernels are built for sm_100a only";
    return IU_ERR_CUDA;
  }
  DeviceGuard guard;
  ce = guard.enter(device);
  if (ce != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(ce);
    return IU_ERR_CUDA;
  }
  iu_engine* e = new iu_engine();
  e->device = device;
  ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) {
    g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(ce);
    delete e;
    return IU_ERR_CUDA;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (ce != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    g_create_error = "cuTensorMapEncodeTiled not available from the driver";
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  cudaStreamDestroy(e->stream);
    delete e;
    return IU_ERR_CUDA;
  }
  e->encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* v = getenv("IU_CONV_VARIANT")) e->conv_variant = atoi(v);
  if (const char* v = getenv("IU_AUTO_BATCH")) e->auto_batch_override = atoi(v);
  if (const char* v = getenv("IU_CONV_PAIR")) e->conv_pair = atoi(v);
  if (const char* v = getenv("IU_CONV_ROW")) e->conv_row = atoi(v);
  if (const char* v = getenv("IU_CONV_BN256")) e->conv_bn256 = atoi(v);
  if (const char* v = getenv("IU_CONV_BM2")) e->conv_bm = atoi(v);
  if (const char* v = getenv("IU_CONV_CLUSTER")) e->conv_cluster = atoi(v);
  if (const char* v = getenv("IU_PLAN_CACHE")) e->plan_cache = std::max(0, atoi(v));
  if (const char* v = getenv("IU_GRAPH")) e->use_graph = atoi(v);
  if (const char* v = getenv("IU_CONV_CHAIN")) e->conv_chain = atoi(v);
  if (const char* v = getenv("IU_CONV_PAIR2")) e->conv_pair2 = atoi(v);
  if (const char* v = getenv("IU_CONV_SMALL_BN")) e->conv_small_bn = atoi(v);
  if (const char* v = getenv("IU_STEM_POOL")) e->stem_pool = atoi(v);
  if (const char* v = getenv("IU_ROW_RES_TMA")) e->row_res_tma = atoi(v);
  if (const char* v = getenv("IU_ROW_TMA")) e->row_tma = atoi(v);
  if (const char* v = getenv("IU_HALO_TMA")) e->halo_tma = atoi(v);
  e->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
  if (const char* v = getenv("IU_SCRATCH_KEEP_MB")) e->scratch_keep = (size_t)std::max(0, atoi(v)) << 20;
  if (const char* v = getenv("IU_CONV_DEBUG")) {
    if (atoi(v) != 0 && cudaMalloc(&e->d_debug, 64 * 16 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(e->d_debug, 0, 64 * 16 * sizeof(unsigned long long));
  }
  *out = e;
  return IU_OK;
}

void iu_engine_destroy(iu_engine* e) {
  if (!e) return;
  DeviceGuard guard;
  guard.enter(e->device);
  cudaStreamSynchronize(e->stream);
  free_all_plans(e);
  free_weights(e);
  for (auto& s : e->scratch)
    if (s.ptr) cudaFree(s.ptr);
  if (e->d_debug) cudaFree(e->d_debug);
  if (e->d_window) cudaFree(e->d_window);
  for (auto& s : e->spans) {
    cudaEventDestroy(s.begin);
    cudaEventDestroy(s.end);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  cudaStreamDestroy(e->stream);
  delete e;
}

void* iu_engine_stream(iu_engine* e) { return e ? (void*)e->stream : nullptr; }

int iu_engine_synchronize(iu_engine* e) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  IU_CUDA(e, cudaStreamSynchronize(e->stream));
  return IU_OK;
}

int iu_engine_load_weights(iu_engine* e, int num_classes, int n_tensors, const char* const* names,
                           const float* const* data, const int64_t* numel) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (num_classes < 1 || num_classes > 16) return e->fail(IU_ERR_INVALID, "num_classes must be in [1, 16]");
  if (!names || !data || !numel) return e->fail(IU_ERR_INVALID, "null tensor table");
  cudaStreamSynchronize(e->stream);
  free_all_plans(e);
  free_weights(e);
  HostTensors ht;
  for (int i = 0; i < n_tensors; ++i) ht.t[names[i]] = {data[i], numel[i]};
  e->num_classes = num_classes;
  rc = build_network(e, ht, num_classes);
  if (rc != IU_OK) {
    const std::string keep = e->err;
    free_weights(e);
    e->err = keep;
    return rc;
  }
  e->loaded = true;
  return IU_OK;
}

int iu_engine_num_classes(const iu_engine* e) { return e ? e->num_classes : 0; }

int iu_engine_set_precision(iu_engine* e, int precision) {
  if (!e) return IU_ERR_INVALID;
  if (precision != IU_PRECISION_FP16 && precision != IU_PRECISION_BF16)
    return e->fail(IU_ERR_INVALID, "precision must be IU_PRECISION_FP16 or IU_PRECISION_BF16");
  const int fp16 = precision == IU_PRECISION_FP16;
  DeviceGuard guard;
  if (fp16 != e->fp16 && e->loaded) {
    guard.enter(e->device);
    cudaStreamSynchronize(e->stream);
    free_all_plans(e);
    free_weights(e);  // packed weights are format specific: the caller must load them again
  }
  e->fp16 = fp16;
  return IU_OK;
}
int iu_engine_precision(const iu_engine* e) { return e ? (e->fp16 ? IU_PRECISION_FP16 : IU_PRECISION_BF16) : -1; }

int iu_engine_set_max_batch(iu_engine* e, int max_batch) {
  if (!e || max_batch < 0) return IU_ERR_INVALID;
  e->max_batch = max_batch;
  return IU_OK;
}

int64_t iu_engine_workspace_bytes(iu_engine* e, int batch, int h, int w) {
  if (!e || !e->loaded || batch < 1 || h % 32 || w % 32) return -1;
  return (int64_t)(e->weight_bytes + plan_bytes(e, (batch + 7) / 8 * 8, h, w));
}

int64_t iu_engine_launch_count(const iu_engine* e) { return e ? e->launches : 0; }

int iu_engine_auto_batch(const iu_engine* e, int h, int w, int count) {
  if (!e || h < 1 || w < 1 || count < 1) return 0;
  return auto_batch(e, h, w, count);
}

int iu_engine_release_workspace(iu_engine* e) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  IU_CUDA(e, cudaStreamSynchronize(e->stream));
  if (e->copy_stream) IU_CUDA(e, cudaStreamSynchronize(e->copy_stream));
  free_all_plans(e);
  const size_t keep = e->scratch_keep;
  e->scratch_keep = 0;
  scratch_trim(e);
  e->scratch_keep = keep;
  return IU_OK;
}

int64_t iu_engine_held_bytes(const iu_engine* e) {
  if (!e) return -1;
  size_t total = e->weight_bytes + e->plan.bytes;
  for (const auto& p : e->cached) total += p.bytes;
  for (const auto& sc : e->scratch) total += sc.bytes;
  return (int64_t)total;
}

int iu_engine_debug_counters(iu_engine* e, unsigned long long* out, int n_values, int reset) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!e->d_debug) return e->fail(IU_ERR_STATE, "debug counters are off (set IU_CONV_DEBUG=1 before creating the engine)");
  if (!out || n_values < 1 || n_values > 64 * 16) return e->fail(IU_ERR_INVALID, "debug_counters: bad arguments");
  IU_CUDA(e, cudaStreamSynchronize(e->stream));
  IU_CUDA(e, cudaMemcpy(out, e->d_debug, (size_t)n_values * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) IU_CUDA(e, cudaMemset(e->d_debug, 0, 64 * 16 * sizeof(unsigned long long)));
  return IU_OK;
}

int iu_engine_profile(iu_engine* e, int enable) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  prof_flush(e);
  e->prof = enable != 0;
  return IU_OK;
}

int iu_engine_profile_read(iu_engine* e, double* ms, int64_t* count, int reset) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!ms || !count) return e->fail(IU_ERR_INVALID, "profile_read: null output");
  prof_flush(e);
  for (int i = 0; i < IU_PROF_CLASSES; ++i) {
    ms[i] = e->prof_ms[i];
    count[i] = e->prof_n[i];
    if (reset) {
      e->prof_ms[i] = 0;
      e->prof_n[i] = 0;
    }
  }
  return IU_OK;
}

int iu_engine_forward(iu_engine* e, const float* x, int batch, int h, int w, float* probs, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, true, &rc, &guard)) return rc;
  if (!x || !probs || batch < 1) return e->fail(IU_ERR_INVALID, "forward: null pointer or empty batch");
  if (h < 32 || w < 32 || h % 32 || w % 32)
    return e->fail(IU_ERR_INVALID, "Wrong input shape height=" + std::to_string(h) + ", width=" + std::to_string(w) +
                                       ". Expected image height and width divisible by 32.");
  const int c = e->num_classes;
  const int bs = auto_batch(e, h, w, batch);
  rc = ensure_plan(e, bs, h, w);
  if (rc != IU_OK) return rc;
  const bool x_dev = is_device_ptr(x), out_dev = is_device_ptr(probs);
  const size_t img = (size_t)h * w;
  // Latency path (one internal batch of at most 64 slices of 512^2, e.g. `predict_slice` on one 256^2 slice): from
  // the second call on a plan its launches run as ONE CUDA graph reading plan.x_in and writing plan.fwd_out.
  if (e->use_graph && bs == batch && !e->prof && !e->d_debug && (size_t)batch * img <= (size_t)64 * 512 * 512) {
    Plan& p = e->plan;
    const size_t out_bytes = (size_t)batch * c * img * 4;
    if (!p.fwd_out) {
      cudaError_t ce = cudaMalloc(&p.fwd_out, out_bytes);
      if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaMalloc(forward output)");
    }
    cudaError_t ce = cudaMemcpyAsync(p.x_in, x, (size_t)batch * img * 4,
                                     x_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaMemcpyAsync(x)");
    p.fwd_calls += 1;
    if (p.graph == nullptr && p.fwd_calls >= 2) {
      // the first call ran eagerly (it also configured every kernel's attributes); capture this one
      const int64_t before = e->launches;
      cudaGraph_t g = nullptr;
      ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
      if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaStreamBeginCapture");
      rc = run_network(e, batch, kEpiSoftmaxNCHW, p.fwd_out, 0, batch, h);
      ce = cudaStreamEndCapture(e->stream, &g);
      p.graph_launches = (int)(e->launches - before);
      e->launches = before;
      if (rc != IU_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      if (ce == cudaSuccess) ce = cudaGraphInstantiate(&p.graph, g, 0);
      if (g) cudaGraphDestroy(g);
      if (ce != cudaSuccess) {
        p.graph = nullptr;
        return e->cuda_fail(ce, "capture of the forward graph");
      }
    }
    if (p.graph) {
      ce = cudaGraphLaunch(p.graph, e->stream);
      if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaGraphLaunch");
      e->launches += p.graph_launches;
    } else {
      rc = run_network(e, batch, kEpiSoftmaxNCHW, p.fwd_out, 0, batch, h);
      if (rc != IU_OK) return rc;
    }
    ce = cudaMemcpyAsync(probs, p.fwd_out, out_bytes, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                         e->stream);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "copy probabilities");
    if (!out_dev) {
      IU_CUDA(e, cudaStreamSynchronize(e->stream));
      return IU_OK;
    }
    return finish(e, flags);
  }
  float* stage = nullptr;
  if (!out_dev) {
    rc = scratch_get(e, (size_t)batch * c * img * 4, (void**)&stage);
    if (rc != IU_OK) return rc;
  }
  float* out_base = out_dev ? probs : stage;
  for (int s = 0; s < batch; s += bs) {
    const int b = std::min(bs, batch - s);
    cudaError_t ce = cudaMemcpyAsync(e->plan.x_in, x + (size_t)s * img, (size_t)b * img * 4,
                                     x_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) {
      scratch_put(e, stage);
      return e->cuda_fail(ce, "cudaMemcpyAsync(x)");
    }
    rc = run_network(e, b, kEpiSoftmaxNCHW, out_base + (size_t)s * c * img, 0, b, h);
    if (rc != IU_OK) {
      scratch_put(e, stage);
      return rc;
    }
  }
  if (!out_dev) {
    cudaError_t ce = cudaMemcpyAsync(probs, stage, (size_t)batch * c * img * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    scratch_put(e, stage);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "copy probabilities to host");
    return IU_OK;
  }
  return finish(e, flags);
}

int iu_engine_gather_slices(iu_engine* e, const void* volume_dev, int dtype, int n, int axis, int start, int count,
                            float* out_dev, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!volume_dev || !out_dev || n < 16 || n % 16 || axis < 0 || axis > 2 || start < 0 || count < 1 ||
      start + count > n || (dtype != IU_DTYPE_U8 && dtype != IU_DTYPE_F32))
    return e->fail(IU_ERR_INVALID, "gather_slices: bad arguments");
  prof_begin(e, IU_PROF_GATHER);
  IU_CUDA(e, launch_gather_slices(volume_dev, dtype == IU_DTYPE_F32, n, axis, start, count, out_dev, e->stream));
  prof_end(e);
  e->launches += 1;
  return finish(e, flags);
}

// Slices from any strided source (device): element (slice i, row r, col c) = base[i*ss + r*sr + c*sc].
// `plan_batch` > 0: run on the activation plan of that batch size even if `count` is smaller (a partial batch), so that
// a caller which interleaves short and long slice ranges does not re-plan in between.
static int predict_slices_impl(iu_engine* e, const void* base_dev, int dtype, int count, int h, int w, long long ss,
                               long long sr, long long sc, float* probs_dev, int slice_offset, int slice_total,
                               int row_block, unsigned flags, int plan_batch) {
  const size_t esz = dtype == IU_DTYPE_F32 ? 4 : 1;
  const int bs = plan_batch > 0 ? plan_batch : auto_batch(e, h, w, count);
  int rc = ensure_plan(e, bs, h, w);
  for (int s = 0; rc == IU_OK && s < count; s += bs) {
    const int b = std::min(bs, count - s);
    prof_begin(e, IU_PROF_GATHER);
    cudaError_t ce = launch_gather_strided(static_cast<const char*>(base_dev) + (size_t)s * ss * esz,
                                           dtype == IU_DTYPE_F32, b, h, w, ss, sr, sc, e->plan.x_in, e->stream);
    prof_end(e);
    if (ce != cudaSuccess) {
      rc = e->cuda_fail(ce, "launch gather_slices");
      break;
    }
    e->launches += 1;
    rc = run_network(e, b, kEpiSoftmaxNHWC, probs_dev, slice_offset + s, slice_total, row_block);
  }
  if (rc != IU_OK) return rc;
  return finish(e, flags);
}

int iu_engine_predict_slices(iu_engine* e, const void* base_dev, int dtype, int count, int h, int w,
                             int64_t stride_slice, int64_t stride_row, int64_t stride_col, float* probs_dev,
                             int slice_offset, int slice_total, int row_block, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, true, &rc, &guard)) return rc;
  if (!base_dev || !probs_dev || count < 1 || (dtype != IU_DTYPE_U8 && dtype != IU_DTYPE_F32) || stride_slice < 0 ||
      stride_row < 0 || stride_col < 0 || row_block < 1 || slice_offset < 0 || slice_offset + count > slice_total)
    return e->fail(IU_ERR_INVALID, "predict_slices: bad arguments");
  if (h < 32 || w < 32 || h % 32 || w % 32)
    return e->fail(IU_ERR_INVALID, "Wrong input shape height=" + std::to_string(h) + ", width=" + std::to_string(w) +
                                       ". Expected image height and width divisible by 32.");
  if (h % row_block) return e->fail(IU_ERR_INVALID, "predict_slices: row_block must divide the image height");
  if (!is_device_ptr(base_dev)) return e->fail(IU_ERR_INVALID, "predict_slices: the slice source must be device memory");
  return predict_slices_impl(e, base_dev, dtype, count, h, w, stride_slice, stride_row, stride_col, probs_dev,
                             slice_offset, slice_total, row_block, flags, 0);
}

static int predict_axis_impl(iu_engine* e, const void* volume, int dtype, int n, int axis, int slice_begin,
                             int slice_count, float* probs_dev, int slice_offset, int slice_total, int row_block,
                             unsigned flags, int plan_batch) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, true, &rc, &guard)) return rc;
  if (!volume || !probs_dev || n < 32 || n % 32 || axis < 0 || axis > 2 || slice_begin < 0 || slice_count < 1 ||
      slice_begin + slice_count > n || (dtype != IU_DTYPE_U8 && dtype != IU_DTYPE_F32) || row_block < 1 ||
      n % row_block || slice_offset < 0 || slice_offset + slice_count > slice_total)
    return e->fail(IU_ERR_INVALID, "predict_axis: bad arguments");
  const size_t esz = dtype == IU_DTYPE_F32 ? 4 : 1;
  const void* vol_dev = volume;
  void* staged = nullptr;
  if (!is_device_ptr(volume)) {
    rc = scratch_get(e, (size_t)n * n * n * esz, &staged);
    if (rc != IU_OK) return rc;
    cudaError_t ce = cudaMemcpyAsync(staged, volume, (size_t)n * n * n * esz, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) {
      scratch_put(e, staged);
      return e->cuda_fail(ce, "cudaMemcpyAsync(volume)");
    }
    vol_dev = staged;
  }
  // volume [z][y][x]: a slice along axis a is image (y,x) | (z,x) | (z,y)
  const long long nn = (long long)n * n;
  const long long ss = axis == 0 ? nn : (axis == 1 ? n : 1), sr = axis == 0 ? n : nn, sc = axis == 2 ? n : 1;
  rc = predict_slices_impl(e, static_cast<const char*>(vol_dev) + (size_t)slice_begin * ss * esz, dtype, slice_count, n,
                           n, ss, sr, sc, probs_dev, slice_offset, slice_total, row_block,
                           staged ? (flags | IU_FLAG_ASYNC) : flags, plan_batch);
  if (staged) {
    cudaStreamSynchronize(e->stream);
    scratch_put(e, staged);
  }
  return rc;
}

int iu_engine_predict_axis(iu_engine* e, const void* volume, int dtype, int n, int axis, int slice_begin,
                           int slice_count, float* probs_dev, int slice_offset, int slice_total, int row_block,
                           unsigned flags) {
  return predict_axis_impl(e, volume, dtype, n, axis, slice_begin, slice_count, probs_dev, slice_offset, slice_total,
                           row_block, flags, 0);
}

// The window factor lives in a device buffer the engine keeps (uploaded again only when its values change), so a
// reduce call neither allocates nor has to drain the stream to give a staging block back.
static int window_on_device(iu_engine* e, const float* g1d_host, int n, const float** out) {
  *out = nullptr;
  if (!g1d_host) return IU_OK;
  const bool same = e->d_window && (int)e->window_host.size() == n &&
                    memcmp(e->window_host.data(), g1d_host, (size_t)n * 4) == 0;
  if (!same) {
    if ((int)e->window_host.size() != n || !e->d_window) {
      IU_CUDA(e, cudaStreamSynchronize(e->stream));  // kernels that still read the old table
      if (e->d_window) cudaFree(e->d_window);
      e->d_window = nullptr;
      IU_CUDA(e, cudaMalloc(&e->d_window, (size_t)n * 4));
    } else {
      IU_CUDA(e, cudaStreamSynchronize(e->stream));
    }
    e->window_host.assign(g1d_host, g1d_host + n);
    IU_CUDA(e, cudaMemcpyAsync(e->d_window, e->window_host.data(), (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  }
  *out = e->d_window;
  return IU_OK;
}

int iu_engine_reduce_planes(iu_engine* e, const float* p0, const float* p1, const float* p2, const int* order, int n_axes,
                            int n, int t, int z0, int zoff, int zcount, int num_classes, const float* g1d_host,
                            float gmax, float lo, uint8_t* out_u8, uint8_t* out_labels, float* out_mean, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!order || n_axes < 1 || n_axes > 3 || n < 1 || t < 1 || z0 < 0 || z0 + t > n || num_classes < 1 ||
      num_classes > 10 || zoff < 0 || zcount < 1 || zoff + zcount > t)
    return e->fail(IU_ERR_INVALID, "reduce: bad arguments");
  ReduceArgs a;
  memset(&a, 0, sizeof(a));
  a.p[0] = p0;
  a.p[1] = p1;
  a.p[2] = p2;
  for (int i = 0; i < 3; ++i) a.order[i] = 0;
  for (int i = 0; i < n_axes; ++i) {
    if (order[i] < 0 || order[i] > 2 || a.p[order[i]] == nullptr)
      return e->fail(IU_ERR_INVALID, "reduce: axis in `order` has no probability buffer");
    a.order[i] = order[i];
  }
  a.n_axes = n_axes;
  a.n = n;
  a.t = t;
  a.z0 = z0;
  a.zoff = zoff;
  a.zcount = zcount;
  a.num_classes = num_classes;
  a.gmax = gmax;
  a.lo = lo;
  a.out_u8 = out_u8;
  a.out_labels = out_labels;
  a.out_mean = out_mean;
  if ((rc = window_on_device(e, g1d_host, n, &a.g1d)) != IU_OK) return rc;
  prof_begin(e, IU_PROF_REDUCE);
  cudaError_t ce = launch_reduce(a, e->stream);
  prof_end(e);
  e->launches += 1;
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch reduce");
  return finish(e, flags);
}

int iu_engine_reduce(iu_engine* e, const float* p0, const float* p1, const float* p2, const int* order, int n_axes,
                     int n, int t, int z0, int num_classes, const float* g1d_host, float gmax, float lo,
                     uint8_t* out_u8, uint8_t* out_labels, float* out_mean, unsigned flags) {
  return iu_engine_reduce_planes(e, p0, p1, p2, order, n_axes, n, t, z0, 0, t, num_classes, g1d_host, gmax, lo, out_u8,
                                 out_labels, out_mean, flags);
}

int iu_engine_predict_volume(iu_engine* e, const void* volume, int dtype, int n, const int* axes, int n_axes,
                             const float* g1d_host, float gmax, float lo, uint8_t* out_u8, uint8_t* out_labels,
                             float* out_mean, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, true, &rc, &guard)) return rc;
  if (!volume || !axes || n_axes < 1 || n_axes > 3 || n < 32 || n % 32 ||
      (dtype != IU_DTYPE_U8 && dtype != IU_DTYPE_F32))
    return e->fail(IU_ERR_INVALID, "predict_volume: bad arguments");
  bool seen[3] = {false, false, false};
  for (int i = 0; i < n_axes; ++i) {
    if (axes[i] < 0 || axes[i] > 2 || seen[axes[i]])
      return e->fail(IU_ERR_INVALID, "predict_volume: axes must be distinct values from {0,1,2}");
    seen[axes[i]] = true;
  }
  const int c = e->num_classes;
  const size_t vox = (size_t)n * n * n;
  const size_t esz = dtype == IU_DTYPE_F32 ? 4 : 1;
  std::vector<void*> held;
  auto release = [&]() {
    cudaStreamSynchronize(e->stream);
    for (void* p : held) scratch_put(e, p);
  };
  auto grab = [&](size_t bytes, void** out) {
    int r = scratch_get(e, bytes, out);
    if (r == IU_OK) held.push_back(*out);
    return r;
  };
  const void* vol_dev = volume;
  const bool vol_host = !is_device_ptr(volume);
  void* staged = nullptr;
  if (vol_host) {
    if ((rc = grab(vox * esz, &staged)) != IU_OK) { release(); return rc; }
    vol_dev = staged;     // uploaded below, once the schedule is known
  }
  float* p[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < n_axes; ++i) {
    if ((rc = grab(vox * c * 4, (void**)&p[axes[i]])) != IU_OK) { release(); return rc; }
  }
  const bool u8_dev = out_u8 && is_device_ptr(out_u8);
  const bool lab_dev = out_labels && is_device_ptr(out_labels);
  const bool mean_dev = out_mean && is_device_ptr(out_mean);
  uint8_t* d_u8 = out_u8;
  uint8_t* d_lab = out_labels;
  float* d_mean = out_mean;
  if (out_u8 && !u8_dev && (rc = grab(vox * c, (void**)&d_u8)) != IU_OK) { release(); return rc; }
  if (out_labels && !lab_dev && (rc = grab(vox, (void**)&d_lab)) != IU_OK) { release(); return rc; }
  if (out_mean && !mean_dev && (rc = grab(vox * c * 4, (void**)&d_mean)) != IU_OK) { release(); return rc; }
  const bool to_host = (out_u8 && !u8_dev) || (out_labels && !lab_dev) || (out_mean && !mean_dev);

  // Results for the host: the order in which the axes are COMPUTED does not enter the arithmetic (each axis has its
  // own buffer; K4 adds them in the caller's order), so axis 0 runs last, in parts of the z range, and each part is
  // reduced and copied out on a second stream while the next one is still in the network.
  if (to_host && seen[0]) {
    if (!e->copy_stream) {
      cudaError_t ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
      if (ce != cudaSuccess) { release(); return e->cuda_fail(ce, "cudaStreamCreate(copy)"); }
    }
    float* g_dev = nullptr;
    if (g1d_host) {
      if ((rc = grab((size_t)n * 4, (void**)&g_dev)) != IU_OK) { release(); return rc; }
      cudaError_t ce = cudaMemcpyAsync(g_dev, g1d_host, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream);
      if (ce != cudaSuccess) { release(); return e->cuda_fail(ce, "cudaMemcpyAsync(window)"); }
    }
    ReduceArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 3; ++i) a.p[i] = p[i];
    for (int i = 0; i < n_axes; ++i) a.order[i] = axes[i];
    a.n_axes = n_axes;
    a.n = n;
    a.t = n;
    a.num_classes = c;
    a.g1d = g_dev;
    a.gmax = gmax;
    a.lo = lo;
    a.out_u8 = out_u8 ? d_u8 : nullptr;
    a.out_labels = out_labels ? d_lab : nullptr;
    a.out_mean = out_mean ? d_mean : nullptr;
    // Parts: whole internal batches (full kernel efficiency), except that the LAST batch is split in two so that the
    // result copy left exposed at the very end is half a batch's worth; all run on the activation plan of the other axes.
    const int plan_batch = auto_batch(e, n, n, n);
    std::vector<int> part_begin;
    for (int zs = 0; zs < n;) {
      part_begin.push_back(zs);
      const int left = n - zs;
      zs += (left <= plan_batch && plan_batch >= 64 && left > plan_batch / 2) ? (left + 1) / 2 : std::min(plan_batch, left);
    }
    part_begin.push_back(n);
    const int kParts = (int)part_begin.size() - 1;
    std::vector<cudaEvent_t> done((size_t)kParts, nullptr);
    cudaError_t ce = cudaSuccess;
    auto copy_part = [&](int k) {
      const int zs = part_begin[k], cnt = part_begin[k + 1] - zs;
      const size_t plane = (size_t)n * n, off = (size_t)zs * plane, len = (size_t)cnt * plane;
      cudaError_t r = cudaStreamWaitEvent(e->copy_stream, done[k], 0);
      if (r == cudaSuccess && out_u8 && !u8_dev)
        r = cudaMemcpyAsync(out_u8 + off * c, d_u8 + off * c, len * c, cudaMemcpyDeviceToHost, e->copy_stream);
      if (r == cudaSuccess && out_labels && !lab_dev)
        r = cudaMemcpyAsync(out_labels + off, d_lab + off, len, cudaMemcpyDeviceToHost, e->copy_stream);
      if (r == cudaSuccess && out_mean && !mean_dev)
        r = cudaMemcpyAsync(out_mean + off * c, d_mean + off * c, len * c * 4, cudaMemcpyDeviceToHost, e->copy_stream);
      return r;
    };
    // Upload: the planes of the FIRST part on the compute stream, the rest on the copy stream under that part's
    // network passes (axis-0 slices are z planes: part 0 needs only its own); then axis 0 part 0, the other axes, and
    // the remaining parts of axis 0, each reduced and copied out under the next one.
    cudaEvent_t uploaded = nullptr;
    if (vol_host) {
      const size_t first = (size_t)part_begin[1] * n * n * esz;
      ce = cudaMemcpyAsync(staged, volume, first, cudaMemcpyHostToDevice, e->stream);
      if (ce == cudaSuccess && first < vox * esz) {
        ce = cudaMemcpyAsync(static_cast<char*>(staged) + first, static_cast<const char*>(volume) + first, vox * esz - first,
                             cudaMemcpyHostToDevice, e->copy_stream);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventRecord(uploaded, e->copy_stream);
      }
    }
    int parts = 0;
    for (int k = 0; k < kParts && rc == IU_OK && ce == cudaSuccess; ++k, ++parts) {
      const int zs = part_begin[k], cnt = part_begin[k + 1] - zs;
      rc = predict_axis_impl(e, vol_dev, dtype, n, 0, zs, cnt, p[0], zs, n, n, IU_FLAG_ASYNC, plan_batch);
      if (rc != IU_OK) break;
      if (k == 0) {
        if (uploaded) ce = cudaStreamWaitEvent(e->stream, uploaded, 0);
        for (int i = 0; i < n_axes && rc == IU_OK && ce == cudaSuccess; ++i) {
          if (axes[i] == 0) continue;
          rc = iu_engine_predict_axis(e, vol_dev, dtype, n, axes[i], 0, n, p[axes[i]], 0, n, n, IU_FLAG_ASYNC);
        }
        if (rc != IU_OK || ce != cudaSuccess) break;
      }
      a.zoff = zs;
      a.zcount = cnt;
      prof_begin(e, IU_PROF_REDUCE);
      ce = launch_reduce(a, e->stream);
      prof_end(e);
      e->launches += 1;
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventRecord(done[k], e->stream);
      // the copy of part k-1 is issued AFTER part k's kernels are queued: with pageable destinations the copy call
      // blocks this thread, and the device then still has a part's worth of work
      if (ce == cudaSuccess && k > 0) ce = copy_part(k - 1);
    }
    if (rc == IU_OK && ce == cudaSuccess && parts > 0) ce = copy_part(parts - 1);
    cudaError_t ce2 = cudaStreamSynchronize(e->stream);
    cudaError_t ce3 = cudaStreamSynchronize(e->copy_stream);
    for (int k = 0; k < kParts; ++k)
      if (done[k]) cudaEventDestroy(done[k]);
    if (uploaded) cudaEventDestroy(uploaded);
    for (void* q : held) scratch_put(e, q);
    scratch_trim(e);
    if (rc != IU_OK) return rc;
    if (ce != cudaSuccess) return e->cuda_fail(ce, "predict_volume (pipelined results)");
    if (ce2 != cudaSuccess) return e->cuda_fail(ce2, "predict_volume");
    if (ce3 != cudaSuccess) return e->cuda_fail(ce3, "predict_volume (copy stream)");
    return IU_OK;
  }

  if (vol_host) {
    cudaError_t ce = cudaMemcpyAsync(staged, volume, vox * esz, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) { release(); return e->cuda_fail(ce, "cudaMemcpyAsync(volume)"); }
  }
  for (int i = 0; i < n_axes; ++i) {
    rc = iu_engine_predict_axis(e, vol_dev, dtype, n, axes[i], 0, n, p[axes[i]], 0, n, n, IU_FLAG_ASYNC);
    if (rc != IU_OK) { release(); return rc; }
  }
  rc = iu_engine_reduce(e, p[0], p[1], p[2], axes, n_axes, n, n, 0, c, g1d_host, gmax, lo, d_u8, d_lab, d_mean,
                        IU_FLAG_ASYNC);
  if (rc != IU_OK) { release(); return rc; }
  cudaError_t ce = cudaSuccess;
  if (out_u8 && !u8_dev) ce = cudaMemcpyAsync(out_u8, d_u8, vox * c, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess && out_labels && !lab_dev)
    ce = cudaMemcpyAsync(out_labels, d_lab, vox, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess && out_mean && !mean_dev)
    ce = cudaMemcpyAsync(out_mean, d_mean, vox * c * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce != cudaSuccess) { release(); return e->cuda_fail(ce, "copy results to host"); }
  ce = cudaStreamSynchronize(e->stream);
  for (void* q : held) scratch_put(e, q);
  scratch_trim(e);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "predict_volume");
  return IU_OK;
}

int iu_engine_extract_block(iu_engine* e, const uint8_t* volume_dev, int d, int h, int w, int i0, int j0, int k0, int s,
                            uint8_t* out_dev, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!volume_dev || !out_dev || d < 1 || h < 1 || w < 1 || s < 1)
    return e->fail(IU_ERR_INVALID, "extract_block: bad arguments");
  cudaError_t ce = launch_extract_block(volume_dev, d, h, w, i0, j0, k0, s, out_dev, e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch extract_block (the block must intersect the volume)");
  e->launches += 1;
  return finish(e, flags);
}

int iu_engine_predict_tiled(iu_engine* e, const uint8_t* volume, int d, int h, int w, int s, int n_blocks,
                            const int* origins, const int* axes, int n_axes, const float* g1d_host, float gmax,
                            float lo, uint8_t* out_u8, uint8_t* out_labels, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, true, &rc, &guard)) return rc;
  if (!volume || !origins || !axes || !g1d_host || n_blocks < 1 || d < 1 || h < 1 || w < 1 || s < 32 || s % 32 ||
      n_axes < 1 || n_axes > 3 || (!out_u8 && !out_labels))
    return e->fail(IU_ERR_INVALID, "predict_tiled: bad arguments (block edge must be a multiple of 32)");
  bool seen[3] = {false, false, false};
  for (int i = 0; i < n_axes; ++i) {
    if (axes[i] < 0 || axes[i] > 2 || seen[axes[i]])
      return e->fail(IU_ERR_INVALID, "predict_tiled: axes must be distinct values from {0,1,2}");
    seen[axes[i]] = true;
  }
  const int c = e->num_classes;
  const size_t vox = (size_t)d * h * w, bvox = (size_t)s * s * s;
  std::vector<void*> held;
  std::vector<cudaEvent_t> events;  // one per z range handed to the copy stream
  auto drop_events = [&]() {
    for (cudaEvent_t ev : events) cudaEventDestroy(ev);
    events.clear();
  };
  auto release = [&]() {
    cudaStreamSynchronize(e->stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    drop_events();
    for (void* p : held) scratch_put(e, p);
  };
  auto grab = [&](size_t bytes, void** out) {
    int r = scratch_get(e, bytes, out);
    if (r == IU_OK) held.push_back(*out);
    return r;
  };
#define IU_TILED_TRY(expr_)  \
  if ((rc = (expr_)) != IU_OK) { \
    release();               \
    return rc;               \
  }
#define IU_TILED_CUDA(call_, what_)            \
  {                                            \
    cudaError_t ce_ = (call_);                 \
    if (ce_ != cudaSuccess) {                  \
      release();                               \
      return e->cuda_fail(ce_, what_);         \
    }                                          \
  }
  const uint8_t* vol_dev = volume;
  if (!is_device_ptr(volume)) {
    void* staged = nullptr;
    IU_TILED_TRY(grab(vox, &staged));
    IU_TILED_CUDA(cudaMemcpyAsync(staged, volume, vox, cudaMemcpyHostToDevice, e->stream), "cudaMemcpyAsync(volume)");
    vol_dev = (const uint8_t*)staged;
  }
  // The fp32 accumulators of predict.py:181-198 (`pred`, `weight`; on disk in the reference) cover a RING of z planes,
  // not the volume: blocks arrive layer by layer in z (the reference's nested i, j, k order), so once a layer's first
  // block starts at plane c0 every plane below c0 is final -- it is normalised, written to the uint8 output (and copied
  // to the host on a second stream) and its ring planes are zeroed for the next layer.  One block edge of planes
  // suffices; a caller whose blocks are not ordered by z gets a ring as deep as the volume (no early hand-back).
  bool z_ordered = true;
  for (int b = 1; b < n_blocks; ++b) z_ordered &= origins[3 * b] >= origins[3 * (b - 1)];
  const int zring = z_ordered ? std::min(d, s) : d;
  const size_t plane = (size_t)h * w;
  float *pred = nullptr, *weight = nullptr, *g_dev = nullptr;
  uint8_t* block = nullptr;
  float* p[3] = {nullptr, nullptr, nullptr};
  IU_TILED_TRY(grab((size_t)zring * plane * c * 4, (void**)&pred));
  IU_TILED_TRY(grab((size_t)zring * plane * 4, (void**)&weight));
  IU_TILED_TRY(grab(bvox, (void**)&block));
  IU_TILED_TRY(grab((size_t)s * 4, (void**)&g_dev));
  for (int i = 0; i < n_axes; ++i) IU_TILED_TRY(grab(bvox * c * 4, (void**)&p[axes[i]]));
  const bool u8_dev = out_u8 && is_device_ptr(out_u8), lab_dev = out_labels && is_device_ptr(out_labels);
  uint8_t *d_u8 = out_u8, *d_lab = out_labels;
  if (out_u8 && !u8_dev) IU_TILED_TRY(grab(vox * c, (void**)&d_u8));
  if (out_labels && !lab_dev) IU_TILED_TRY(grab(vox, (void**)&d_lab));
  const bool to_host = (out_u8 && !u8_dev) || (out_labels && !lab_dev);
  if (to_host && !e->copy_stream)
    IU_TILED_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
  IU_TILED_CUDA(cudaMemsetAsync(pred, 0, (size_t)zring * plane * c * 4, e->stream), "cudaMemsetAsync(pred)");
  IU_TILED_CUDA(cudaMemsetAsync(weight, 0, (size_t)zring * plane * 4, e->stream), "cudaMemsetAsync(weight)");
  IU_TILED_CUDA(cudaMemcpyAsync(g_dev, g1d_host, (size_t)s * 4, cudaMemcpyHostToDevice, e->stream), "cudaMemcpyAsync(window)");
  // planes [za, zb) are final: normalise (predict.py:252-255), hand the ring planes back, copy the range out
  auto finish_planes = [&](int za, int zb) -> cudaError_t {
    cudaError_t ce = cudaSuccess;
    for (int z = za; z < zb && ce == cudaSuccess;) {
      const int r0 = z % zring;
      const int cnt = std::min(zb - z, zring - r0);        // contiguous in the ring
      const size_t vx = (size_t)cnt * plane;
      ce = launch_finalise(pred + (size_t)r0 * plane * c, weight + (size_t)r0 * plane, vx, c,
                           d_u8 ? d_u8 + (size_t)z * plane * c : nullptr, d_lab ? d_lab + (size_t)z * plane : nullptr, e->stream);
      e->launches += 1;
      if (ce == cudaSuccess) ce = cudaMemsetAsync(pred + (size_t)r0 * plane * c, 0, vx * c * 4, e->stream);
      if (ce == cudaSuccess) ce = cudaMemsetAsync(weight + (size_t)r0 * plane, 0, vx * 4, e->stream);
      z += cnt;
    }
    if (ce == cudaSuccess && to_host && zb > za) {
      cudaEvent_t ev = nullptr;
      ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (ce != cudaSuccess) return ce;
      events.push_back(ev);
      ce = cudaEventRecord(ev, e->stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->copy_stream, ev, 0);
      const size_t off = (size_t)za * plane, len = (size_t)(zb - za) * plane;
      if (ce == cudaSuccess && out_u8 && !u8_dev)
        ce = cudaMemcpyAsync(out_u8 + off * c, d_u8 + off * c, len * c, cudaMemcpyDeviceToHost, e->copy_stream);
      if (ce == cudaSuccess && out_labels && !lab_dev)
        ce = cudaMemcpyAsync(out_labels + off, d_lab + off, len, cudaMemcpyDeviceToHost, e->copy_stream);
    }
    return ce;
  };
  int zlo = 0;                                             // first plane that is not final yet
  for (int b = 0; b < n_blocks; ++b) {                     // the reference's block order (predict.py:235)
    const int i0 = origins[3 * b], j0 = origins[3 * b + 1], k0 = origins[3 * b + 2];
    if (z_ordered && std::min(std::max(i0, 0), d) > zlo) {
      const int c0 = std::min(std::max(i0, 0), d);
      cudaError_t ce = finish_planes(zlo, c0);
      if (ce != cudaSuccess) {
        release();
        return e->cuda_fail(ce, "finalise (streamed z range)");
      }
      zlo = c0;
    }
    IU_TILED_CUDA(launch_extract_block(vol_dev, d, h, w, i0, j0, k0, s, block, e->stream), "launch extract_block");
    e->launches += 1;
    for (int i = 0; i < n_axes; ++i)
      IU_TILED_TRY(iu_engine_predict_axis(e, block, IU_DTYPE_U8, s, axes[i], 0, s, p[axes[i]], 0, s, s, IU_FLAG_ASYNC));
    ReduceArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 3; ++i) a.p[i] = p[i];
    for (int i = 0; i < n_axes; ++i) a.order[i] = axes[i];
    a.n_axes = n_axes;
    a.n = s;
    a.t = s;
    a.z0 = 0;
    a.num_classes = c;
    a.g1d = g_dev;
    a.gmax = gmax;
    a.lo = lo;
    a.blend_pred = pred;
    a.blend_weight = weight;
    a.gd = d;
    a.gh = h;
    a.gw = w;
    a.gd_ring = zring;
    const int org[3] = {i0, j0, k0}, dims[3] = {d, h, w};
    for (int k = 0; k < 3; ++k) {
      a.b0[k] = org[k];
      a.l0[k] = org[k] < 0 ? -org[k] : 0;                                    // predict.py:401-404
      a.l1[k] = org[k] + s > dims[k] ? dims[k] - org[k] : s;
    }
    prof_begin(e, IU_PROF_REDUCE);
    cudaError_t ce = launch_reduce(a, e->stream);
    prof_end(e);
    e->launches += 1;
    IU_TILED_CUDA(ce, "launch reduce (blend)");
  }
  {
    cudaError_t ce = finish_planes(zlo, d);
    if (ce != cudaSuccess) {
      release();
      return e->cuda_fail(ce, "finalise");
    }
  }
#undef IU_TILED_TRY
#undef IU_TILED_CUDA
  (void)flags;  // host outputs and the scratch hand-back need the streams drained: always synchronous
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (to_host) {
    const cudaError_t ce2 = cudaStreamSynchronize(e->copy_stream);
    if (ce == cudaSuccess) ce = ce2;
  }
  drop_events();
  for (void* q : held) scratch_put(e, q);
  scratch_trim(e);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "predict_tiled");
  return IU_OK;
}

int iu_engine_blend_block(iu_engine* e, const float* p0, const float* p1, const float* p2, const int* order, int n_axes,
                          int s, int num_classes, const float* g1d_host, float gmax, float lo, float* pred_dev,
                          float* weight_dev, int d, int h, int w, const int* origin, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!order || n_axes < 1 || n_axes > 3 || s < 1 || num_classes < 1 || num_classes > 10 || !g1d_host || !pred_dev ||
      !weight_dev || !origin)
    return e->fail(IU_ERR_INVALID, "blend_block: bad arguments");
  ReduceArgs a;
  memset(&a, 0, sizeof(a));
  a.p[0] = p0;
  a.p[1] = p1;
  a.p[2] = p2;
  for (int i = 0; i < n_axes; ++i) {
    if (order[i] < 0 || order[i] > 2 || a.p[order[i]] == nullptr)
      return e->fail(IU_ERR_INVALID, "blend_block: axis in `order` has no probability buffer");
    a.order[i] = order[i];
  }
  a.n_axes = n_axes;
  a.n = s;
  a.t = s;
  a.num_classes = num_classes;
  a.gmax = gmax;
  a.lo = lo;
  a.blend_pred = pred_dev;
  a.blend_weight = weight_dev;
  a.gd = d;
  a.gh = h;
  a.gw = w;
  const int dims[3] = {d, h, w};
  for (int k = 0; k < 3; ++k) {
    a.b0[k] = origin[k];
    a.l0[k] = origin[k] < 0 ? -origin[k] : 0;
    a.l1[k] = origin[k] + s > dims[k] ? dims[k] - origin[k] : s;
  }
  float* g_dev = nullptr;
  if ((rc = scratch_get(e, (size_t)s * 4, (void**)&g_dev)) != IU_OK) return rc;
  cudaError_t ce = cudaMemcpyAsync(g_dev, g1d_host, (size_t)s * 4, cudaMemcpyHostToDevice, e->stream);
  a.g1d = g_dev;
  if (ce == cudaSuccess) ce = launch_reduce(a, e->stream);
  e->launches += 1;
  cudaStreamSynchronize(e->stream);  // the window table goes back to the pool
  scratch_put(e, g_dev);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch reduce (blend)");
  return finish(e, flags);
}

int iu_engine_finalise(iu_engine* e, const float* pred_dev, const float* weight_dev, int64_t voxels, int num_classes,
                       uint8_t* out_u8_dev, uint8_t* out_labels_dev, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!pred_dev || !weight_dev || voxels < 1 || num_classes < 1 || (!out_u8_dev && !out_labels_dev))
    return e->fail(IU_ERR_INVALID, "finalise: bad arguments");
  cudaError_t ce = launch_finalise(pred_dev, weight_dev, (size_t)voxels, num_classes, out_u8_dev, out_labels_dev, e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch finalise");
  e->launches += 1;
  return finish(e, flags);
}

int iu_engine_to_chunks(iu_engine* e, const void* volume_dev, int d, int h, int w, int elem, int chunk_elem, int cz,
                        int cy, int cx, void* staged_dev, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!volume_dev || !staged_dev) return e->fail(IU_ERR_INVALID, "to_chunks: null buffer");
  cudaError_t ce = launch_chunk_layout((const uint8_t*)volume_dev, (uint8_t*)staged_dev, d, h, w, elem, chunk_elem, cz, cy, cx, true,
                                       e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch to_chunks");
  e->launches += 1;
  return finish(e, flags);
}

int iu_engine_from_chunks(iu_engine* e, const void* staged_dev, int d, int h, int w, int elem, int chunk_elem, int cz,
                          int cy, int cx, void* volume_dev, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!volume_dev || !staged_dev) return e->fail(IU_ERR_INVALID, "from_chunks: null buffer");
  cudaError_t ce = launch_chunk_layout((const uint8_t*)staged_dev, (uint8_t*)volume_dev, d, h, w, elem, chunk_elem, cz,
                                       cy, cx, false, e->stream);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch from_chunks");
  e->launches += 1;
  return finish(e, flags);
}

int iu_engine_zoom_nearest(iu_engine* e, const void* src_dev, const int* src_dims, void* dst_dev, const int* dst_dims,
                           const int* t0, const int* t1, const int* t2, const int* t3, int item_bytes, unsigned flags) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!src_dev || !src_dims || !dst_dims || !t0 || !t1 || !t2 || !t3)
    return e->fail(IU_ERR_INVALID, "zoom_nearest: null argument");
  size_t n = 0, total = 1;
  for (int k = 0; k < 4; ++k) {
    if (src_dims[k] < 0 || dst_dims[k] < 0) return e->fail(IU_ERR_INVALID, "zoom_nearest: negative extent");
    n += (size_t)dst_dims[k];
    total *= (size_t)dst_dims[k];
  }
  if (total == 0) return IU_OK;  // an empty level (the reference halves the class axis too: C = 1 -> 0)
  if (!dst_dev) return e->fail(IU_ERR_INVALID, "zoom_nearest: null destination");
  if (item_bytes != 1 && item_bytes != 2 && item_bytes != 4 && item_bytes != 8)
    return e->fail(IU_ERR_INVALID, "zoom_nearest: item size must be 1, 2, 4 or 8 bytes");
  const int* tabs[4] = {t0, t1, t2, t3};
  std::vector<int> host(n);
  size_t off = 0;
  for (int k = 0; k < 4; ++k) {
    for (int i = 0; i < dst_dims[k]; ++i) {
      const int v = tabs[k][i];
      if (v < -1 || v >= src_dims[k]) return e->fail(IU_ERR_INVALID, "zoom_nearest: table entry outside the source");
      host[off + i] = v;
    }
    off += (size_t)dst_dims[k];
  }
  int* tab_dev = nullptr;
  if ((rc = scratch_get(e, n * sizeof(int), (void**)&tab_dev)) != IU_OK) return rc;
  cudaError_t ce = cudaMemcpyAsync(tab_dev, host.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream);
  const int* dz = tab_dev;
  const int* dy = dz + dst_dims[0];
  const int* dx = dy + dst_dims[1];
  const int* dc = dx + dst_dims[2];
  if (ce == cudaSuccess)
    ce = launch_zoom_gather((const uint8_t*)src_dev, src_dims, (uint8_t*)dst_dev, dst_dims, dz, dy, dx, dc, item_bytes,
                            e->stream);
  e->launches += 1;
  cudaStreamSynchronize(e->stream);  // the tables (pageable host copy, pooled device copy) are released here
  scratch_put(e, tab_dev);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "launch zoom_nearest");
  return finish(e, flags);
}

int iu_engine_conv_test(iu_engine* e, const void* src0, int cin0, const void* src1, int cin1, int batch, int h_in,
                        int w_in, int ksize, int stride, const float* weight, const float* bias, int cout,
                        const void* residual, int relu, int up2x, void* out) {
  int rc;
  DeviceGuard guard;
  if (!check_engine(e, false, &rc, &guard)) return rc;
  if (!src0 || !weight || !bias || !out || batch < 1 || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2) ||
      cin0 < 16 || cout % 16 || (src1 == nullptr) != (cin1 == 0) || batch % 8)
    return e->fail(IU_ERR_INVALID, "conv_test: bad arguments (batch must be a multiple of 8)");
  const int nseg = src1 ? 2 : 1;
  const int cin_total = cin0 + cin1;
  const int cin_min = src1 ? std::min(cin0, cin1) : cin0;
  int kc, bn;
  if (!pick_tile(cin_min, cout, &kc, &bn) || cin0 % kc || cin1 % kc || cout % bn)
    return e->fail(IU_ERR_INVALID, "conv_test: no kernel variant for these channel counts");
  const int ktot = ksize * ksize * cin_total;
  std::vector<uint16_t> packed((size_t)cout * ktot, 0);
  pack_segment(packed, e->fp16, ktot, 0, weight, cout, cin_total, 0, cin0, ksize);
  if (src1) pack_segment(packed, e->fp16, ktot, ksize * ksize * cin0, weight, cout, cin_total, cin0, cin1, ksize);
  void* d_w = nullptr;
  float* d_b = nullptr;
  if ((rc = scratch_get(e, packed.size() * 2, &d_w)) != IU_OK) return rc;
  if ((rc = scratch_get(e, (size_t)cout * 4, (void**)&d_b)) != IU_OK) {
    scratch_put(e, d_w);
    return rc;
  }
  cudaMemcpyAsync(d_w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, e->stream);
  cudaMemcpyAsync(d_b, bias, (size_t)cout * 4, cudaMemcpyHostToDevice, e->stream);
  const int pad = ksize / 2;
  const int src0_up = (up2x >> 1) & 1;  // bit 1: src0 is [batch][h_in/2][w_in/2][cin0], read through a 2x nearest upsample
  up2x &= 1;
  if (src0_up && (ksize != 3 || stride != 1 || (h_in | w_in) % 2)) {
    scratch_put(e, d_w);
    scratch_put(e, d_b);
    return e->fail(IU_ERR_INVALID, "conv_test: an upsampled source needs a 3x3 stride-1 conv and even h_in, w_in");
  }
  const int out_h = (h_in + 2 * pad - ksize) / stride + 1, out_w = (w_in + 2 * pad - ksize) / stride + 1;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  set_tiling(&a, batch, out_h, out_w);
  a.nseg = nseg;
  a.seg[0] = {cin0, ksize, stride, pad, src0_up};
  a.seg[1] = {cin1, ksize, stride, pad, 0};
  rc = src0_up ? IU_OK
               : encode_act_map(e, &a.amap[0], src0, cin0, w_in, h_in, batch, kc, a.tw * stride, a.th * stride, a.nb,
                                stride);
  if (rc == IU_OK && src1)
    rc = encode_act_map(e, &a.amap[1], src1, cin1, w_in, h_in, batch, kc, a.tw * stride, a.th * stride, a.nb, stride);
  if (rc == IU_OK) rc = encode_weight_map(e, &a.bmap, d_w, ktot, cout, kc, bn);
  if (rc == IU_OK && kc == 64 && cout % 128 == 0) rc = encode_weight_map(e, &a.bmap2, d_w, ktot, cout, 64, 64);
  if (rc == IU_OK && kc == 64 && cout % 256 == 0) {
    rc = encode_weight_map(e, &a.bmap256, d_w, ktot, cout, 64, 256);
    a.use_bn256 = rc == IU_OK;
  }
  void* d_wf = nullptr;
  void* d_wu = nullptr;
  if (rc == IU_OK) {
    a.cout = cout;
    a.bias = d_b;
    a.residual = (const __nv_bfloat16*)residual;
    a.out = out;
    a.relu = relu;
    a.fp16 = e->fp16;
    a.up2x = up2x;
    a.mode = kEpiBf16;
    a.src_ptr[0] = (const __nv_bfloat16*)src0;
    a.src_ptr[1] = (const __nv_bfloat16*)src1;
    const int row_mode = (e->conv_row && ksize == 3) ? conv_row_mode(a) : 0;
    if (row_mode && (rc = ensure_identity(e)) == IU_OK) {
      const int ktot_f = 3 * cin_total, kcr = conv_row_kc(cout, row_mode);
      std::vector<uint16_t> pf((size_t)3 * cout * ktot_f, 0);
      pack_fold_segment(pf, e->fp16, ktot_f, 0, weight, cout, cout, cin_total, 0, cin0);
      if (src1) pack_fold_segment(pf, e->fp16, ktot_f, 3 * cin0, weight, cout, cout, cin_total, cin0, cin1);
      if ((rc = scratch_get(e, pf.size() * 2, &d_wf)) == IU_OK) {
        cudaMemcpyAsync(d_wf, pf.data(), pf.size() * 2, cudaMemcpyHostToDevice, e->stream);
        cudaStreamSynchronize(e->stream);  // `pf` is a local
        rc = encode_weight_map(e, &a.bmapf, d_wf, ktot_f, 3 * cout, kcr, 3 * cout);
        if (rc == IU_OK && src0_up) {
          std::vector<uint16_t> pu((size_t)4 * cout * 3 * cin0, 0);
          pack_up_fold(pu, e->fp16, weight, cout, cout, cin_total, cin0);
          if ((rc = scratch_get(e, pu.size() * 2, &d_wu)) == IU_OK) {
            cudaMemcpyAsync(d_wu, pu.data(), pu.size() * 2, cudaMemcpyHostToDevice, e->stream);
            cudaStreamSynchronize(e->stream);  // `pu` is a local
            rc = encode_weight_map(e, &a.bmapu, d_wu, 3 * cin0, 4 * cout, kcr, 4 * cout);
          }
        }
        if (rc == IU_OK && residual) rc = encode_weight_map(e, &a.bmapi, e->d_ident, 64, 64, kcr, cout);
        if (rc == IU_OK)
          rc = encode_act_map(e, &a.omap, out, cout, out_w, out_h, batch, cout, 128, conv_row_store_rows(cout), 1, 1);
        if (rc == IU_OK && residual && cout == 64 && row_mode == 1 && e->row_res_tma) {
          rc = encode_act_map(e, &a.rmap, residual, cout, out_w, out_h, batch, cout, 128, conv_row_store_rows(cout), 1, 1);
          a.res_tma = rc == IU_OK ? e->row_res_tma : 0;
        }
        if (rc == IU_OK && e->row_tma && conv_row_tma_applicable(a)) {
          rc = encode_act_map(e, &a.rowmap, a.src_ptr[0], 64, out_w, out_h, batch, 64, 130, 3, 1, 1);
          a.row_tma = rc == IU_OK ? e->row_tma : 0;
        }
        if (rc == IU_OK) a.use_row = 1;
      }
    }
  }
  if (rc == IU_OK && e->halo_tma && kc == 64 && conv_halo_tma_applicable(a)) {
    // halo tiles by TMA (no minimum image / grid size here: the unit tests drive the edge cases through it)
    const int pitch = e->halo_tma == 24 ? 24 : kHaloTile + 2;
    rc = encode_act_map(e, &a.hmap[0], src0, cin0, out_w, out_h, batch, 64, pitch, kHaloTile + 2, 1, 1);
    if (rc == IU_OK && src1)
      rc = encode_act_map(e, &a.hmap[1], src1, cin1, out_w, out_h, batch, 64, pitch, kHaloTile + 2, 1, 1);
    a.halo_tma = rc == IU_OK ? pitch : 0;
  }
  if (rc == IU_OK) {
    cudaError_t ce = launch_conv(e, a, kc, bn);
    e->launches += 1;
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) rc = e->cuda_fail(ce, "conv_test");
  }
  cudaStreamSynchronize(e->stream);
  scratch_put(e, d_w);
  scratch_put(e, d_b);
  scratch_put(e, d_wf);
  scratch_put(e, d_wu);
  return rc;
}

}  // extern "C"
