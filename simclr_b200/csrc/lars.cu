// LARS multi-tensor update (tf2/lars_optimizer.py:83-137), HBM-bound:
// pass 1 reads w,g (8 B/param) for the norms, pass 2 reads w,g,v and writes w,v
// (20 B/param).  Reduction order is fixed (per-chunk partials, then a fixed
// tree per tensor), so all replicas apply bit-identical updates.
#include "common.cuh"

namespace simclr {
namespace {

constexpr int LT = 256;

struct LarsArgs {
  float* const* w;
  const float* const* g;
  float* const* v;
  const int64_t* numel;
  const int32_t* flags;
  const int32_t* chunk_tensor;
  const int64_t* chunk_offset;
  const int64_t* tensor_chunk_begin;
  int64_t chunk_elems;
  const float* lr;
  float momentum, weight_decay, eeta;
  float* partials;
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < (LT >> 5)) ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

__global__ void __launch_bounds__(LT) lars_norms_kernel(LarsArgs a) {
  __shared__ float sh[32];
  const int64_t chunk = blockIdx.x;
  const int t = a.chunk_tensor[chunk];
  const int32_t fl = a.flags[t];
  if (!(fl & 2)) return;                       // no layer adaptation: norms unused
  const int64_t off = a.chunk_offset[chunk];
  const int64_t n = min(a.chunk_elems, a.numel[t] - off);
  const float* __restrict__ w = a.w[t] + off;
  const float* __restrict__ g = a.g[t] + off;
  const float wd = (fl & 1) ? a.weight_decay : 0.f;
  float sw = 0.f, sg = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  const int64_t n4 = vec ? (n >> 2) : 0;
  for (int64_t i = threadIdx.x; i < n4; i += LT) {
    const float4 ww = reinterpret_cast<const float4*>(w)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float x;
    sw += ww.x * ww.x + ww.y * ww.y + ww.z * ww.z + ww.w * ww.w;
    x = fmaf(wd, ww.x, gg.x); sg = fmaf(x, x, sg);
    x = fmaf(wd, ww.y, gg.y); sg = fmaf(x, x, sg);
    x = fmaf(wd, ww.z, gg.z); sg = fmaf(x, x, sg);
    x = fmaf(wd, ww.w, gg.w); sg = fmaf(x, x, sg);
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += LT) {
    const float x = fmaf(wd, w[i], g[i]);
    sw = fmaf(w[i], w[i], sw);
    sg = fmaf(x, x, sg);
  }
  const float rw = block_sum(sw, sh);
  const float rg = block_sum(sg, sh);
  if (threadIdx.x == 0) { a.partials[chunk * 2] = rw; a.partials[chunk * 2 + 1] = rg; }
}

__global__ void __launch_bounds__(LT) lars_update_kernel(LarsArgs a) {
  __shared__ float s_trust;
  const int64_t chunk = blockIdx.x;
  const int t = a.chunk_tensor[chunk];
  const int32_t fl = a.flags[t];
  const float lr = *a.lr;
  if (threadIdx.x < 32) {
    float trust = 1.f;
    if (fl & 2) {
      // fixed-order reduction of this tensor's chunk partials (identical on every block / replica)
      const int64_t c0 = a.tensor_chunk_begin[t], c1 = a.tensor_chunk_begin[t + 1];
      double sw = 0.0, sg = 0.0;
      for (int64_t c = c0 + threadIdx.x; c < c1; c += 32) { sw += a.partials[c * 2]; sg += a.partials[c * 2 + 1]; }
      sw = warp_sum(sw); sg = warp_sum(sg);
      const float w_norm = sqrtf((float)sw), g_norm = sqrtf((float)sg);
      // tf2/lars_optimizer.py:103-107
      trust = (w_norm > 0.f) ? ((g_norm > 0.f) ? (a.eeta * w_norm / g_norm) : 1.f) : 1.f;
    }
    if (threadIdx.x == 0) s_trust = trust;
  }
  __syncthreads();
  const float scaled_lr = lr * s_trust;
  const float wd = (fl & 1) ? a.weight_decay : 0.f;
  const float mom = a.momentum;
  const int64_t off = a.chunk_offset[chunk];
  const int64_t n = min(a.chunk_elems, a.numel[t] - off);
  float* __restrict__ w = a.w[t] + off;
  const float* __restrict__ g = a.g[t] + off;
  float* __restrict__ v = a.v[t] + off;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? (n >> 2) : 0;
  for (int64_t i = threadIdx.x; i < n4; i += LT) {
    float4 ww = reinterpret_cast<float4*>(w)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    // next_v = momentum*v + scaled_lr*(g + wd*w);  w -= next_v   (:96-115)
    vv.x = fmaf(mom, vv.x, scaled_lr * fmaf(wd, ww.x, gg.x)); ww.x -= vv.x;
    vv.y = fmaf(mom, vv.y, scaled_lr * fmaf(wd, ww.y, gg.y)); ww.y -= vv.y;
    vv.z = fmaf(mom, vv.z, scaled_lr * fmaf(wd, ww.z, gg.z)); ww.z -= vv.z;
    vv.w = fmaf(mom, vv.w, scaled_lr * fmaf(wd, ww.w, gg.w)); ww.w -= vv.w;
    reinterpret_cast<float4*>(w)[i] = ww;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += LT) {
    const float nv = fmaf(mom, v[i], scaled_lr * fmaf(wd, w[i], g[i]));
    v[i] = nv;
    w[i] -= nv;
  }
}

}  // namespace
}  // namespace simclr

using namespace simclr;

extern "C" int simclr_lars_apply(int64_t n_tensors, int64_t n_chunks, const void* w_ptrs, const void* g_ptrs,
                                 const void* v_ptrs, const int64_t* numels, const int32_t* flags,
                                 const int32_t* chunk_tensor, const int64_t* chunk_offset,
                                 const int64_t* tensor_chunk_begin, int64_t chunk_elems, const float* lr,
                                 float momentum, float weight_decay, float eeta, float* partials,
                                 void* stream) {
  SIMCLR_CHECK_ARG(w_ptrs && g_ptrs && v_ptrs && numels && flags && chunk_tensor && chunk_offset &&
                       tensor_chunk_begin && lr && partials, "lars_apply: null pointer");
  SIMCLR_CHECK_ARG(n_tensors > 0 && n_chunks >= n_tensors, "lars_apply: bad table sizes");
  SIMCLR_CHECK_ARG(chunk_elems > 0 && chunk_elems % 4 == 0, "lars_apply: chunk_elems must be a positive multiple of 4");
  LarsArgs a;
  a.w = (float* const*)w_ptrs; a.g = (const float* const*)g_ptrs; a.v = (float* const*)v_ptrs;
  a.numel = numels; a.flags = flags; a.chunk_tensor = chunk_tensor; a.chunk_offset = chunk_offset;
  a.tensor_chunk_begin = tensor_chunk_begin; a.chunk_elems = chunk_elems; a.lr = lr;
  a.momentum = momentum; a.weight_decay = weight_decay; a.eeta = eeta; a.partials = partials;
  cudaStream_t st = (cudaStream_t)stream;
  lars_norms_kernel<<<(unsigned)n_chunks, LT, 0, st>>>(a);
  SIMCLR_CHECK_LAUNCH();
  lars_update_kernel<<<(unsigned)n_chunks, LT, 0, st>>>(a);
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}
