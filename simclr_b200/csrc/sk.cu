// Selective-kernel block pieces (tf2/resnet.py:217-277 SK_Conv2D) and the ResNet-D
// shortcut pooling (tf2/resnet.py:333-340,401-408).  All HBM-bound
// elementwise / reduction kernels over NHWC tensors viewed as [N][HW][C].
#include "common.cuh"

namespace simclr {
namespace {

// y[n, c] = mean_hw(x[n, hw, c] + x[n, hw, f + c])     (tf2/resnet.py:265-266)
template <typename T>
__global__ void sk_pool_kernel(const T* __restrict__ x, float* __restrict__ g, int64_t N, int HW, int f) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * f) return;
  const int64_t n = idx / f; const int c = (int)(idx % f);
  const T* p = x + n * HW * (int64_t)(2 * f) + c;
  float s = 0.f;
  for (int i = 0; i < HW; ++i) s += to_f<T>(p[(int64_t)i * 2 * f]) + to_f<T>(p[(int64_t)i * 2 * f + f]);
  g[idx] = s / (float)HW;
}

// mixing = softmax over the two streams of logits [N][2f]; out = x0*m0 + x1*m1   (:270-275)
template <typename T>
__global__ void sk_mix_fwd_kernel(const T* __restrict__ x, const float* __restrict__ logits, float* __restrict__ mixing,
                                  T* __restrict__ out, int64_t total, int HW, int f) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % f);
    const int64_t nhw = idx / f;
    const int64_t n = nhw / HW;
    const float l0 = logits[n * 2 * f + c], l1 = logits[n * 2 * f + f + c];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float m0 = e0 / (e0 + e1), m1 = e1 / (e0 + e1);
    if (nhw % HW == 0) { mixing[n * 2 * f + c] = m0; mixing[n * 2 * f + f + c] = m1; }
    const float x0 = to_f<T>(x[nhw * 2 * f + c]), x1 = to_f<T>(x[nhw * 2 * f + f + c]);
    out[idx] = from_f<T>(x0 * m0 + x1 * m1);
  }
}

// dmix_s[n,c] = sum_hw dout*x_s ; softmax backward -> dlogits [N][2f]
template <typename T>
__global__ void sk_mix_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                                         const float* __restrict__ mixing, float* __restrict__ dlogits, int64_t N,
                                         int HW, int f) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * f) return;
  const int64_t n = idx / f; const int c = (int)(idx % f);
  float d0 = 0.f, d1 = 0.f;
  for (int i = 0; i < HW; ++i) {
    const int64_t nhw = n * HW + i;
    const float g = to_f<T>(dout[nhw * f + c]);
    d0 = fmaf(g, to_f<T>(x[nhw * 2 * f + c]), d0);
    d1 = fmaf(g, to_f<T>(x[nhw * 2 * f + f + c]), d1);
  }
  const float m0 = mixing[n * 2 * f + c], m1 = mixing[n * 2 * f + f + c];
  const float dot = m0 * d0 + m1 * d1;
  dlogits[n * 2 * f + c] = m0 * (d0 - dot);
  dlogits[n * 2 * f + f + c] = m1 * (d1 - dot);
}

// dx[n,hw,s*f+c] = dout[n,hw,c]*m_s[n,c] + dg[n,c]/HW   (dg: gradient of the pooled features)
template <typename T>
__global__ void sk_mix_bwd_apply_kernel(const T* __restrict__ dout, const float* __restrict__ mixing,
                                        const float* __restrict__ dg, T* __restrict__ dx, int64_t total, int HW, int f) {
  const float inv = 1.f / (float)HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % f);
    const int64_t nhw = idx / f;
    const int64_t n = nhw / HW;
    const float g = to_f<T>(dout[idx]);
    const float pool = dg[n * f + c] * inv;
    dx[nhw * 2 * f + c] = from_f<T>(fmaf(g, mixing[n * 2 * f + c], pool));
    dx[nhw * 2 * f + f + c] = from_f<T>(fmaf(g, mixing[n * 2 * f + f + c], pool));
  }
}

// AveragePooling2D(2, stride): stride 2 -> FixedPadding(2) (zero row/col after) + 'VALID' (divisor 4);
// stride 1 -> 'SAME' (pad after, divisor counts valid elements only).
template <typename T, bool BWD>
__global__ void avgpool2x2_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t total, int H, int W, int C,
                                  int Ho, int Wo, int stride) {
  // FWD: one thread per output element.  BWD: one thread per input-gradient element (gather form).
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    int64_t p = idx / C;
    if (!BWD) {
      const int wo = (int)(p % Wo); p /= Wo;
      const int ho = (int)(p % Ho);
      const int64_t n = p / Ho;
      float s = 0.f; int cnt = 0;
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
          const int h = ho * stride + dh, w = wo * stride + dw;
          if (h < H && w < W) { s += to_f<T>(in[((n * H + h) * W + w) * (int64_t)C + c]); ++cnt; }
        }
      out[idx] = from_f<T>(s / (float)(stride == 1 ? cnt : 4));
    } else {
      const int w = (int)(p % W); p /= W;
      const int h = (int)(p % H);
      const int64_t n = p / H;
      float s = 0.f;
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
          const int th = h - dh, tw = w - dw;
          if (th < 0 || tw < 0 || th % stride || tw % stride) continue;
          const int ho = th / stride, wo = tw / stride;
          if (ho >= Ho || wo >= Wo) continue;
          float div = 4.f;
          if (stride == 1) div = (float)((ho + 1 < H ? 2 : 1) * (wo + 1 < W ? 2 : 1));
          s += to_f<T>(in[((n * Ho + ho) * Wo + wo) * (int64_t)C + c]) / div;
        }
      out[idx] = from_f<T>(s);
    }
  }
}

// ---- squeeze-and-excitation (tf2/resnet.py:280-311) ------------------------------------------
// out = sigmoid(l[n,c]) * x[n,hw,c]
template <typename T>
__global__ void se_scale_fwd_kernel(const T* __restrict__ x, const float* __restrict__ logits, T* __restrict__ out,
                                    int64_t total, int HW, int C) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t n = idx / ((int64_t)HW * C);
    const float g = 1.f / (1.f + expf(-logits[n * C + c]));
    out[idx] = from_f<T>(to_f<T>(x[idx]) * g);
  }
}
// dl[n,c] = sigma'(l) * sum_hw dout*x
template <typename T>
__global__ void se_scale_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                                           const float* __restrict__ logits, float* __restrict__ dlogits, int64_t N,
                                           int HW, int C) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * C) return;
  const int64_t n = idx / C; const int c = (int)(idx % C);
  float d = 0.f;
  for (int i = 0; i < HW; ++i) {
    const int64_t o = (n * HW + i) * (int64_t)C + c;
    d = fmaf(to_f<T>(dout[o]), to_f<T>(x[o]), d);
  }
  const float g = 1.f / (1.f + expf(-logits[idx]));
  dlogits[idx] = d * g * (1.f - g);
}
// dx = dout*sigmoid(l) + dmean[n,c]/HW   (dmean: gradient w.r.t. the squeezed mean)
template <typename T>
__global__ void se_scale_bwd_apply_kernel(const T* __restrict__ dout, const float* __restrict__ logits,
                                          const float* __restrict__ dmean, T* __restrict__ dx, int64_t total, int HW,
                                          int C) {
  const float inv = 1.f / (float)HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t n = idx / ((int64_t)HW * C);
    const float g = 1.f / (1.f + expf(-logits[n * C + c]));
    dx[idx] = from_f<T>(fmaf(to_f<T>(dout[idx]), g, dmean[n * C + c] * inv));
  }
}
// y = max(x, 0) in place / dx = dy * [y > 0] in place, fp32 [n]
__global__ void relu_kernel(float* __restrict__ x, const float* __restrict__ mask_src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mask_src == nullptr) x[i] = fmaxf(x[i], 0.f);
    else x[i] = mask_src[i] > 0.f ? x[i] : 0.f;
  }
}

inline unsigned grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace
}  // namespace simclr

using namespace simclr;
typedef __nv_bfloat16 bf16;

extern "C" {

int simclr_sk_pool(const void* x, int dtype, float* g, int64_t N, int64_t HW, int64_t f, void* stream) {
  SIMCLR_CHECK_ARG(x && g && N > 0 && HW > 0 && f > 0, "sk_pool: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((N * f + 255) / 256);
  if (dtype == SIMCLR_F32) sk_pool_kernel<float><<<grid, 256, 0, st>>>((const float*)x, g, N, (int)HW, (int)f);
  else if (dtype == SIMCLR_BF16) sk_pool_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, g, N, (int)HW, (int)f);
  else { set_error("sk_pool: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_sk_mix_fwd(const void* x, const float* logits, float* mixing, void* out, int dtype, int64_t N, int64_t HW,
                      int64_t f, void* stream) {
  SIMCLR_CHECK_ARG(x && logits && mixing && out && N > 0 && HW > 0 && f > 0, "sk_mix_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * HW * f;
  if (dtype == SIMCLR_F32) sk_mix_fwd_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)x, logits, mixing, (float*)out, total, (int)HW, (int)f);
  else if (dtype == SIMCLR_BF16) sk_mix_fwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)x, logits, mixing, (bf16*)out, total, (int)HW, (int)f);
  else { set_error("sk_mix_fwd: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_sk_mix_bwd_reduce(const void* dout, const void* x, const float* mixing, float* dlogits, int dtype, int64_t N,
                             int64_t HW, int64_t f, void* stream) {
  SIMCLR_CHECK_ARG(dout && x && mixing && dlogits && N > 0 && HW > 0 && f > 0, "sk_mix_bwd_reduce: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((N * f + 255) / 256);
  if (dtype == SIMCLR_F32) sk_mix_bwd_reduce_kernel<float><<<grid, 256, 0, st>>>((const float*)dout, (const float*)x, mixing, dlogits, N, (int)HW, (int)f);
  else if (dtype == SIMCLR_BF16) sk_mix_bwd_reduce_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dout, (const bf16*)x, mixing, dlogits, N, (int)HW, (int)f);
  else { set_error("sk_mix_bwd_reduce: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_sk_mix_bwd_apply(const void* dout, const float* mixing, const float* dg, void* dx, int dtype, int64_t N,
                            int64_t HW, int64_t f, void* stream) {
  SIMCLR_CHECK_ARG(dout && mixing && dg && dx && N > 0 && HW > 0 && f > 0, "sk_mix_bwd_apply: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * HW * f;
  if (dtype == SIMCLR_F32) sk_mix_bwd_apply_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)dout, mixing, dg, (float*)dx, total, (int)HW, (int)f);
  else if (dtype == SIMCLR_BF16) sk_mix_bwd_apply_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)dout, mixing, dg, (bf16*)dx, total, (int)HW, (int)f);
  else { set_error("sk_mix_bwd_apply: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_se_scale_fwd(const void* x, const float* logits, void* out, int dtype, int64_t N, int64_t HW, int64_t C,
                        void* stream) {
  SIMCLR_CHECK_ARG(x && logits && out && N > 0 && HW > 0 && C > 0, "se_scale_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * HW * C;
  if (dtype == SIMCLR_F32) se_scale_fwd_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)x, logits, (float*)out, total, (int)HW, (int)C);
  else if (dtype == SIMCLR_BF16) se_scale_fwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)x, logits, (bf16*)out, total, (int)HW, (int)C);
  else { set_error("se_scale_fwd: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_se_scale_bwd_reduce(const void* dout, const void* x, const float* logits, float* dlogits, int dtype,
                               int64_t N, int64_t HW, int64_t C, void* stream) {
  SIMCLR_CHECK_ARG(dout && x && logits && dlogits && N > 0 && HW > 0 && C > 0, "se_scale_bwd_reduce: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((N * C + 255) / 256);
  if (dtype == SIMCLR_F32) se_scale_bwd_reduce_kernel<float><<<grid, 256, 0, st>>>((const float*)dout, (const float*)x, logits, dlogits, N, (int)HW, (int)C);
  else if (dtype == SIMCLR_BF16) se_scale_bwd_reduce_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dout, (const bf16*)x, logits, dlogits, N, (int)HW, (int)C);
  else { set_error("se_scale_bwd_reduce: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_se_scale_bwd_apply(const void* dout, const float* logits, const float* dmean, void* dx, int dtype,
                              int64_t N, int64_t HW, int64_t C, void* stream) {
  SIMCLR_CHECK_ARG(dout && logits && dmean && dx && N > 0 && HW > 0 && C > 0, "se_scale_bwd_apply: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * HW * C;
  if (dtype == SIMCLR_F32) se_scale_bwd_apply_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)dout, logits, dmean, (float*)dx, total, (int)HW, (int)C);
  else if (dtype == SIMCLR_BF16) se_scale_bwd_apply_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)dout, logits, dmean, (bf16*)dx, total, (int)HW, (int)C);
  else { set_error("se_scale_bwd_apply: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_relu_inplace(float* x, const float* mask_src, int64_t n, void* stream) {
  SIMCLR_CHECK_ARG(x && n > 0, "relu_inplace: bad args");
  relu_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, mask_src, n);
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_avgpool2x2_fwd(const void* x, void* y, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, int64_t stride,
                          void* stream) {
  SIMCLR_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2), "avgpool2x2_fwd: bad args");
  const int64_t Ho = stride == 1 ? H : (H - 1) / 2 + 1, Wo = stride == 1 ? W : (W - 1) / 2 + 1;
  const int64_t total = N * Ho * Wo * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SIMCLR_F32) avgpool2x2_kernel<float, false><<<grid_for(total), 256, 0, st>>>((const float*)x, (float*)y, total, (int)H, (int)W, (int)C, (int)Ho, (int)Wo, (int)stride);
  else if (dtype == SIMCLR_BF16) avgpool2x2_kernel<bf16, false><<<grid_for(total), 256, 0, st>>>((const bf16*)x, (bf16*)y, total, (int)H, (int)W, (int)C, (int)Ho, (int)Wo, (int)stride);
  else { set_error("avgpool2x2_fwd: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_avgpool2x2_bwd(const void* dy, void* dx, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, int64_t stride,
                          void* stream) {
  SIMCLR_CHECK_ARG(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2), "avgpool2x2_bwd: bad args");
  const int64_t Ho = stride == 1 ? H : (H - 1) / 2 + 1, Wo = stride == 1 ? W : (W - 1) / 2 + 1;
  const int64_t total = N * H * W * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SIMCLR_F32) avgpool2x2_kernel<float, true><<<grid_for(total), 256, 0, st>>>((const float*)dy, (float*)dx, total, (int)H, (int)W, (int)C, (int)Ho, (int)Wo, (int)stride);
  else if (dtype == SIMCLR_BF16) avgpool2x2_kernel<bf16, true><<<grid_for(total), 256, 0, st>>>((const bf16*)dy, (bf16*)dx, total, (int)H, (int)W, (int)C, (int)Ho, (int)Wo, (int)stride);
  else { set_error("avgpool2x2_bwd: unknown dtype"); return SIMCLR_ERR_INVALID_ARG; }
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

}  // extern "C"
