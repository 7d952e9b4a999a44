// The two other optimizers of `build_optimizer` (tf2/model.py:29-44): Keras SGD with Nesterov momentum and
// Keras Adam, as element-wise updates over the flat parameter / gradient / slot buffers (HBM-bound: SGD
// 16 B/param read + 8 written, Adam 16 + 12).  `hyper` is a device array staged by the host before the
// launch (CUDA-graph capturable): SGD {lr}; Adam {lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)}.
#include "common.cuh"

namespace simclr {
namespace {

// tf.keras.optimizers.SGD(lr, momentum, nesterov=True):  v <- m v - lr g;  w <- w + m v - lr g
// (nesterov=False: w <- w + v)
__global__ void sgd_momentum_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ v, int64_t n,
                                    const float* __restrict__ hyper, float momentum, int nesterov) {
  const float lr = hyper[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float vi = momentum * v[i] - lr * gi;
    v[i] = vi;
    w[i] += nesterov ? (momentum * vi - lr * gi) : vi;
  }
}

// tf.keras.optimizers.Adam(lr): m <- m + (g - m)(1 - b1); v <- v + (g^2 - v)(1 - b2); w <- w - lr_t m / (sqrt(v) + eps)
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, const float* __restrict__ hyper, float beta1, float beta2, float eps) {
  const float lr_t = hyper[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.f - beta2);
    m[i] = mi; v[i] = vi;
    w[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

inline unsigned grid_of(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace
}  // namespace simclr

using namespace simclr;

extern "C" {

int simclr_sgd_momentum_apply(float* w, const float* g, float* v, int64_t n, const float* hyper_dev, float momentum,
                              int nesterov, void* stream) {
  SIMCLR_CHECK_ARG(w && g && v && hyper_dev && n > 0, "sgd_momentum_apply: bad arguments");
  sgd_momentum_kernel<<<grid_of(n), 256, 0, (cudaStream_t)stream>>>(w, g, v, n, hyper_dev, momentum, nesterov);
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

int simclr_adam_apply(float* w, const float* g, float* m, float* v, int64_t n, const float* hyper_dev, float beta1,
                      float beta2, float eps, void* stream) {
  SIMCLR_CHECK_ARG(w && g && m && v && hyper_dev && n > 0, "adam_apply: bad arguments");
  adam_kernel<<<grid_of(n), 256, 0, (cudaStream_t)stream>>>(w, g, m, v, n, hyper_dev, beta1, beta2, eps);
  SIMCLR_CHECK_LAUNCH();
  return SIMCLR_OK;
}

}  // extern "C"
