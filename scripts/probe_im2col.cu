// Probe of TMA im2col semantics on sm_100a (no public docs in this sandbox): loads pixel columns from
// an NHWC fp32 tensor whose elements encode their own (n, h, w, c), and prints what landed in smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/probe_im2col scripts/probe_im2col.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int PIX = 16, CPP = 32;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int ow, int oh, float* out) {
  __shared__ __align__(128) float tile[PIX * CPP];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
  for (int i = threadIdx.x; i < PIX * CPP; i += blockDim.x) tile[i] = -7.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(PIX * CPP * 4) : "memory");
    const uint16_t o_w = (uint16_t)ow, o_h = (uint16_t)oh;
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
        ::"r"(dst), "l"(&tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(bar_a), "h"(o_w), "h"(o_h) : "memory");
  }
  // bounded wait
  uint32_t ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PIX * CPP; i += blockDim.x) out[i] = tile[i];
  if (threadIdx.x == 0) out[PIX * CPP] = ok ? 1.f : 0.f;
}

int main() {
  const int N = 3, H = 6, W = 8, C = 64;
  std::vector<float> hx((size_t)N * H * W * C);
  for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c)
    hx[(((size_t)n * H + h) * W + w) * C + c] = (float)((((n + 1) * 16 + h) * 16 + w) * 256 + c);
  float* dx; cudaMalloc(&dx, hx.size() * 4); cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  float* dout; cudaMalloc(&dout, (PIX * CPP + 1) * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qr);
  if (!fn) { printf("no cuTensorMapEncodeIm2col\n"); return 1; }
  EncodeIm2col enc = (EncodeIm2col)fn;

  struct Case { int stride, lo, up; int c, w, h, n, ow, oh; const char* what; };
  const Case cases[] = {
      {1, -1, -1, 0, -1, -1, 0, 0, 0, "s1 pad1 3x3: base (-1,-1) off (0,0)"},
      {1, -1, -1, 0, -1, -1, 0, 1, 1, "s1 pad1 3x3: base (-1,-1) off (1,1)"},
      {1, -1, -1, 0, -1, -1, 0, 2, 2, "s1 pad1 3x3: base (-1,-1) off (2,2)"},
      {1, -1, -1, 32, 3, 4, 0, 2, 0, "s1: c=32 base w=3 h=4 off w=2 h=0 (wraps rows, crosses image)"},
      {1, -1, -1, 0, 0, 0, 0, 0, 0, "s1: base (0,0) off 0"},
      {1, -1, -1, 0, 2, 4, 2, 1, 1, "s1: last image base w=2 h=4 n=2 (runs past N)"},
      {2, -1, -1, 0, -1, -1, 0, 0, 0, "s2 pad1 3x3: base (-1,-1) off (0,0)"},
      {2, -1, -1, 0, -1, -1, 0, 1, 2, "s2 pad1 3x3: base (-1,-1) off w=1 h=2"},
      {2, -1, -1, 0, 3, 1, 1, 1, 1, "s2: base w=3 h=1 n=1 off (1,1)"},
      {1, 0, 0, 0, 0, 0, 0, 0, 0, "s1 1x1 no pad: base 0"},
      {1, -2, 0, 0, -2, -2, 0, 1, 2, "dgrad-like: lower -2 upper 0, base (-2,-2), off w=1 h=2"},
  };
  for (const Case& k : cases) {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lo[2] = {k.lo, k.lo}, up[2] = {k.up, k.up};
    cuuint32_t es[4] = {1, (cuuint32_t)k.stride, (cuuint32_t)k.stride, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, gdim, gstr, lo, up, CPP, PIX, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== %s  (encode rc=%d)\n", k.what, (int)r);
    if (r != CUDA_SUCCESS) continue;
    probe<<<1, 128>>>(tm, k.c, k.w, k.h, k.n, k.ow, k.oh, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("   kernel error %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> ho(PIX * CPP + 1);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    printf("   completed=%d :", (int)ho[PIX * CPP]);
    for (int p = 0; p < PIX; ++p) {
      const float v = ho[p * CPP];
      if (v == 0.f) printf(" [0]");
      else if (v == -7.f) printf(" [-]");
      else { const int iv = (int)v; const int c = iv & 255, w = (iv >> 8) & 15, h = (iv >> 12) & 15, n = (iv >> 16) - 1;
             printf(" n%dh%dw%dc%d", n, h, w, c); }
    }
    printf("\n");
  }
  return 0;
}
