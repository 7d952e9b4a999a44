// Probe for a halo-reuse WGRAD (not part of the product): the activation operand of wgrad is MN-major
// ([pixel][64 channels] rows of 128 B; the reduction K runs over pixels).  If the input halo is loaded once as a
// slab, tap (r, s) is the slab seen from pixel offset d = r*Wp + s, i.e. a shift along K by d rows of 128 B -- and
// TWO taps can fill one M = 128 operand if the second 64-channel atom may sit LBO = (d_b - d_a) * 128 bytes after
// the first (atoms OVERLAP in memory).  Questions: (1) does an MN-major SWIZZLE_128B descriptor accept a start
// address that is not 1024-byte aligned, (2) may LBO be any multiple of 128 bytes, smaller than an atom?
//
// X[p][c] = p for c < 32, c for c >= 32 (exact in bf16); B[k][n] = (k == n): D[m][n] = A[n][m].
// Expected D[m][n] = X[n + d(m / 64)][m % 64].
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -I simclr_b200/csrc -o build_tmp/probe_umma_mn_shift scripts/probe_umma_mn_shift.cu
#include "tc_common.cuh"
#include <vector>

namespace simclr { void set_error(const char*, ...) {} bool accumulate_prezeroed() { return false; } }
using namespace simclr::tc;

constexpr int ROWS = 256, COLS = 64, M = 128, N = 64, KPIX = 64;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int da, int db, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tile = smem;                       // [256 px][128 B]
  uint8_t* b_tile = smem + ROWS * 128;          // [64 px][64 n] MN-major: 64 rows of 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + KPIX * 128);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, ROWS * 128 + KPIX * 128);
    tma_load_2d(a_tile, &tm_a, bar, 0, 0);
    tma_load_2d(b_tile, &tm_b, bar, 0, 0);
    mbar_wait(bar, 0, 1);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(a_tile) + (uint32_t)da * 128u;
    const uint32_t lbo = (uint32_t)(db - da) * 128u;
    const uint32_t b_addr = smem_u32(b_tile);
    constexpr uint32_t IDESC = make_idesc(false, M, N, true, true);
#pragma unroll
    for (int k = 0; k < 4; ++k) {       // 16 pixels per MMA
      const uint64_t ad = smem_desc_sw128(a_addr + k * 16 * 128, lbo, 1024);
      const uint64_t bd = smem_desc_sw128(b_addr + k * 16 * 128, KPIX * 128, 1024);
      umma<false>(tmem, ad, bd, IDESC, k != 0 ? 1u : 0u);
    }
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0, 2);
  tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t acc[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, acc);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c * 32 + i] = __uint_as_float(acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  std::vector<__nv_bfloat16> ha(ROWS * COLS), hb(KPIX * N);
  for (int p = 0; p < ROWS; ++p) for (int c = 0; c < COLS; ++c) ha[p * COLS + c] = __float2bfloat16(c < 32 ? (float)p : (float)c);
  for (int k = 0; k < KPIX; ++k) for (int n = 0; n < N; ++n) hb[k * N + n] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *da_, *db_; float* dout;
  cudaMalloc(&da_, ha.size() * 2); cudaMalloc(&db_, hb.size() * 2); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(da_, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db_, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, da_, 2, ROWS, COLS, COLS * 2, ROWS, COLS) || make_tmap_2d(&tb, db_, 2, KPIX, N, N * 2, KPIX, N)) {
    printf("tensor map encode failed\n"); return 1;
  }
  const size_t smem = 1024 + ROWS * 128 + KPIX * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> ho(M * N);
  const int cases[][2] = {{0, 64}, {0, 8}, {8, 16}, {0, 1}, {1, 2}, {3, 4}, {0, 58}, {1, 59}, {58, 59}, {59, 60}, {116, 117}, {118, 118}};
  for (auto& cs : cases) {
    const int da = cs[0], db = cs[1];
    cudaMemset(dout, 0xff, M * N * 4);
    probe<<<1, 128, smem>>>(ta, tb, da, db, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift a=%d b=%d: kernel error %s\n", da, db, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(ho.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
    int bad0 = 0, bad1 = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        const float want = (m % 64) < 32 ? (float)(n + (m < 64 ? da : db)) : (float)(m % 64);
        if (ho[m * N + n] != want) { if (m < 64) ++bad0; else ++bad1; }
      }
    printf("tap shifts a=%3d b=%3d (LBO %5d B): atom 0 wrong %4d / 4096, atom 1 wrong %4d / 4096 | D[0][0..2]=%g %g %g  D[64][0..2]=%g %g %g  D[65][0]=%g\n",
           da, db, (db - da) * 128, bad0, bad1, ho[0], ho[1], ho[2], ho[64 * N], ho[64 * N + 1], ho[64 * N + 2], ho[65 * N]);
  }
  return 0;
}
