// Probe for the "halo reuse" plan of DESIGN.md section 9 (not part of the product): can a K-major,
// 128B-swizzled UMMA operand descriptor start at an arbitrary 128-byte row of a larger smem tile?
//
// A 3x3 stride-1 conv reads nine shifted windows of the same input pixels.  If the input halo is
// loaded ONCE as [pixels][128 B] rows (TMA, SWIZZLE_128B), tap (r, s) is the same tile seen from
// row offset r*Wp + s.  That works only if tcgen05.mma applies the 128B swizzle from the absolute
// shared-memory address bits (or honours the descriptor's base_offset field) when the start address
// is not a multiple of 1024 bytes.  This program loads a [192][64] bf16 matrix whose elements encode
// their own (row, column), runs D = A_window * I (identity B, K = 64) for several row shifts and both
// settings of base_offset, and prints which source rows/columns arrived in D.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -I simclr_b200/csrc -o gpurun_out/probe_umma_shift scripts/probe_umma_shift.cu
// Run  :  gpurun -- './gpurun_out/probe_umma_shift'
#include "tc_common.cuh"
#include <vector>

namespace simclr { void set_error(const char*, ...) {} }   // the probe does not link the library
using namespace simclr::tc;

constexpr int ROWS = 192, COLS = 64, M = 128, N = 64;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int shift, int use_base_offset,
      float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tile = smem;                       // [192][128 B]
  uint8_t* b_tile = smem + ROWS * 128;          // [64][128 B]   (24576 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + N * 128);
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
    mbar_arrive_expect_tx(bar, ROWS * 128 + N * 128);
    tma_load_2d(a_tile, &tm_a, bar, 0, 0);
    tma_load_2d(b_tile, &tm_b, bar, 0, 0);
    mbar_wait(bar, 0, 1);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(a_tile) + (uint32_t)shift * 128u;
    const uint32_t b_addr = smem_u32(b_tile);
    constexpr uint32_t IDESC = make_idesc(false, M, N, false, false);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = smem_desc_sw128(a_addr + k * 32, 16, 1024);
      if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;     // matrix base offset, bits [49,52)
      const uint64_t bd = smem_desc_sw128(b_addr + k * 32, 16, 1024);
      umma<false>(tmem, ad, bd, IDESC, k != 0 ? 1u : 0u);
    }
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0, 2);
  tc_fence_after();
  // D row m lives in TMEM lane m: warp w reads lanes 32w .. 32w+31
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
  // A[r][c] = r for c < 32, c for c >= 32 (exact in bf16);  B = identity (so D[m][n] = A_window[m][n])
  std::vector<__nv_bfloat16> ha(ROWS * COLS), hb(N * COLS);
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < COLS; ++c) ha[r * COLS + c] = __float2bfloat16(c < 32 ? (float)r : (float)c);
  for (int n = 0; n < N; ++n) for (int k = 0; k < COLS; ++k) hb[n * COLS + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, da, 2, ROWS, COLS, COLS * 2, ROWS, COLS) || make_tmap_2d(&tb, db, 2, N, COLS, COLS * 2, N, COLS)) {
    printf("tensor map encode failed\n"); return 1;
  }
  const size_t smem = 1024 + ROWS * 128 + N * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> ho(M * N);
  const int shifts[] = {0, 8, 1, 3, 7, 9, 58};
  for (int use_bo = 0; use_bo < 2; ++use_bo) {
    for (int shift : shifts) {
      cudaMemset(dout, 0xff, M * N * 4);
      probe<<<1, 128, smem>>>(ta, tb, shift, use_bo, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: kernel error %s\n", shift, use_bo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(ho.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
      int bad_rows = 0, bad_cols = 0;
      for (int m = 0; m < M; ++m) {
        for (int n = 0; n < 32; ++n) if (ho[m * N + n] != (float)(shift + m)) { ++bad_rows; break; }
        for (int n = 32; n < N; ++n) if (ho[m * N + n] != (float)n) { ++bad_cols; break; }
      }
      printf("shift %2d base_offset=%d : rows wrong %3d / 128, column order wrong %3d / 128 | D[0..3][0]=%g %g %g %g  D[0][32..35]=%g %g %g %g\n",
             shift, use_bo, bad_rows, bad_cols, ho[0], ho[N], ho[2 * N], ho[3 * N], ho[32], ho[33], ho[34], ho[35]);
    }
  }
  return 0;
}
