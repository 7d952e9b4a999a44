// Probe for the stem kernel (not part of the product): a K-major, NON-swizzled UMMA operand whose core matrices
// OVERLAP in shared memory.
//
// The 7x7/2 stem reads, for output pixel q and filter row r, the 8 input pixels 2q-4 .. 2q+3 of one image row
// (4 stored channels: 64 contiguous bytes = four 16-byte "pixel pairs").  Neighbouring output pixels start one pair
// (16 bytes) apart.  In the canonical no-swizzle K-major layout a core matrix is 8 rows x 16 bytes with a 16-byte
// row pitch, the next core matrix along K sits LBO bytes further, the next 8 rows SBO bytes further.  With
// LBO = 16 and SBO = 128 the operand "row m, K chunk c" is read from start + 16*(m + c): exactly the slab of input
// pixels, no im2col copy.  This program checks that tcgen05.mma accepts such a descriptor (LBO smaller than a core
// matrix, start address any multiple of 16 bytes) and reads what the formula says.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -I simclr_b200/csrc -o gpurun_out/probe_umma_nosw_overlap scripts/probe_umma_nosw_overlap.cu -lcuda
#include "tc_common.cuh"
#include <vector>

namespace simclr { void set_error(const char*, ...) {} }
using namespace simclr::tc;

constexpr int M = 128, N = 64, PAIRS = 512;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm_b, int start_bytes, int kstep, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* slab = smem;                         // [PAIRS][16 B]
  uint8_t* b_tile = smem + PAIRS * 16;          // [64][128 B] SWIZZLE_128B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + N * 128);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < PAIRS * 8; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(slab)[i] = __float2bfloat16((float)(i % 256));
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, N * 128);
    tma_load_2d(b_tile, &tm_b, bar, 0, 0);
    mbar_wait(bar, 0, 1);
    tc_fence_after();
    constexpr uint32_t IDESC = make_idesc(false, M, N, false, false);
    constexpr uint32_t HI_NONE = ((128u >> 4) & 0x3FFFu) | (1u << 14);              // SBO = 128 B, layout_type 0
    const uint64_t ad = desc_pack(desc_lo(smem_u32(slab) + (uint32_t)start_bytes, 16), HI_NONE);   // LBO = 16 B
    const uint64_t bd = desc_pack(desc_lo(smem_u32(b_tile) + (uint32_t)kstep * 32u, 16), desc_hi_sw128(1024));
    umma<false>(tmem, ad, bd, IDESC, 0u);
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
  // B [64 n][64 k]: B[n][k] = 1 if n == k (so with K step `kstep`, D[m][n] = A[m][n - 16*kstep] for those 16 n)
  std::vector<__nv_bfloat16> hb(N * 64);
  for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16* db; float* dout;
  cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tb;
  if (make_tmap_2d(&tb, db, 2, N, 64, 64 * 2, N, 64)) { printf("tensor map encode failed\n"); return 1; }
  const size_t smem = 1024 + PAIRS * 16 + N * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> ho(M * N);
  const int starts[] = {0, 16, 32, 48, 1856, 1856 + 32, 2 * 1856 + 16};
  int total_bad = 0;
  for (int kstep = 0; kstep < 2; ++kstep) {
    for (int start : starts) {
      cudaMemset(dout, 0xff, M * N * 4);
      probe<<<1, 128, smem>>>(tb, start, kstep, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("start %d: kernel error %s\n", start, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(ho.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          const int k = n - 16 * kstep;            // A column selected by the identity block of this K step
          const float want = (k >= 0 && k < 16) ? (float)((start / 2 + 8 * m + k) % 256) : 0.f;   // element index of A[m][k]
          if (ho[m * N + n] != want) ++bad;
        }
      total_bad += bad;
      printf("B k-step %d, A start +%5d B : %5d / %d wrong | D[0][%d..]=%g %g ... D[1][%d]=%g D[127][%d]=%g\n", kstep, start, bad, M * N,
             16 * kstep, ho[16 * kstep], ho[16 * kstep + 1], 16 * kstep, ho[N + 16 * kstep], 16 * kstep + 15, ho[127 * N + 16 * kstep + 15]);
    }
  }
  printf(total_bad == 0 ? "PASS: overlapping no-swizzle K-major core matrices (LBO 16 B, SBO 128 B) read start + 16*(m + chunk)\n" : "FAIL\n");
  return total_bad != 0;
}
