// Test-only probe kernels (NOT part of the product library): pin down tcgen05 operand layouts on
// real hardware before the attention kernel relies on them.
//   mode 0: SS MMA, A and B K-major tiles loaded by TMA with swizzle `sw` (32/64/128 bytes)
//   mode 1: TS MMA, A written to TMEM by tcgen05.st as packed bf16 pairs (row = lane), B as mode 0
// D[128 x N] fp32 is written to global row-major.
#include "../../llmseg_b200/csrc/common.cuh"
using namespace llmseg;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const bf16* __restrict__ A, float* __restrict__ D, int N, int KA, int sw, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 128 rows * sw bytes
  uint8_t* sB = smem + 128 * 128;      // N rows * sw bytes (<= 256*128)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 128 + 256 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = *tptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (mode == 0 ? 128 * sw : 0) + N * sw);
    if (mode == 0) tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  const uint32_t acol = 256;  // A operand columns in TMEM (mode 1)
  if (mode == 1) {
    uint32_t r[32];
    const int row = threadIdx.x;
    for (int c = 0; c < 32; ++c) {
      if (c < KA / 2) {
        uint32_t lo = reinterpret_cast<const uint16_t*>(A)[row * KA + 2 * c];
        uint32_t hi = reinterpret_cast<const uint16_t*>(A)[row * KA + 2 * c + 1];
        r[c] = lo | (hi << 16);
      } else r[c] = 0;
    }
    tmem_st32(tbase + (uint32_t(warp * 32) << 16) + acol, r);
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t lt = sw == 128 ? UMMA_SW128 : (sw == 64 ? UMMA_SW64 : UMMA_SW32);
    const uint32_t sbo = 8 * sw;
    for (int k = 0; k < KA / 16; ++k) {
      const uint64_t db = umma_smem_desc(smem_u32(sB) + k * 32, sbo, lt);
      if (mode == 0) {
        const uint64_t da = umma_smem_desc(smem_u32(sA) + k * 32, sbo, lt);
        umma_ss(tbase, da, db, idesc, k != 0);
      } else {
        umma_ts(tbase, tbase + acol + k * 8, db, idesc, k != 0);
      }
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tmem_ld16(tbase + (uint32_t(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) D[(size_t)threadIdx.x * N + c + e] = __uint_as_float(r[e]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

extern "C" int probe_mma(const void* A, const void* B, float* D, int N, int KA, int sw, int mode, void* stream) {
  // A: bf16 [128, KA], B: bf16 [N, KA]; KA*2 == sw bytes
  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {(uint64_t)KA, 128}, dimsB[2] = {(uint64_t)KA, (uint64_t)N};
  uint64_t str[1] = {(uint64_t)KA * 2};
  uint32_t boxA[2] = {(uint32_t)KA, 128}, boxB[2] = {(uint32_t)KA, (uint32_t)N};
  if (int e = make_tmap_bf16(&tmA, A, 2, dimsA, str, boxA, sw)) return e;
  if (int e = make_tmap_bf16(&tmB, B, 2, dimsB, str, boxB, sw)) return e;
  const int smem = 128 * 128 + 256 * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, (const bf16*)A, D, N, KA, sw, mode);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
