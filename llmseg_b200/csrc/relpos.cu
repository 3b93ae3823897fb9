// llmseg_b200 — decomposed rel-pos prep for the SAM 14x14 windows (reference image_encoder.py:354-392):
//   qext[row, j]      = q[row] · rel_pos_h[qh + 13 - j] / scale      j < 14
//   qext[row, 14 + j] = q[row] · rel_pos_w[qw + 13 - j] / scale      j < 14        (row = qh*14 + qw)
// One skinny GEMM QR = q · [rel_h | rel_w]ᵀ (M = all window rows, N = 64, K = 80) followed by a per-row
// shifted gather.  HBM-bound (160 B read + 64 B written per row, ~6 flop/B): the table stays resident in
// shared memory, q tiles stream through a 6-deep TMA ring, the 128x64 accumulators rotate through four
// TMEM stages, and eight epilogue warps (two tiles in flight) stage their own 32 rows in shared memory to
// turn the dynamic column index into conflict-free LDS, then write 64 contiguous bytes per row.
#include <atomic>

#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
namespace {

struct RCfg {
  static constexpr int A_BYTES = 16384 + 4096;  // [128 x 64] SW128 + [128 x 16] SW32
  static constexpr int STAGES = 6;
  static constexpr int R_BYTES = 8192 + 2048;   // [64 x 64] SW128 + [64 x 16] SW32
  static constexpr int ACC = 4;                 // TMEM accumulator stages of 64 columns
  static constexpr int STG_BYTES = 8 * 64 * 32 * 4;
  static constexpr int OFF_R = STAGES * A_BYTES;
  static constexpr int OFF_STG = OFF_R + 10240;
  static constexpr int OFF_BAR = OFF_STG + STG_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

struct RelDev {
  bf16* qext;
  int M, seq, seq_pad, n_tiles;
  float inv_scale;
};

__global__ void __launch_bounds__(320, 1)
relpos_win_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                  const __grid_constant__ CUtensorMap tmRa, const __grid_constant__ CUtensorMap tmRb,
                  const RelDev p) {
  using C = RCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = full + C::STAGES;       // [STAGES]
  uint64_t* acc_full = empty + C::STAGES;   // [ACC]
  uint64_t* acc_empty = acc_full + C::ACC;  // [ACC]
  uint64_t* r_full = acc_empty + C::ACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(r_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQa);
    tma_prefetch_desc(&tmQb);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < C::ACC; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(r_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // prologue done; everything below may read what the previous kernel wrote

  // producer / MMA warps: warp-uniform loops, one elected lane issues (operands stay in uniform registers)
  if (warp == 0) {
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(r_full, C::R_BYTES);
      tma_load_2d(smem + C::OFF_R, &tmRa, r_full, 0, 0);
      tma_load_2d(smem + C::OFF_R + 8192, &tmRb, r_full, 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sa = smem + stage * C::A_BYTES;
      if (issuer) {
        mbar_expect_tx(&full[stage], C::A_BYTES);
        tma_load_2d(sa, &tmQa, &full[stage], 0, tile * 128);
        tma_load_2d(sa + 16384, &tmQb, &full[stage], 64, tile * 128);
      }
      __syncwarp();
      if (++stage == C::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const bool issuer = elect_one();
    const uint32_t sr = smem_u32(smem + C::OFF_R);
    const uint64_t dR0 = umma_smem_desc(sr, 1024, UMMA_SW128), dR1 = umma_smem_desc(sr + 8192, 256, UMMA_SW32);
    const uint32_t smem_base = smem_u32(smem);
    mbar_wait(r_full, 0);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[as], aphase ^ 1);
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_base + stage * C::A_BYTES;
      const uint32_t d_tmem = tmem_base + as * 64;
      const uint64_t dA0 = umma_smem_desc(sa, 1024, UMMA_SW128), dA1 = umma_smem_desc(sa + 16384, 256, UMMA_SW32);
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d_tmem, dA0 + 2 * k, dR0 + 2 * k, idesc, k != 0);
        umma_ss(d_tmem, dA1, dR1, idesc, 1);
        umma_commit(&empty[stage]);
        umma_commit(&acc_full[as]);
      }
      __syncwarp();
      if (++stage == C::STAGES) {
        stage = 0;
        phase ^= 1;
      }
      if (++as == C::ACC) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ---- epilogue: warp group g = (warp-2)/4 takes every other tile of this CTA; thread = one row ----
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;
    float* stg = reinterpret_cast<float*>(smem + C::OFF_STG) + (warp - 2) * (64 * 32);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int as = it & (C::ACC - 1);
      const uint32_t aphase = (it >> 2) & 1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * 64;
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) stg[e * 32 + lane] = __uint_as_float(r0[e]);
#pragma unroll
        for (int e = 0; e < 32; ++e) stg[(32 + e) * 32 + lane] = __uint_as_float(r1[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);  // accumulators are in shared memory now
      const int row = tile * 128 + quarter * 32 + lane;
      const int s = row % p.seq_pad;
      const bool live = row < p.M && s < p.seq;
      const int qh = s / 14, qw = s - qh * 14;
      const float* sh = stg + (qh + 13) * 32 + lane;
      const float* sw = stg + (32 + qw + 13) * 32 + lane;
      float v[32];
#pragma unroll
      for (int j = 0; j < 14; ++j) {
        v[j] = sh[-j * 32] * p.inv_scale;
        v[14 + j] = sw[-j * 32] * p.inv_scale;
      }
      v[28] = v[29] = v[30] = v[31] = 0.f;
      if (live) {
        uint4* dst = reinterpret_cast<uint4*>(p.qext + (size_t)row * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(v[8 * g], v[8 * g + 1]);
          o.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
          o.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]);
          o.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
          dst[g] = o;
        }
      }
      __syncwarp();  // staging rows are reused by this warp's next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// q: bf16 [rows, 80] (rows = bh * seq_pad), rel_hw: bf16 [64, 80] (rel_h rows 0..26, rel_w rows 32..58),
// qext: bf16 [rows, 32].  Called by llmseg_relpos_prep for grid 14 / head_dim 80.
int launch_relpos_win(const void* q, const void* rel_hw, int rows, int seq, int seq_pad, float inv_scale,
                      void* qext, cudaStream_t stream) {
  using C = RCfg;
  CUtensorMap tmQa, tmQb, tmRa, tmRb;
  {
    uint64_t dims[2] = {80, (uint64_t)rows};
    uint64_t str[1] = {160};
    uint32_t a[2] = {64, 128}, b[2] = {16, 128};
    if (int e = make_tmap_bf16(&tmQa, q, 2, dims, str, a, 128)) return e;
    if (int e = make_tmap_bf16(&tmQb, q, 2, dims, str, b, 32)) return e;
  }
  {
    uint64_t dims[2] = {80, 64};
    uint64_t str[1] = {160};
    uint32_t a[2] = {64, 64}, b[2] = {16, 64};
    if (int e = make_tmap_bf16(&tmRa, rel_hw, 2, dims, str, a, 128)) return e;
    if (int e = make_tmap_bf16(&tmRb, rel_hw, 2, dims, str, b, 32)) return e;
  }
  RelDev d{};
  d.qext = static_cast<bf16*>(qext);
  d.M = rows;
  d.seq = seq;
  d.seq_pad = seq_pad;
  d.n_tiles = (rows + 127) / 128;
  d.inv_scale = inv_scale;
  static bool attr_done = false;
  if (!attr_done) {
    LLMSEG_CUDA(cudaFuncSetAttribute(relpos_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = d.n_tiles < sms ? d.n_tiles : sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr, 0);
  LLMSEG_CUDA(cudaLaunchKernelEx(&cfg, relpos_win_kernel, tmQa, tmQb, tmRa, tmRb, d));
  g_launches.fetch_add(1);
  return 0;
}

}  // namespace llmseg
