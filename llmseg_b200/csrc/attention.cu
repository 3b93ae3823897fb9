// llmseg_b200 — fused flash-style attention for sm_100a:  O = softmax(scale·[q|qext]·[k|kext]ᵀ + mask)·V
//
// One CTA = one (batch·head, 128-query tile); 320 threads, two CTAs co-resident per SM so one
// CTA's softmax overlaps the other's tensor-core work.
//   warp 0      TMA producer: Q tile once, then K tile and Vᵀ tile per 128-key step (single buffers
//               with separate full/empty barriers: K(j+1) streams in under softmax(j)+PV(j))
//   warp 1      TMEM allocator + MMA issuer (one lane):
//                 S(128×128 fp32, TMEM cols 0..127)   = Q·Kᵀ   tcgen05.mma SS, K-major operands
//                 O(128×HD  fp32, TMEM cols 128..)   += P·V    tcgen05.mma TS: P is read from TMEM
//   warps 2..9  softmax: thread = (query row, 64-key half) (tcgen05.ld 32x32b).  Two passes over S in
//               TMEM (row max, then exp2 → bf16 P written back over S), lazy rescale of O (only when
//               the running max grew by > 2^8), final O/l → HBM.
// S never reaches shared or global memory; HBM traffic is Q, K, V, O only.
//
// Decomposed rel-pos (reference image_encoder.py:354-392) enters through the extended reduction
// columns qext/kext (see include/llmseg_b200.h): windows use 32 extra columns (rel_h and rel_w),
// global attention 64 extra columns (rel_w) plus a per-(row, 64-key block) additive constant (rel_h).
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
namespace {

constexpr int BM = 128;  // queries per CTA
constexpr int BN = 128;  // keys per step
constexpr float LOG2E = 1.4426950408889634f;

template <int HD, int EXT>
struct ACfg {
  static constexpr int Q0_BYTES = 128 * 128;                                     // [128 x 64] SW128
  static constexpr int Q1_BYTES = HD == 80 ? 128 * 32 : (HD == 128 ? 128 * 128 : 0);
  static constexpr int QX_BYTES = EXT == 1 ? 128 * 64 : (EXT == 2 ? 128 * 128 : 0);
  static constexpr int E_BYTES = EXT ? 16384 : 0;  // EXT1: 2 tiles x [128 x 32]; EXT2: [128 x 64]
  static constexpr int V_CHUNK = HD * 128;         // [HD x 64 keys] SW128
  static constexpr int OFF_Q0 = 0;
  static constexpr int OFF_Q1 = OFF_Q0 + Q0_BYTES;
  static constexpr int OFF_QX = OFF_Q1 + Q1_BYTES;
  static constexpr int OFF_E = OFF_QX + QX_BYTES;
  static constexpr int OFF_K0 = OFF_E + E_BYTES;
  static constexpr int OFF_K1 = OFF_K0 + Q0_BYTES;
  static constexpr int OFF_V = OFF_K1 + Q1_BYTES;
  static constexpr int OFF_BAR = OFF_V + 2 * V_CHUNK;
  static constexpr int OFF_XCH = OFF_BAR + 128;                 // float[2][2][128] row-max exchange + float[2][128] row sums
  static constexpr int SMEM_BYTES = OFF_XCH + 3072 + 1024;
  static constexpr int Q_TX = Q0_BYTES + Q1_BYTES + QX_BYTES + E_BYTES;
  static constexpr int K_TX = Q0_BYTES + Q1_BYTES;
  static constexpr int V_TX = 2 * V_CHUNK;
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = 128;
};

struct AttnDev {
  int dbg;  // developer timing knob (LLMSEG_ATTN_DBG): 1 no ex2, 2 no TMEM score loads, 3 no MMAs, 4 no P stores
  bf16* out;
  int ldo;
  int heads, seq, seq_pad;
  float c1;  // scale * log2(e)
  int causal;
  const int* kv_len;
  const bf16* row_bias;  // [(bh), seq_pad, 64] or null
  const int* out_row_map;  // [(batch*seq)] output row of query (b, s), negative = not written; or null
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- softmax building blocks (thread = one query row x 64 keys).  MASK is a template parameter and the
// callers branch on it explicitly: when the compiler if-converts a runtime mask test it executes the
// index/compare/select instructions for every element of every tile (measured: 2x the instruction count
// on the unmasked SAM-global tiles), and instruction issue is the limiter of this kernel. ----
template <bool MASK>
__device__ __forceinline__ float softmax_row_max(uint32_t t_s, float c1, float add, int key0, int lim) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(t_s + c * 32, r);
    tmem_ld_wait();
    if (MASK) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float t = fmaf(__uint_as_float(r[e]), c1, add);
        mx = fmaxf(mx, (key0 + c * 32 + e < lim) ? t : -INFINITY);
      }
    } else {
      float mc = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; ++e) mc = fmaxf(mc, __uint_as_float(r[e]));
      mx = fmaxf(mx, fmaf(mc, c1, add));  // c1 > 0: max commutes with the affine map
    }
  }
  return mx;
}

// P = exp2(s*c1 + addm) for this thread's 64 keys, packed bf16 into pk[32]; returns the max exponent
// argument (relative to the reference max folded into addm) and accumulates the row sum.
template <bool MASK>
__device__ __forceinline__ float softmax_exp_regs(uint32_t t_s, float c1, float addm, int key0, int lim,
                                                  float& lsum, uint32_t* pk, int dbg = 0) {
  float mx = -INFINITY;
  // both 32-column loads are issued before the single wait: one TMEM round trip per tile instead of two
  uint32_t rr[64];
  if (dbg == 2) {
#pragma unroll
    for (int e = 0; e < 64; ++e) rr[e] = __float_as_uint(0.001f * (float)e);
  } else {
    tmem_ld32(t_s, rr);
    tmem_ld32(t_s + 32, rr + 32);
    tmem_ld_wait();
  }
  // packed pairs (FFMA2 / FADD2): this loop is issue-bound — trading half of the ex2 for an FMA-pipe polynomial made
  // every variant of the kernel 10-20 % SLOWER (profiles/round2_attn.md), so instructions are what to save here
  const float2 c2 = make_float2(c1, c1), a2 = make_float2(addm, addm);
  float2 ls = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t* r = rr + c * 32;
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float2 t = ffma2(make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, a2);
      if (MASK) {
        if (key0 + c * 32 + e >= lim) t.x = -INFINITY;
        if (key0 + c * 32 + e + 1 >= lim) t.y = -INFINITY;
      }
      mx = fmaxf(mx, fmaxf(t.x, t.y));
      const float p0 = dbg == 1 ? t.x * 0.01f : ex2(t.x), p1 = dbg == 1 ? t.y * 0.01f : ex2(t.y);
      ls = fadd2(ls, make_float2(p0, p1));
      pk[c * 16 + (e >> 1)] = pack_bf16(p0, p1);
    }
  }
  lsum += ls.x + ls.y;
  return mx;
}

// slow path: recompute P against a new max and store it straight to TMEM
template <bool MASK>
__device__ __forceinline__ float softmax_exp_store(uint32_t t_s, float c1, float addm, int key0, int lim) {
  float lsum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(t_s + c * 32, r);
    tmem_ld_wait();
    uint32_t pq[16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float t0 = fmaf(__uint_as_float(r[e]), c1, addm);
      float t1 = fmaf(__uint_as_float(r[e + 1]), c1, addm);
      if (MASK) {
        if (key0 + c * 32 + e >= lim) t0 = -INFINITY;
        if (key0 + c * 32 + e + 1 >= lim) t1 = -INFINITY;
      }
      const float p0 = ex2(t0), p1 = ex2(t1);
      lsum += p0 + p1;
      pq[e >> 1] = pack_bf16(p0, p1);
    }
    tmem_st16(t_s + c * 16, pq);
  }
  return lsum;
}

template <int HD, int EXT>
__global__ void __launch_bounds__(320, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
            const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQx,
            const __grid_constant__ CUtensorMap tmE, const AttnDev p) {
  using C = ACfg<HD, EXT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* bar_s = bars + 5;
  uint64_t* bar_p = bars + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQa);
    tma_prefetch_desc(&tmKa);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_p, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // prologue done; everything below may read what the previous kernel wrote

  int kv_limit = p.seq;
  if (p.kv_len) kv_limit = min(kv_limit, max(p.kv_len[b], 1));
  int kv_hi = kv_limit;
  if (p.causal) kv_hi = min(kv_hi, q0 + BM);
  const int n_tiles = (kv_hi + BN - 1) / BN;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (all 32 lanes walk the loop, one elected lane issues: uniform control flow keeps the operands in
    //  uniform registers — see the MMA issuer below)
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(bar_q, C::Q_TX);
      tma_load_3d(smem + C::OFF_Q0, &tmQa, bar_q, 0, q0, bh);
      if (HD == 80) tma_load_3d(smem + C::OFF_Q1, &tmQb, bar_q, 64, q0, bh);
      if (HD == 128) tma_load_3d(smem + C::OFF_Q1, &tmQa, bar_q, 64, q0, bh);
      if (EXT == 1) {
        tma_load_3d(smem + C::OFF_QX, &tmQx, bar_q, 0, q0, bh);
        tma_load_2d(smem + C::OFF_E, &tmE, bar_q, 0, 0);
        tma_load_2d(smem + C::OFF_E + 8192, &tmE, bar_q, 0, 128);
      }
      if (EXT == 2) {
        tma_load_3d(smem + C::OFF_QX, &tmQx, bar_q, 0, q0, bh);
        tma_load_2d(smem + C::OFF_E, &tmE, bar_q, 0, 0);
      }
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const int key0 = j * BN;
      mbar_wait(k_empty, ph ^ 1);
      if (issuer) {
        mbar_expect_tx(k_full, C::K_TX);
        tma_load_3d(smem + C::OFF_K0, &tmKa, k_full, 0, key0, bh);
        if (HD == 80) tma_load_3d(smem + C::OFF_K1, &tmKb, k_full, 64, key0, bh);
        if (HD == 128) tma_load_3d(smem + C::OFF_K1, &tmKa, k_full, 64, key0, bh);
      }
      __syncwarp();
      mbar_wait(v_empty, ph ^ 1);
      if (issuer) {
        mbar_expect_tx(v_full, C::V_TX);
        tma_load_3d(smem + C::OFF_V, &tmV, v_full, key0, 0, bh);
        tma_load_3d(smem + C::OFF_V + C::V_CHUNK, &tmV, v_full, key0 + 64, 0, bh);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The loop runs on all 32 lanes with the tcgen05 instructions under an elected-lane predicate.  With
    // the loop inside `if (lane == 0)` ptxas wrapped each of the 17 MMAs of a tile in a ~20-instruction
    // ELECT / R2UR loop, and the issuer thread — not the tensor pipe, the MUFU or TMEM — set the
    // ~2300 clk tile time every variant in profiles/r01f_attention_investigation.md ran into.
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, HD);
      const bool issuer = elect_one();
      const uint32_t sQ0 = smem_u32(smem + C::OFF_Q0), sQ1 = smem_u32(smem + C::OFF_Q1);
      const uint32_t sQX = smem_u32(smem + C::OFF_QX), sE = smem_u32(smem + C::OFF_E);
      const uint32_t sK0 = smem_u32(smem + C::OFF_K0), sK1 = smem_u32(smem + C::OFF_K1);
      const uint32_t sV = smem_u32(smem + C::OFF_V);
      const uint32_t tS = tmem_base, tO = tmem_base + C::O_COL;
      // every operand lives at a fixed shared-memory address: the descriptors are loop invariants
      const uint64_t dQ0 = umma_smem_desc(sQ0, 1024, UMMA_SW128), dK0 = umma_smem_desc(sK0, 1024, UMMA_SW128);
      const uint64_t dQ1s = umma_smem_desc(sQ1, 256, UMMA_SW32), dK1s = umma_smem_desc(sK1, 256, UMMA_SW32);
      const uint64_t dQ1 = umma_smem_desc(sQ1, 1024, UMMA_SW128), dK1 = umma_smem_desc(sK1, 1024, UMMA_SW128);
      const uint64_t dQXw = umma_smem_desc(sQX, 512, UMMA_SW64), dEw = umma_smem_desc(sE, 512, UMMA_SW64);
      const uint64_t dQXg = umma_smem_desc(sQX, 1024, UMMA_SW128), dEg = umma_smem_desc(sE, 1024, UMMA_SW128);
      const uint64_t dV0 = umma_smem_desc(sV, 1024, UMMA_SW128);
      const uint64_t dV1 = umma_smem_desc(sV + C::V_CHUNK, 1024, UMMA_SW128);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_full, ph);
        tc_fence_after();
        if (issuer) {
          // ---- S = [q|qext]·[k|kext]ᵀ ----  (+32 bytes along K = +2 in the descriptor's address field)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dQ0 + 2 * k, dK0 + 2 * k, idesc_s, k != 0);
          if (HD == 80) umma_ss(tS, dQ1s, dK1s, idesc_s, 1);
          if (HD == 128) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tS, dQ1 + 2 * k, dK1 + 2 * k, idesc_s, 1);
          }
          if (EXT == 1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_ss(tS, dQXw + 2 * k, dEw + (uint64_t)(j * 512 + 2 * k), idesc_s, 1);
          }
          if (EXT == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tS, dQXg + 2 * k, dEg + 2 * k, idesc_s, 1);
          }
          umma_commit(k_empty);
          umma_commit(bar_s);
        }
        __syncwarp();
        // ---- O += P·V ----
        mbar_wait(bar_p, ph);
        mbar_wait(v_full, ph);
        tc_fence_after();
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 8; ++k)  // P(keys 0..63) sits at S columns [0,32), P(keys 64..127) at [64,96)
            umma_ts(tO, tS + (k >> 2) * 64 + (k & 3) * 8, (k < 4 ? dV0 : dV1) + 2 * (k & 3), idesc_o,
                    (j | k) != 0);
          umma_commit(v_empty);
        }
        __syncwarp();
      }
      if (issuer) umma_commit(bar_s);  // final: all PV done
      __syncwarp();
    }
  } else {
    // ================================ softmax / correction / epilogue ================================
    // 8 warps: two per TMEM lane quarter; warp "half" h owns score columns [64h, 64h+64) of its 32 rows
    // and a share of the O columns.  Row maxima / sums are exchanged through shared memory.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + row_in_tile;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t tS_mine = t_row + half * 64;
    const uint32_t tO = t_row + C::O_COL;
    float* xmax = reinterpret_cast<float*>(smem + C::OFF_XCH);        // [2 parity][2 half][128]
    float* xsum = xmax + 512;                                          // [2 half][128]
    constexpr int O_CHUNKS = HD / 16;
    const int oc0 = half == 0 ? 0 : (O_CHUNKS + 1) / 2;
    const int oc1 = half == 0 ? (O_CHUNKS + 1) / 2 : O_CHUNKS;
    const int lim = p.causal ? min(kv_limit, q_row + 1) : kv_limit;  // key valid iff key < lim
    const bf16* rb = nullptr;
    if (EXT == 2) rb = p.row_bias + ((size_t)bh * p.seq_pad + min(q_row, p.seq_pad - 1)) * 64;
    const float c1 = p.c1;
    float m = -INFINITY, l = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const int key0 = j * BN + half * 64;
      float add = 0.f;
      if (EXT == 2) add = __bfloat162float(rb[2 * j + half]) * LOG2E;
      const bool need_mask = (key0 + 64 > kv_limit) || (p.causal && key0 + 63 > q0);
      mbar_wait(bar_s, ph);
      tc_fence_after();

      // ---- single pass (fast path): P = exp2(t - m) against the RUNNING max m while tracking this
      //      tile's max; S is read from TMEM once.  Only when some row's max grew by more than 2^8 —
      //      always on the first tile, rarely afterwards — the slow path rescales O and recomputes P. ----
      float mx, lsum = 0.f;
      uint32_t pk[32];
      if (j == 0) {
        mx = need_mask ? softmax_row_max<true>(tS_mine, c1, add, key0, lim)
                       : softmax_row_max<false>(tS_mine, c1, add, key0, lim);
      } else {
        const float m_fast = (m == -INFINITY) ? 0.f : m;
        mx = (need_mask ? softmax_exp_regs<true>(tS_mine, c1, add - m_fast, key0, lim, lsum, pk)
                        : softmax_exp_regs<false>(tS_mine, c1, add - m_fast, key0, lim, lsum, pk)) +
             m_fast;  // back to absolute (log2-domain) units
      }
      xmax[(ph * 2 + half) * 128 + row_in_tile] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, xmax[(ph * 2 + (half ^ 1)) * 128 + row_in_tile]);
      const float m_new = fmaxf(m, mx);

      // both halves of a row quarter see identical (m, m_new) and therefore take the same branch
      if (j == 0 || __any_sync(0xffffffffu, m_new > m + 8.0f)) {
        if (j > 0) {
          float f = ex2(m - m_new);
          if (m_new == -INFINITY) f = 1.f;
#pragma unroll 1
          for (int c = oc0; c < oc1; ++c) {
            uint32_t r[16];
            tmem_ld16(tO + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * f);
            tmem_st16(tO + c * 16, r);
          }
          l *= f;
        }
        m = m_new;
        const float addm = add - ((m == -INFINITY) ? 0.f : m);
        lsum = need_mask ? softmax_exp_store<true>(tS_mine, c1, addm, key0, lim)
                         : softmax_exp_store<false>(tS_mine, c1, addm, key0, lim);
      } else {
        tmem_st16(tS_mine, pk);
        tmem_st16(tS_mine + 16, pk + 16);
      }
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
    }

    // ---- epilogue: O / l → bf16 → out[b*seq + q_row, h*HD + d] ----
    if (n_tiles > 0) {
      mbar_wait(bar_s, n_tiles & 1);
      tc_fence_after();
    }
    xsum[half * 128 + row_in_tile] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xsum[(half ^ 1) * 128 + row_in_tile];
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    const int h = bh - b * p.heads;
    long long out_r = (long long)b * p.seq + q_row;
    if (p.out_row_map != nullptr) out_r = q_row < p.seq ? p.out_row_map[out_r] : -1;
    bf16* orow = p.out + (size_t)(out_r < 0 ? 0 : out_r) * p.ldo + h * HD;
#pragma unroll 1
    for (int c = oc0; c < oc1; ++c) {
      uint32_t r[16];
      if (n_tiles > 0) {
        tmem_ld16(tO + c * 16, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0;
      }
      if (q_row < p.seq && out_r >= 0) {
        uint4 o0, o1;
        o0.x = pack_bf16(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
        o0.y = pack_bf16(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
        o0.z = pack_bf16(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
        o0.w = pack_bf16(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        o1.x = pack_bf16(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l);
        o1.y = pack_bf16(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l);
        o1.z = pack_bf16(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l);
        o1.w = pack_bf16(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 16) = o0;
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// v3 fast path, one 32-column chunk already in registers: P = exp2(s*c1 + addm) packed into pk16[16]; tracks the
// chunk's max exponent argument (two chains) and the row sum (two chains: a single FADD chain of 64 was latency).
template <bool MASK>
__device__ __forceinline__ void softmax_exp_chunk(const uint32_t* r, float c1, float addm, int key0, int lim,
                                                  float& mx0, float& mx1, float& l0, float& l1, uint32_t* pk16) {
  const float2 c2 = make_float2(c1, c1), a2 = make_float2(addm, addm);
  float2 ls = make_float2(l0, l1);
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float2 t = ffma2(make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, a2);
    if (MASK) {
      if (key0 + e >= lim) t.x = -INFINITY;
      if (key0 + e + 1 >= lim) t.y = -INFINITY;
    }
    mx0 = fmaxf(mx0, t.x);
    mx1 = fmaxf(mx1, t.y);
    const float p0 = ex2(t.x), p1 = ex2(t.y);
    ls = fadd2(ls, make_float2(p0, p1));
    pk16[e >> 1] = pack_bf16(p0, p1);
  }
  l0 = ls.x;
  l1 = ls.y;
}

// =============================================================================================
// v3: 64-key score tiles DOUBLE-BUFFERED in TMEM.
//
// v1 above serialises  S(j) -> softmax(j) -> PV(j) -> S(j+1)  inside a CTA (one score buffer) and leans on the
// second CTA of the SM for overlap: ncu (profiles/r02_summary.md, r02o) had the tensor pipe 53 % and the MUFU
// 61 % active, 21.7 % of all samples softmax warps waiting for their next score tile and the MMA warp waiting
// for P 61 % of its time.  Here a CTA keeps TWO score buffers of 64 keys (TMEM columns [0,64) and [64,128),
// O behind them: 128 + HD <= 256 columns, so two CTAs still share an SM) and the MMA warp runs one tile ahead:
//
//        S(0)  S(1) PV(0)  S(2) PV(1)  S(3) PV(2) ...
//
// S(j+1) is computed under softmax(j), so the softmax warps — the MUFU / TMEM-read side, which bounds this
// kernel at ~1024 clk per 128x128 scores against 896 clk of MMAs at head_dim 80 — go from one tile straight to
// the next.  tcgen05.mma executes in issue order, so S(j+2) overwriting the buffer P(j) was read from needs no
// barrier; PV(j-1) may still be in flight when softmax(j) starts, so the (rare) O rescale waits for it.
//   warp 0      TMA producer: Q (+qext, kext) once, then K(j) / V^T(j) through a 2-3 stage ring
//   warp 1      TMEM allocator + MMA issuer (one elected lane, warp-uniform loop)
//   warps 2..5  softmax: thread = one query row x the tile's 64 keys (no cross-warp exchange, no named
//               barrier); single pass over S against the running max, lazy O rescale (max grew by > 2^8)
// =============================================================================================
constexpr int BN3 = 64;

template <int HD, int EXT>
struct A3Cfg {
  static constexpr int Q0_BYTES = 128 * 128;                                     // [128 x 64] SW128
  static constexpr int Q1_BYTES = HD == 80 ? 128 * 32 : (HD == 128 ? 128 * 128 : 0);
  static constexpr int QX_BYTES = EXT == 1 ? 128 * 64 : (EXT == 2 ? 128 * 128 : 0);
  static constexpr int E_BYTES = EXT == 1 ? 256 * 64 : (EXT == 2 ? 64 * 128 : 0);  // EXT1: [256 x 32] SW64; EXT2: [64 x 64] SW128
  static constexpr int K0_BYTES = BN3 * 128;                                     // [64 x 64] SW128
  static constexpr int K1_BYTES = HD == 80 ? BN3 * 32 : (HD == 128 ? BN3 * 128 : 0);
  static constexpr int V_BYTES = HD * 128;                                       // [HD x 64 keys] SW128
  static constexpr int STAGE_BYTES = K0_BYTES + K1_BYTES + V_BYTES;
  static constexpr int STAGES = HD == 128 ? 2 : 3;
  static constexpr int OFF_Q0 = 0;
  static constexpr int OFF_Q1 = OFF_Q0 + Q0_BYTES;
  static constexpr int OFF_QX = OFF_Q1 + Q1_BYTES;
  static constexpr int OFF_E = OFF_QX + QX_BYTES;
  static constexpr int OFF_ST = OFF_E + E_BYTES;
  static constexpr int OFF_BAR = OFF_ST + STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int Q_TX = Q0_BYTES + Q1_BYTES + QX_BYTES + E_BYTES;
  static constexpr int K_TX = K0_BYTES + K1_BYTES;
  static constexpr int V_TX = V_BYTES;
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = 128;
  static_assert(OFF_ST % 1024 == 0 && STAGE_BYTES % 1024 == 0 && (K0_BYTES + K1_BYTES) % 1024 == 0, "swizzle atoms");
  static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
};

template <int HD, int EXT>
__global__ void __launch_bounds__(192, 2)
attn3_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
             const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQx,
             const __grid_constant__ CUtensorMap tmE, const AttnDev p) {
  using C = A3Cfg<HD, EXT>;
  constexpr int ST = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* k_full = bars + 1;          // [ST]
  uint64_t* k_empty = bars + 1 + ST;    // [ST]
  uint64_t* v_full = bars + 1 + 2 * ST; // [ST]
  uint64_t* v_empty = bars + 1 + 3 * ST;
  uint64_t* bar_s = bars + 1 + 4 * ST;  // [2] score buffer (j & 1) holds S(j)
  uint64_t* bar_p = bar_s + 2;          // [2] P(j) written over it (4 elected arrivals)
  uint64_t* bar_pv = bar_p + 2;         // PV(j) complete (one completion per tile)
  uint64_t* bar_done = bar_pv + 1;      // every PV complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQa);
    tma_prefetch_desc(&tmKa);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 1 + 4 * ST + 2; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_p[0], 4);
    mbar_init(&bar_p[1], 4);
    mbar_init(bar_pv, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // prologue done; everything below may read what the previous kernel wrote

  int kv_limit = p.seq;
  if (p.kv_len) kv_limit = min(kv_limit, max(p.kv_len[b], 1));
  int kv_hi = kv_limit;
  if (p.causal) kv_hi = min(kv_hi, q0 + BM);
  const int n_tiles = (kv_hi + BN3 - 1) / BN3;

  if (warp == 0) {
    // ================================ TMA producer ================================
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(bar_q, C::Q_TX);
      tma_load_3d(smem + C::OFF_Q0, &tmQa, bar_q, 0, q0, bh);
      if (HD == 80) tma_load_3d(smem + C::OFF_Q1, &tmQb, bar_q, 64, q0, bh);
      if (HD == 128) tma_load_3d(smem + C::OFF_Q1, &tmQa, bar_q, 64, q0, bh);
      if (EXT) {
        tma_load_3d(smem + C::OFF_QX, &tmQx, bar_q, 0, q0, bh);
        tma_load_2d(smem + C::OFF_E, &tmE, bar_q, 0, 0);
      }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int key0 = j * BN3;
      uint8_t* st = smem + C::OFF_ST + s * C::STAGE_BYTES;
      mbar_wait(&k_empty[s], ph ^ 1);
      if (issuer) {
        mbar_expect_tx(&k_full[s], C::K_TX);
        tma_load_3d(st, &tmKa, &k_full[s], 0, key0, bh);
        if (HD == 80) tma_load_3d(st + C::K0_BYTES, &tmKb, &k_full[s], 64, key0, bh);
        if (HD == 128) tma_load_3d(st + C::K0_BYTES, &tmKa, &k_full[s], 64, key0, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      if (issuer) {
        mbar_expect_tx(&v_full[s], C::V_TX);
        tma_load_3d(st + C::K0_BYTES + C::K1_BYTES, &tmV, &v_full[s], key0, 0, bh);
      }
      __syncwarp();
      if (++s == ST) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN3);
    constexpr uint32_t idesc_o = umma_idesc_bf16(BM, HD);
    const bool issuer = elect_one();
    const uint32_t sQ0 = smem_u32(smem + C::OFF_Q0), sQ1 = smem_u32(smem + C::OFF_Q1);
    const uint32_t sQX = smem_u32(smem + C::OFF_QX), sE = smem_u32(smem + C::OFF_E);
    const uint32_t sSt = smem_u32(smem + C::OFF_ST);
    const uint32_t tO = tmem_base + C::O_COL;
    const uint64_t dQ0 = umma_smem_desc(sQ0, 1024, UMMA_SW128);
    const uint64_t dQ1s = umma_smem_desc(sQ1, 256, UMMA_SW32);
    const uint64_t dQ1 = umma_smem_desc(sQ1, 1024, UMMA_SW128);
    const uint64_t dQXw = umma_smem_desc(sQX, 512, UMMA_SW64), dEw = umma_smem_desc(sE, 512, UMMA_SW64);
    const uint64_t dQXg = umma_smem_desc(sQX, 1024, UMMA_SW128), dEg = umma_smem_desc(sE, 1024, UMMA_SW128);
    int ks = 0, vs = 0;          // ring positions of the next K / V tile to consume
    uint32_t kph = 0, vph = 0;
    auto issue_s = [&](int j) {
      const uint32_t sk = sSt + ks * C::STAGE_BYTES;
      const uint32_t tS = tmem_base + (j & 1) * BN3;
      mbar_wait(&k_full[ks], kph);
      tc_fence_after();
      if (issuer) {
        const uint64_t dK0 = umma_smem_desc(sk, 1024, UMMA_SW128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, dQ0 + 2 * k, dK0 + 2 * k, idesc_s, k != 0);
        if (HD == 80) umma_ss(tS, dQ1s, umma_smem_desc(sk + C::K0_BYTES, 256, UMMA_SW32), idesc_s, 1);
        if (HD == 128) {
          const uint64_t dK1 = umma_smem_desc(sk + C::K0_BYTES, 1024, UMMA_SW128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dQ1 + 2 * k, dK1 + 2 * k, idesc_s, 1);
        }
        if (EXT == 1) {  // kext rows [64j, 64j+64) of the [256 x 32] one-hot table: 64 rows x 64 B = 4096 B per tile
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_ss(tS, dQXw + 2 * k, dEw + (uint64_t)(j * 256 + 2 * k), idesc_s, 1);
        }
        if (EXT == 2) {  // a 64-key tile is one grid row: the same 64 x 64 identity for every tile
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dQXg + 2 * k, dEg + 2 * k, idesc_s, 1);
        }
        umma_commit(&k_empty[ks]);
        umma_commit(&bar_s[j & 1]);
      }
      __syncwarp();
      if (++ks == ST) {
        ks = 0;
        kph ^= 1;
      }
    };
    auto issue_pv = [&](int j) {
      const uint32_t sv = sSt + vs * C::STAGE_BYTES + C::K0_BYTES + C::K1_BYTES;
      const uint32_t tP = tmem_base + (j & 1) * BN3;
      mbar_wait(&bar_p[j & 1], (uint32_t)((j >> 1) & 1));
      mbar_wait(&v_full[vs], vph);
      tc_fence_after();
      if (issuer) {
        const uint64_t dV = umma_smem_desc(sv, 1024, UMMA_SW128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tO, tP + k * 8, dV + 2 * k, idesc_o, (j | k) != 0);
        umma_commit(&v_empty[vs]);
        umma_commit(bar_pv);
      }
      __syncwarp();
      if (++vs == ST) {
        vs = 0;
        vph ^= 1;
      }
    };
    mbar_wait(bar_q, 0);
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_s(j + 1);
      issue_pv(j);
    }
    if (issuer) umma_commit(bar_done);
    __syncwarp();
  } else {
    // ================================ softmax / correction / epilogue ================================
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + row_in_tile;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t tO = t_row + C::O_COL;
    const int lim = p.causal ? min(kv_limit, q_row + 1) : kv_limit;  // key valid iff key < lim
    const bf16* rb = nullptr;
    if (EXT == 2) rb = p.row_bias + ((size_t)bh * p.seq_pad + min(q_row, p.seq_pad - 1)) * 64;
    const float c1 = p.c1;
    float m = -INFINITY, l = 0.f;
    float add_next = 0.f;
    if (EXT == 2) add_next = __bfloat162float(rb[0]) * LOG2E;
    // The TMEM -> register path (64 B/clk per SM: 512 clk for a CTA's 128 x 64 fp32 tile) and the MUFU (512 clk for
    // its 8192 ex2) are the two bounds of this loop, and ncu on the first v3 (profiles/round2_attn3.md) showed each
    // softmax warp running them one after the other: load all 64 columns, wait, then 64 ex2 — 1858 clk per tile and
    // warp, MUFU 57 % busy.  So the loop is software-pipelined: the tile's second 32 columns are in flight while the
    // first 32 are exponentiated, and the first 32 columns of the NEXT tile (whose MMA finished long ago) are
    // requested before the second half is touched (v4 below; here, with P aliasing S, the next tile's scores are not
    // complete early enough for that).
    uint32_t ra[32], rc[32];
    const bool pre = false;

    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int key0 = j * BN3;
      const uint32_t tS = t_row + sb * BN3;
      const float add = add_next;
      if (EXT == 2 && j + 1 < n_tiles) add_next = __bfloat162float(rb[j + 1]) * LOG2E;
      const bool need_mask = (key0 + BN3 > kv_limit) || (p.causal && key0 + BN3 - 1 > q0);
      float lsum = 0.f;
      if (j == 0) {
        mbar_wait(&bar_s[sb], ph);
        tc_fence_after();
        const float mx = need_mask ? softmax_row_max<true>(tS, c1, add, key0, lim)
                                   : softmax_row_max<false>(tS, c1, add, key0, lim);
        m = mx;
        const float addm = add - ((m == -INFINITY) ? 0.f : m);
        lsum = need_mask ? softmax_exp_store<true>(tS, c1, addm, key0, lim)
                         : softmax_exp_store<false>(tS, c1, addm, key0, lim);
      } else {
        if (!pre) {
          mbar_wait(&bar_s[sb], ph);
          tc_fence_after();
          tmem_ld32(tS, ra);
        }
        // fast path: P against the RUNNING max while tracking this tile's max; S is read from TMEM once
        const float m_fast = (m == -INFINITY) ? 0.f : m;
        const float addm = add - m_fast;
        uint32_t pk[32];
        float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        tmem_ld_wait();
        tmem_ld32(tS + 32, rc);
        if (need_mask) softmax_exp_chunk<true>(ra, c1, addm, key0, lim, mx0, mx1, l0, l1, pk);
        else softmax_exp_chunk<false>(ra, c1, addm, key0, lim, mx0, mx1, l0, l1, pk);
        tmem_ld_wait();
        if (need_mask) softmax_exp_chunk<true>(rc, c1, addm, key0 + 32, lim, mx0, mx1, l0, l1, pk + 16);
        else softmax_exp_chunk<false>(rc, c1, addm, key0 + 32, lim, mx0, mx1, l0, l1, pk + 16);
        const float mx_rel = fmaxf(mx0, mx1);
        lsum = l0 + l1;
        const bool grow = (m == -INFINITY) ? (mx_rel > -INFINITY) : (mx_rel > 8.0f);
        if (__any_sync(0xffffffffu, grow)) {
          // some row's max grew by more than 2^8: rescale O (after PV(j-1), which may still be running) and
          // recompute this tile's P against the new max
          mbar_wait(bar_pv, (uint32_t)((j - 1) & 1));
          tc_fence_after();
          const float g = fmaxf(mx_rel, 0.f);
          const float m_new = (m == -INFINITY) ? mx_rel : m + g;
          float f = (m == -INFINITY) ? 1.f : ex2(-g);
          if (m_new == -INFINITY) f = 1.f;
#pragma unroll 1
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tO + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * f);
            tmem_st16(tO + c * 16, r);
          }
          l *= f;
          m = m_new;
          const float addm2 = add - ((m == -INFINITY) ? 0.f : m);
          lsum = need_mask ? softmax_exp_store<true>(tS, c1, addm2, key0, lim)
                           : softmax_exp_store<false>(tS, c1, addm2, key0, lim);
        } else {
          tmem_st32(tS, pk);
        }
      }
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[sb]);
    }

    // ---- epilogue: O / l → bf16 → out[b*seq + q_row, h*HD + d] ----
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    const int h = bh - b * p.heads;
    long long out_r = (long long)b * p.seq + q_row;
    if (p.out_row_map != nullptr) out_r = q_row < p.seq ? p.out_row_map[out_r] : -1;
    bf16* orow = p.out + (size_t)(out_r < 0 ? 0 : out_r) * p.ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(tO + c * 16, r);
      tmem_ld_wait();
      if (q_row < p.seq && out_r >= 0) {
        uint4 o0, o1;
        o0.x = pack_bf16(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
        o0.y = pack_bf16(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
        o0.z = pack_bf16(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
        o0.w = pack_bf16(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        o1.x = pack_bf16(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l);
        o1.y = pack_bf16(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l);
        o1.z = pack_bf16(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l);
        o1.w = pack_bf16(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 16) = o0;
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// =============================================================================================
// Window kernel: SAM 14x14 windowed attention (196 keys, hd 80, 32 rel-pos extension columns).
//
// The generic kernel spends most of a window CTA's life in its prologue (6400 two-tile CTAs per
// layer at batch 8; ncu: 193 us, 216 TF/s).  Here ONE persistent CTA per SM loops over
// (window, head) items; a whole window fits one 208-key score tile, so softmax is exact in a single
// step (no online rescale) and the kernel keeps, per item:
//   S_t = [q_t|qext_t].[k|kext]^T   two M=128 x N=208 MMAs (query rows 0..127 / 128..255)
//   P_t  -> written over S_t columns [0,104);  O_t = P_t.V accumulates into S_t columns [112,192)
//   TMEM: S_0 at columns [0,208), S_1 at [256,464)  (O aliases the dead tail of S)
// K/V are double-buffered across items, the constant one-hot kext tile is loaded once per CTA,
// Q of the next item streams in as soon as both score MMAs of the current item have been issued.
//   warps: 0 TMA producer, 1 MMA issuer + TMEM alloc, 2..5 softmax rows of tile 0, 6..9 of tile 1
// =============================================================================================
struct WCfg {
  static constexpr int KEYS = 208;
  static constexpr int QT_BYTES = 16384 + 4096 + 8192;   // [128x64] SW128 + [128x16] SW32 + [128x32] SW64
  static constexpr int E_BYTES = 208 * 64;               // [208x32] SW64
  static constexpr int E_ALLOC = 14336;
  static constexpr int K0_BYTES = 208 * 128;             // [208x64] SW128
  static constexpr int K1_BYTES = 208 * 32;              // [208x16] SW32
  static constexpr int K_ALLOC = 26624 + 7168;
  static constexpr int V_CHUNK = 80 * 128;               // [80x64 keys] SW128
  static constexpr int V_TAIL = 80 * 32;                 // [80x16 keys] SW32
  static constexpr int V_ALLOC = 3 * 10240 + 3072;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_E = OFF_Q + 2 * QT_BYTES;
  static constexpr int OFF_K = OFF_E + E_ALLOC;
  static constexpr int OFF_V = OFF_K + 2 * K_ALLOC;
  static constexpr int OFF_BAR = OFF_V + 2 * V_ALLOC;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int Q_TX = 2 * QT_BYTES;
  static constexpr int K_TX = K0_BYTES + K1_BYTES;
  static constexpr int V_TX = 3 * V_CHUNK + V_TAIL;
  static constexpr int O_COL = 112;                      // O_t inside the S_t column range
};

__global__ void __launch_bounds__(320, 1)
attn_win_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                const __grid_constant__ CUtensorMap tmQx, const __grid_constant__ CUtensorMap tmKa,
                const __grid_constant__ CUtensorMap tmKb, const __grid_constant__ CUtensorMap tmE,
                const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb,
                const AttnDev p, const int n_items) {
  using C = WCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_e = bars + 0;
  uint64_t* q_full = bars + 1;    // [2] Q tile t of the current item landed
  uint64_t* q_empty = bars + 3;   // [2] score MMA of tile t done with its Q tile
  uint64_t* k_full = bars + 5;    // [2]
  uint64_t* k_empty = bars + 7;   // [2]
  uint64_t* v_full = bars + 9;    // [2]
  uint64_t* v_empty = bars + 11;  // [2]
  uint64_t* bar_s = bars + 13;    // [2] score tile t ready
  uint64_t* bar_p = bars + 15;    // [2] P_t written          (4 elected arrivals)
  uint64_t* bar_o = bars + 17;    // [2] O_t complete
  uint64_t* o_free = bars + 19;   // [2] O_t/S_t columns released by the epilogue (4 elected arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQa);
    tma_prefetch_desc(&tmKa);
    tma_prefetch_desc(&tmVa);
    for (int i = 0; i < 21; ++i) mbar_init(&bars[i], 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar_p[t], 4);
      mbar_init(&o_free[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // prologue done; everything below may read what the previous kernel wrote

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (warp-uniform loops, elected-lane issue: see attn_kernel)
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(bar_e, C::E_BYTES);
      tma_load_2d(smem + C::OFF_E, &tmE, bar_e, 0, 0);
    }
    __syncwarp();
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t kvph = (it >> 1) & 1;
      uint8_t* sk = smem + C::OFF_K + st * C::K_ALLOC;
      uint8_t* sv = smem + C::OFF_V + st * C::V_ALLOC;
      mbar_wait(&k_empty[st], kvph ^ 1);
      if (issuer) {
        mbar_expect_tx(&k_full[st], C::K_TX);
        tma_load_3d(sk, &tmKa, &k_full[st], 0, 0, item);
        tma_load_3d(sk + 26624, &tmKb, &k_full[st], 64, 0, item);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], kvph ^ 1);
      if (issuer) {
        mbar_expect_tx(&v_full[st], C::V_TX);
#pragma unroll
        for (int c = 0; c < 3; ++c) tma_load_3d(sv + c * C::V_CHUNK, &tmVa, &v_full[st], c * 64, 0, item);
        tma_load_3d(sv + 3 * C::V_CHUNK, &tmVb, &v_full[st], 192, 0, item);
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&q_empty[t], (it & 1) ^ 1);
        if (issuer) {
          uint8_t* sq = smem + C::OFF_Q + t * C::QT_BYTES;
          mbar_expect_tx(&q_full[t], C::QT_BYTES);
          tma_load_3d(sq, &tmQa, &q_full[t], 0, t * 128, item);
          tma_load_3d(sq + 16384, &tmQb, &q_full[t], 64, t * 128, item);
          tma_load_3d(sq + 20480, &tmQx, &q_full[t], 0, t * 128, item);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, C::KEYS);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 80);
      const bool issuer = elect_one();
      const uint32_t sE = smem_u32(smem + C::OFF_E);
      const uint64_t dE = umma_smem_desc(sE, 512, UMMA_SW64);
      mbar_wait(bar_e, 0);
      // The two query tiles of an item run half a period apart:  S0(i)  PV1(i-1)  S1(i)  PV0(i)  S0(i+1) ...
      // Tile 0's warps are in their softmax (MUFU-bound) while tile 1's read O and store, and vice versa, so each
      // softmax has the MUFU pipe to itself and the tensor pipe works for one tile while the other one's threads
      // are busy.  Issued back to back (S0 S1 ... PV0 PV1, both softmaxes at once) an item was one serial chain:
      // 10.7k clk against a 3.3k clk MUFU floor (profiles/r01k).
      auto issue_s = [&](int t, uint32_t sk) {
        const uint32_t sq = smem_u32(smem + C::OFF_Q + t * C::QT_BYTES);
        const uint32_t tS = tmem_base + t * 256;
        const uint64_t dK0 = umma_smem_desc(sk, 1024, UMMA_SW128);
        const uint64_t dK1 = umma_smem_desc(sk + 26624, 256, UMMA_SW32);
        const uint64_t dQ0 = umma_smem_desc(sq, 1024, UMMA_SW128);
        const uint64_t dQ1 = umma_smem_desc(sq + 16384, 256, UMMA_SW32);
        const uint64_t dQX = umma_smem_desc(sq + 20480, 512, UMMA_SW64);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dQ0 + 2 * k, dK0 + 2 * k, idesc_s, k != 0);
          umma_ss(tS, dQ1, dK1, idesc_s, 1);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_ss(tS, dQX + 2 * k, dE + 2 * k, idesc_s, 1);
          umma_commit(&bar_s[t]);
          umma_commit(&q_empty[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, uint32_t sv) {
        const uint32_t tS = tmem_base + t * 256;
        const uint64_t dV = umma_smem_desc(sv, 1024, UMMA_SW128);
        const uint64_t dVt = umma_smem_desc(sv + 3 * C::V_CHUNK, 256, UMMA_SW32);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 12; ++k)
            umma_ts(tS + C::O_COL, tS + k * 8, dV + (uint64_t)((k >> 2) * (C::V_CHUNK >> 4) + (k & 3) * 2), idesc_o,
                    k != 0);
          umma_ts(tS + C::O_COL, tS + 96, dVt, idesc_o, 1);
          umma_commit(&bar_o[t]);
        }
        __syncwarp();
      };
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = it & 1, kvph = (it >> 1) & 1;
        const uint32_t sk = smem_u32(smem + C::OFF_K + st * C::K_ALLOC);
        const uint32_t sv = smem_u32(smem + C::OFF_V + st * C::V_ALLOC);
        // ---- S0(i)
        mbar_wait(&q_full[0], ph);
        mbar_wait(&k_full[st], kvph);
        if (it > 0) mbar_wait(&o_free[0], ph ^ 1);  // previous item's O_0 / S_0 columns drained
        tc_fence_after();
        issue_s(0, sk);
        // ---- PV1(i-1): the other V buffer (its v_full was waited for by PV0(i-1))
        if (it > 0) {
          mbar_wait(&bar_p[1], ph ^ 1);
          tc_fence_after();
          issue_pv(1, smem_u32(smem + C::OFF_V + (st ^ 1) * C::V_ALLOC));
          if (issuer) umma_commit(&v_empty[st ^ 1]);
          __syncwarp();
        }
        // ---- S1(i)
        mbar_wait(&q_full[1], ph);
        if (it > 0) mbar_wait(&o_free[1], ph ^ 1);
        tc_fence_after();
        issue_s(1, sk);
        if (issuer) umma_commit(&k_empty[st]);
        __syncwarp();
        // ---- PV0(i)
        mbar_wait(&v_full[st], kvph);
        mbar_wait(&bar_p[0], ph);
        tc_fence_after();
        issue_pv(0, sv);
      }
      if (it > 0) {  // PV1 of the last item
        const int st = (it - 1) & 1;
        mbar_wait(&bar_p[1], (uint32_t)((it - 1) & 1));
        tc_fence_after();
        issue_pv(1, smem_u32(smem + C::OFF_V + st * C::V_ALLOC));
        if (issuer) umma_commit(&v_empty[st]);
        __syncwarp();
      }
    }
  } else {
    // ================================ softmax: thread = one query row ================================
    const int t = warp >= 6 ? 1 : 0;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = t * 128 + row_in_tile;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
    const float c1 = p.c1;
    const int seq = p.seq;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(&bar_s[t], ph);
      tc_fence_after();
      // ONE pass over the score tile.  TMEM reads run at 64 B/clk per SM: a [256 x 208] fp32 item is 3.3k clk per
      // read, and the separate row-max pass this replaces made the kernel TMEM-read-bound (2 reads + O = 8k of its
      // 9.7k clk per item, profiles/r02m).  Now every 32-key chunk is read once: its maximum joins a RUNNING maximum
      // m_run, P = exp2(s*c1 - m_run) goes back to TMEM as bf16, and only when some row's maximum grows by more than
      // 2^8 over what its earlier chunks were scaled with (warp-uniform vote; rare after the first chunk) the
      // chunks already written are rescaled in place — the same lazy rescale the multi-tile kernel applies to O.
      float l0 = 0.f, l1 = 0.f;
      float m_run = -INFINITY;
      auto chunk_max = [&](const uint32_t* r, int n) {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          if (e + 0 < n) m0 = fmaxf(m0, __uint_as_float(r[e]));
          if (e + 1 < n) m1 = fmaxf(m1, __uint_as_float(r[e + 1]));
          if (e + 2 < n) m2 = fmaxf(m2, __uint_as_float(r[e + 2]));
          if (e + 3 < n) m3 = fmaxf(m3, __uint_as_float(r[e + 3]));
        }
        return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c1;  // c1 > 0: max commutes with the scaling
      };
      // join chunk c's maximum (log2 units); on a large jump rescale P chunks [0, c) and the row sums
      auto advance_max = [&](float cm, int c) {
        if (c == 0) {
          m_run = cm;
          return;
        }
        const bool grow = cm > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float f = grow ? ex2(m_run - cm) : 1.0f;
          if (grow) m_run = cm;
          l0 *= f;
          l1 *= f;
          tmem_st_wait();
#pragma unroll 1
          for (int cc = 0; cc < c; ++cc) {
            uint32_t pq[16];
            tmem_ld16(tS + cc * 16, pq);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float2 v = unpack_bf16(pq[e]);
              pq[e] = pack_bf16(v.x * f, v.y * f);
            }
            tmem_st16(tS + cc * 16, pq);
          }
        }
      };
      auto exp_chunk = [&](const uint32_t* r, int c) {
        advance_max(chunk_max(r, 32), c);
        const float moff = -m_run;
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(r[e]), c1, moff));
          const float p1 = ex2(fmaf(__uint_as_float(r[e + 1]), c1, moff));
          l0 += p0;
          l1 += p1;
          pk[e >> 1] = pack_bf16(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);  // P chunk c lies inside S chunk c/2, already in registers
      };
      {
        uint32_t ra[32], rb[32];
        tmem_ld32(tS, ra);
#pragma unroll
        for (int c = 0; c < 6; c += 2) {
          tmem_ld_wait();
          tmem_ld32(tS + (c + 1) * 32, rb);
          exp_chunk(ra, c);
          tmem_ld_wait();
          if (c + 2 < 6) tmem_ld32(tS + (c + 2) * 32, ra);
          else tmem_ld16(tS + 192, ra);
          exp_chunk(rb, c + 1);
        }
        tmem_ld_wait();
        // keys 192..207: valid below seq (196)
        advance_max(chunk_max(ra, seq - 192 < 16 ? seq - 192 : 16), 6);
        const float moff = -m_run;
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float p0 = (192 + e < seq) ? ex2(fmaf(__uint_as_float(ra[e]), c1, moff)) : 0.f;
          const float p1 = (192 + e + 1 < seq) ? ex2(fmaf(__uint_as_float(ra[e + 1]), c1, moff)) : 0.f;
          l0 += p0;
          l1 += p1;
          pk[e >> 1] = pack_bf16(p0, p1);
        }
#pragma unroll
        for (int e = 8; e < 16; ++e) pk[e] = 0;
        tmem_st16(tS + 96, pk);  // columns [96,104) = keys 192..207, [104,112) scratch
      }
      const float l = l0 + l1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[t]);

      // ---- epilogue: O / l -> out ----
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
      const int b = item / p.heads, h = item - b * p.heads;
      long long out_r = (long long)b * seq + q_row;
      if (p.out_row_map != nullptr) out_r = q_row < seq ? p.out_row_map[out_r] : -1;
      bf16* orow = p.out + (size_t)(out_r < 0 ? 0 : out_r) * p.ldo + h * 80;
      mbar_wait(&bar_o[t], ph);
      tc_fence_after();
      {
        uint32_t r[80];
#pragma unroll
        for (int c = 0; c < 5; ++c) tmem_ld16(tS + C::O_COL + c * 16, r + c * 16);
        tmem_ld_wait();
        if (q_row < seq && out_r >= 0) {
#pragma unroll
          for (int c = 0; c < 10; ++c) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[8 * c + 0]) * inv_l, __uint_as_float(r[8 * c + 1]) * inv_l);
            o.y = pack_bf16(__uint_as_float(r[8 * c + 2]) * inv_l, __uint_as_float(r[8 * c + 3]) * inv_l);
            o.z = pack_bf16(__uint_as_float(r[8 * c + 4]) * inv_l, __uint_as_float(r[8 * c + 5]) * inv_l);
            o.w = pack_bf16(__uint_as_float(r[8 * c + 6]) * inv_l, __uint_as_float(r[8 * c + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c * 8) = o;
          }
        }
      }
      // (fetching this thread's out_row_map entry at the top of the item instead of right before its use measured
      //  SLOWER too — 94.0 vs 86.2 us, same box, two library builds, scripts/gpu_attn_win_time.py: the kernel's two query
      //  tiles run half a period apart and anything that shifts one warp group's timing moves them into each other)
      // (releasing the columns right after the TMEM read, before the row stores, measured SLOWER twice — 134-141 us
      //  against 118 with the two-pass softmax, 98 against 93.6 with the single pass: the stores are part of what
      //  keeps the two tiles half a period apart)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// LLMSEG_ATTN_WIN=0 disables the window kernel (falls back to the generic one)
bool use_attn_win() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LLMSEG_ATTN_WIN");
    mode = e ? atoi(e) : 1;
  }
  return mode == 1;
}

int num_sms_attn() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int launch_attn_win(const llmseg_attn_params* p, cudaStream_t stream) {
  using C = WCfg;
  const int BH = p->batch * p->heads;
  CUtensorMap tmQa, tmQb, tmQx, tmKa, tmKb, tmE, tmVa, tmVb;
  {
    uint64_t dims[3] = {80, (uint64_t)p->seq_pad, (uint64_t)BH};
    uint64_t str[2] = {160, (uint64_t)p->seq_pad * 160};
    uint32_t qa[3] = {64, 128, 1}, qb[3] = {16, 128, 1}, ka[3] = {64, 208, 1}, kb[3] = {16, 208, 1};
    if (int e = make_tmap_bf16(&tmQa, p->q, 3, dims, str, qa, 128)) return e;
    if (int e = make_tmap_bf16(&tmQb, p->q, 3, dims, str, qb, 32)) return e;
    if (int e = make_tmap_bf16(&tmKa, p->k, 3, dims, str, ka, 128)) return e;
    if (int e = make_tmap_bf16(&tmKb, p->k, 3, dims, str, kb, 32)) return e;
  }
  {
    uint64_t dims[3] = {32, (uint64_t)p->seq_pad, (uint64_t)BH};
    uint64_t str[2] = {64, (uint64_t)p->seq_pad * 64};
    uint32_t box[3] = {32, 128, 1};
    if (int e = make_tmap_bf16(&tmQx, p->qext, 3, dims, str, box, 64)) return e;
    uint64_t edims[2] = {32, 256};
    uint64_t estr[1] = {64};
    uint32_t ebox[2] = {32, 208};
    if (int e = make_tmap_bf16(&tmE, p->kext, 2, edims, estr, ebox, 64)) return e;
  }
  {
    uint64_t dims[3] = {(uint64_t)p->seq_pad, 80, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)p->seq_pad * 2, (uint64_t)p->seq_pad * 160};
    uint32_t va[3] = {64, 80, 1}, vb[3] = {16, 80, 1};
    if (int e = make_tmap_bf16(&tmVa, p->vt, 3, dims, str, va, 128)) return e;
    if (int e = make_tmap_bf16(&tmVb, p->vt, 3, dims, str, vb, 32)) return e;
  }
  AttnDev d{};
  d.out = static_cast<bf16*>(p->out);
  d.ldo = p->ldo;
  d.heads = p->heads;
  d.seq = p->seq;
  d.seq_pad = p->seq_pad;
  d.c1 = p->scale * LOG2E;
  d.out_row_map = p->out_row_map;
  static bool attr_done = false;
  if (!attr_done) {
    LLMSEG_CUDA(cudaFuncSetAttribute(attn_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  const int sms = num_sms_attn();
  const int grid = BH < sms ? BH : sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr, 0);
  LLMSEG_CUDA(cudaLaunchKernelEx(&cfg, attn_win_kernel, tmQa, tmQb, tmQx, tmKa, tmKb, tmE, tmVa, tmVb, d, BH));
  g_launches.fetch_add(1);
  return 0;
}

template <int HD, int EXT>
int launch_attn(const llmseg_attn_params* p, cudaStream_t stream) {
  using C = ACfg<HD, EXT>;
  const int BH = p->batch * p->heads;
  CUtensorMap tmQa, tmQb, tmKa, tmKb, tmV, tmQx, tmE;
  {
    uint64_t dims[3] = {(uint64_t)HD, (uint64_t)p->seq_pad, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)HD * 2, (uint64_t)p->seq_pad * HD * 2};
    uint32_t boxa[3] = {64, 128, 1};
    if (int e = make_tmap_bf16(&tmQa, p->q, 3, dims, str, boxa, 128)) return e;
    if (int e = make_tmap_bf16(&tmKa, p->k, 3, dims, str, boxa, 128)) return e;
    tmQb = tmQa;
    tmKb = tmKa;
    if (HD == 80) {
      uint32_t boxb[3] = {16, 128, 1};
      if (int e = make_tmap_bf16(&tmQb, p->q, 3, dims, str, boxb, 32)) return e;
      if (int e = make_tmap_bf16(&tmKb, p->k, 3, dims, str, boxb, 32)) return e;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)p->seq_pad, (uint64_t)HD, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)p->seq_pad * 2, (uint64_t)p->seq_pad * HD * 2};
    uint32_t box[3] = {64, (uint32_t)HD, 1};
    if (int e = make_tmap_bf16(&tmV, p->vt, 3, dims, str, box, 128)) return e;
  }
  tmQx = tmQa;
  tmE = tmQa;
  if (EXT) {
    const int xc = EXT == 1 ? 32 : 64;
    uint64_t dims[3] = {(uint64_t)xc, (uint64_t)p->seq_pad, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)xc * 2, (uint64_t)p->seq_pad * xc * 2};
    uint32_t box[3] = {(uint32_t)xc, 128, 1};
    if (int e = make_tmap_bf16(&tmQx, p->qext, 3, dims, str, box, xc * 2)) return e;
    uint64_t edims[2] = {(uint64_t)xc, (uint64_t)(EXT == 1 ? 256 : 128)};
    uint64_t estr[1] = {(uint64_t)xc * 2};
    uint32_t ebox[2] = {(uint32_t)xc, 128};
    if (int e = make_tmap_bf16(&tmE, p->kext, 2, edims, estr, ebox, xc * 2)) return e;
  }
  AttnDev d{};
  d.out = static_cast<bf16*>(p->out);
  d.ldo = p->ldo;
  d.heads = p->heads;
  d.seq = p->seq;
  d.seq_pad = p->seq_pad;
  d.c1 = p->scale * LOG2E;
  d.causal = p->causal;
  d.kv_len = p->kv_len;
  d.row_bias = static_cast<const bf16*>(p->row_bias);
  d.out_row_map = p->out_row_map;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("LLMSEG_ATTN_DBG");
      dbg = e ? atoi(e) : 0;
    }
    d.dbg = dbg;
  }

  auto kern = attn_kernel<HD, EXT>;
  static bool attr_done = false;
  if (!attr_done) {
    LLMSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid((p->seq + BM - 1) / BM, BH);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr, 0);
  LLMSEG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQa, tmQb, tmKa, tmKb, tmV, tmQx, tmE, d));
  g_launches.fetch_add(1);
  return 0;
}

template <int HD, int EXT>
int launch_attn3(const llmseg_attn_params* p, cudaStream_t stream) {
  using C = A3Cfg<HD, EXT>;
  constexpr int SMEM = C::SMEM_BYTES;
  const int BH = p->batch * p->heads;
  CUtensorMap tmQa, tmQb, tmKa, tmKb, tmV, tmQx, tmE;
  {
    uint64_t dims[3] = {(uint64_t)HD, (uint64_t)p->seq_pad, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)HD * 2, (uint64_t)p->seq_pad * HD * 2};
    uint32_t qa[3] = {64, 128, 1}, ka[3] = {64, BN3, 1};
    if (int e = make_tmap_bf16(&tmQa, p->q, 3, dims, str, qa, 128)) return e;
    if (int e = make_tmap_bf16(&tmKa, p->k, 3, dims, str, ka, 128)) return e;
    tmQb = tmQa;
    tmKb = tmKa;
    if (HD == 80) {
      uint32_t qb[3] = {16, 128, 1}, kb[3] = {16, BN3, 1};
      if (int e = make_tmap_bf16(&tmQb, p->q, 3, dims, str, qb, 32)) return e;
      if (int e = make_tmap_bf16(&tmKb, p->k, 3, dims, str, kb, 32)) return e;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)p->seq_pad, (uint64_t)HD, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)p->seq_pad * 2, (uint64_t)p->seq_pad * HD * 2};
    uint32_t box[3] = {BN3, (uint32_t)HD, 1};
    if (int e = make_tmap_bf16(&tmV, p->vt, 3, dims, str, box, 128)) return e;
  }
  tmQx = tmQa;
  tmE = tmQa;
  if (EXT) {
    const int xc = EXT == 1 ? 32 : 64;
    uint64_t dims[3] = {(uint64_t)xc, (uint64_t)p->seq_pad, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)xc * 2, (uint64_t)p->seq_pad * xc * 2};
    uint32_t box[3] = {(uint32_t)xc, 128, 1};
    if (int e = make_tmap_bf16(&tmQx, p->qext, 3, dims, str, box, xc * 2)) return e;
    uint64_t edims[2] = {(uint64_t)xc, (uint64_t)(EXT == 1 ? 256 : 128)};
    uint64_t estr[1] = {(uint64_t)xc * 2};
    uint32_t ebox[2] = {(uint32_t)xc, (uint32_t)(EXT == 1 ? 256 : 64)};
    if (int e = make_tmap_bf16(&tmE, p->kext, 2, edims, estr, ebox, xc * 2)) return e;
  }
  AttnDev d{};
  d.out = static_cast<bf16*>(p->out);
  d.ldo = p->ldo;
  d.heads = p->heads;
  d.seq = p->seq;
  d.seq_pad = p->seq_pad;
  d.c1 = p->scale * LOG2E;
  d.causal = p->causal;
  d.kv_len = p->kv_len;
  d.row_bias = static_cast<const bf16*>(p->row_bias);
  d.out_row_map = p->out_row_map;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((p->seq + BM - 1) / BM, BH);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr, 0);
  auto kern = attn3_kernel<HD, EXT>;
  static bool attr_done = false;
  if (!attr_done) {
    LLMSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_done = true;
  }
  LLMSEG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQa, tmQb, tmKa, tmKb, tmV, tmQx, tmE, d));
  g_launches.fetch_add(1);
  return 0;
}

// Which kernel: measured on one box, interleaved (profiles/round2_attn.md) — the double-buffered 64-key kernel wins
// where the sequence is short or the head wide (LLaMA causal T=319: 36 -> 28 us, T=767: 33 -> 28, CLIP 20 -> 17,
// DINOv2 4097 tokens 880 -> 825 us); the SAM global layers (rel-pos extension columns: 13 instead of 9 score MMAs
// per 128 keys, and twice the softmax warps per SM to hide their latencies) stay 8-10 % faster on the 128-key kernel.
// LLMSEG_ATTN_V1 = 1 / 0 forces one of them (read per call: A/B runs flip it between launches).
// (Round 2 also measured v3's schedule with v1's warp count — two softmax threads per row sharing a reference maximum two
// tiles late, no per-tile barrier: correct, 967-1052 us against v1's 803-895 and v3's 894-964 on the SAM global shape.
// Halving the keys a thread handles per visit costs more than the added warps hide; profiles/round2_attn.md.)
template <int HD, int EXT>
int launch_attn_any(const llmseg_attn_params* p, cudaStream_t stream) {
  const char* e = getenv("LLMSEG_ATTN_V1");
  const bool v1 = e != nullptr ? atoi(e) == 1 : EXT != 0;
  return v1 ? launch_attn<HD, EXT>(p, stream) : launch_attn3<HD, EXT>(p, stream);
}

}  // namespace
}  // namespace llmseg

using namespace llmseg;

extern "C" int llmseg_attention(const llmseg_attn_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LLMSEG_REQUIRE(p != nullptr, LLMSEG_EARG, "llmseg_attention: null params");
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(p->q && p->k && p->vt && p->out, LLMSEG_EARG, "llmseg_attention: null q/k/vt/out");
  LLMSEG_REQUIRE(p->batch > 0 && p->heads > 0 && p->seq > 0 && p->seq_pad >= p->seq &&
                     p->seq_pad % 8 == 0,
                 LLMSEG_ESHAPE, "llmseg_attention: batch=%d heads=%d seq=%d seq_pad=%d", p->batch,
                 p->heads, p->seq, p->seq_pad);
  LLMSEG_REQUIRE(p->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0, LLMSEG_EALIGN,
                 "llmseg_attention: out / ldo not 16-byte aligned");
  LLMSEG_REQUIRE(p->scale > 0.f, LLMSEG_EARG, "llmseg_attention: scale must be positive");
  const int hd = p->head_dim, ext = p->ext_cols;
  if (ext != 0) {
    LLMSEG_REQUIRE(hd == 80 && (ext == 32 || ext == 64) && p->qext && p->kext && !p->causal,
                   LLMSEG_ESHAPE, "llmseg_attention: rel-pos extension needs head_dim 80, ext 32|64");
    LLMSEG_REQUIRE(ext == 32 ? (p->seq <= 256 && p->row_bias == nullptr)
                             : (p->row_bias != nullptr && p->seq <= 64 * 64),
                   LLMSEG_ESHAPE, "llmseg_attention: ext_cols=%d inconsistent with seq=%d / row_bias", ext,
                   p->seq);
  }
  if (hd == 64 && ext == 0) return launch_attn_any<64, 0>(p, stream);
  if (hd == 128 && ext == 0) return launch_attn_any<128, 0>(p, stream);
  if (hd == 80 && ext == 0) return launch_attn_any<80, 0>(p, stream);
  if (hd == 80 && ext == 32 && p->seq >= 192 && p->seq <= 208 && p->kv_len == nullptr && use_attn_win())
    return launch_attn_win(p, stream);
  if (hd == 80 && ext == 32) return launch_attn_any<80, 1>(p, stream);
  if (hd == 80 && ext == 64) return launch_attn_any<80, 2>(p, stream);
  return set_error(LLMSEG_ESHAPE, "llmseg_attention: unsupported head_dim=%d ext_cols=%d", hd, ext);
}
