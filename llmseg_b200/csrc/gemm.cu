// llmseg_b200 — persistent warp-specialised tcgen05 GEMM  C = epilogue(A · Wᵀ)  for sm_100a.
//
//   warp 0 (1 lane)  TMA producer: A[128×64] and W[BN×64] bf16 tiles, 128B-swizzled, mbarrier ring
//   warp 1 (1 lane)  MMA issuer:   tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16 ×4 per stage,
//                                  fp32 accumulators in TMEM, double-buffered (2×BN columns)
//   warp 2           TMEM allocator
//   warps 4..11      epilogue (2 per TMEM lane quarter): tcgen05.ld 32 lanes × 32 columns → bias / act /
//                    residual / layout → HBM
//
// The epilogue variants replace the reference's separate elementwise passes:
//   PLAIN   bias + {GELU, quick-GELU, ReLU} + residual (+ row scatter for window un-partition,
//           reference image_encoder.py:291-318, and broadcast residual for pos_embed, :111-113)
//   SWIGLU  LLaMA MLP  silu(gate)·up  with gate/up rows interleaved in W
//   QKV     split into per-head Q, K, Vᵀ (reference image_encoder.py:238-242) with optional
//           rotate-half RoPE on q,k (transformers LlamaAttention.apply_rotary_pos_emb)
#include <atomic>
#include <cstdlib>

#include <climits>
#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
int launch_relpos_win(const void* q, const void* rel_hw, int rows, int seq, int seq_pad, float inv_scale,
                      void* qext, cudaStream_t stream);  // relpos.cu

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle span
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmDev {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  bf16* C;
  int ldc;
  const bf16* bias;
  const bf16* residual;
  int ldr;
  int res_mod;
  int act;
  const int* out_row_map;
  bf16* q;
  bf16* k;
  bf16* vt;
  int heads, head_dim, seq_in, seq_pad;
  const bf16* rope_cos;
  const bf16* rope_sin;
  // RELPOS mode (llmseg_relpos_prep)
  int rp_grid, rp_seq, rp_seq_pad, rp_ext;
  float rp_inv_scale;
  bf16* rp_qext;
  bf16* rp_rh;
  int cluster;    // CTAs per cluster along M (1, 2 or 4)
  int n_fastest;  // work-item order (see kernel)
  // stream-K tail of the CTA-pair kernel (0 = off): the last, partial wave of tiles is cut along K into
  // one contiguous range of sk_w k-blocks per CTA pair; see gemm2_kernel.
  // LayerNorm / RMSNorm of the A rows folded into the epilogue (see include/llmseg_b200.h: row_stats)
  const float2* row_stats;
  int stats_parts_in;    // 0: row_stats = (mean, rstd) per row; > 0: that many (sum, sum of squares) partials per row
  float norm_inv_dim, norm_eps;
  int norm_rms;
  int tma_store;         // PLAIN, no row map: C leaves through shared memory + TMA tile stores (pair kernel)
  int tma_res;           // pair kernel, BN=256: the residual tile arrives through TMA into shared memory (the ring
                         // gives up its last stage for the 8 x 4 KB landing buffers), see epilogue_tile
  float2* stats_out;     // PLAIN: per-row (sum, sum of squares) partials of this GEMM's output, or null
  int stats_parts_out;   // partials per row = num_n_tiles * 2 (one per epilogue warp sharing a row)
  float2* stats_final;   // (mean, rstd) per output row, finished in-kernel by the last tile of a row block
  int* stats_counters;   // [num_m_tiles] arrivals per 128-row block (zero before and after the launch)
  float so_inv_dim, so_eps;
  int so_rms;
  int sk_w;
  float* sk_ws;   // fp32 partial tiles: [pair][cta rank][BN/4][128 rows] float4
  int* sk_flags;  // [pair][cta rank]: 1 = that pair's partial is complete
};

constexpr int MODE_RELPOS = 3;

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // rel-pos mode: K <= 128 needs only 2 stages; 4 x 16 KB fp32 staging tiles (one per TMEM lane quarter)
  static constexpr int RP_STAGES = 2;
  static constexpr int RP_STAGING = 4 * 128 * 32 * 4;
  static constexpr int RP_SMEM_BYTES = RP_STAGES * STAGE_BYTES + RP_STAGING + 1024 + 256;
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu2 (packed-FMA exact-erf GELU) lives in common.cuh: the proposal kernels (amg.cu) use it too
__device__ __forceinline__ float fast_sigmoid(float x) {
  return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x));
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == LLMSEG_ACT_QUICK_GELU) return x * fast_sigmoid(1.702f * x);
  if (act == LLMSEG_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

__device__ __forceinline__ uint4 ldg16(const bf16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// ---- epilogue: 32 fp32 accumulator columns of one row, PLAIN mode ----------------------------
// bias / residual vectors for the whole 32-column chunk are fetched by the caller BEFORE it waits for
// the TMEM load, so their latency overlaps it (they were serialised on the critical path: ncu showed
// 17 % of all warp samples on the first use of the bias registers).
struct EpiPrefetch {
  uint4 bias[4];
  uint4 res[4];
};
__device__ __forceinline__ void epi_prefetch(const GemmDev& p, EpiPrefetch& pf, int out_row, int n0, bool live) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = n0 + g * 8;
    pf.bias[g] = make_uint4(0, 0, 0, 0);
    pf.res[g] = make_uint4(0, 0, 0, 0);
    if (live && col < p.N) {
      if (p.bias) pf.bias[g] = ldg16(p.bias + col);
      if (p.residual && !p.tma_res) {
        const int rr = p.res_mod > 0 ? out_row % p.res_mod : out_row;
        pf.res[g] = ldg16(p.residual + (size_t)rr * p.ldr + col);
      }
    }
  }
}
// stage_row != nullptr: the packed bf16 groups go to this row of the warp's 128B-swizzled [32 x 64] staging
// tile (16-byte chunk index chunk0 + j/8, XOR row&7) instead of global memory; a TMA store follows.
// res_row != nullptr: the residual of this row's 32 columns sits in a 64B-swizzled [32 x 32] landing tile that a
// TMA load filled (16-byte chunk index j/8 XOR res_swz) instead of in pf.res.
// All arithmetic runs on packed fp32 pairs (FFMA2 / FADD2): acc * rstd + bias is ONE instruction per pair
// (rs = 1 without a folded norm, bias = 0 without a bias), residual one, the row statistics two — the epilogue
// of the short-K GEMMs is what the MMA warp waits for (profiles/r02e), so instructions here are kernel time.
__device__ __forceinline__ void epi_plain(const GemmDev& p, const uint32_t* r, float rs, int out_row, int n0,
                                          const EpiPrefetch& pf, float2& st_s, float2& st_ss,
                                          uint8_t* stage_row = nullptr, int chunk0 = 0, int swz = 0,
                                          const uint8_t* res_row = nullptr, int res_swz = 0) {
  const float2 rs2 = make_float2(rs, rs);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int col = n0 + j;
    if (col >= p.N) break;
    float2 v[4];
    {
      const uint4 b = pf.bias[j >> 3];  // zeros when there is no bias
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        v[e] = ffma2(make_float2(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1])), rs2,
                     unpack_bf16(bw[e]));
    }
    if (p.act == LLMSEG_ACT_GELU) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = gelu2(v[e]);
    } else if (p.act != LLMSEG_ACT_NONE) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = make_float2(apply_act(v[e].x, p.act), apply_act(v[e].y, p.act));
    }
    if (p.residual) {
      const uint4 b = res_row != nullptr ? *reinterpret_cast<const uint4*>(res_row + (((j >> 3) ^ res_swz) << 4))
                                         : pf.res[j >> 3];
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = fadd2(v[e], unpack_bf16(bw[e]));
    }
    if (p.stats_out != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        st_s = fadd2(st_s, v[e]);
        st_ss = ffma2(v[e], v[e], st_ss);
      }
    }
    uint4 o;
    o.x = pack_bf16(v[0].x, v[0].y);
    o.y = pack_bf16(v[1].x, v[1].y);
    o.z = pack_bf16(v[2].x, v[2].y);
    o.w = pack_bf16(v[3].x, v[3].y);
    if (stage_row != nullptr)
      *reinterpret_cast<uint4*>(stage_row + (((chunk0 + (j >> 3)) ^ swz) << 4)) = o;
    else
      *reinterpret_cast<uint4*>(p.C + (size_t)out_row * p.ldc + col) = o;
  }
}

// ---- SWIGLU: columns (2j, 2j+1) = (gate_j, up_j) ---------------------------------------------
__device__ __forceinline__ void epi_swiglu(const GemmDev& p, const uint32_t* r, int out_row, int n0) {
#pragma unroll
  for (int j = 0; j < 32; j += 16) {
    const int col = n0 + j;
    if (col >= p.N) break;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = __uint_as_float(r[j + 2 * e]);
      const float u = __uint_as_float(r[j + 2 * e + 1]);
      o[e] = g * fast_sigmoid(g) * u;
    }
    uint4 w;
    w.x = pack_bf16(o[0], o[1]);
    w.y = pack_bf16(o[2], o[3]);
    w.z = pack_bf16(o[4], o[5]);
    w.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(p.C + (size_t)out_row * p.ldc + (col >> 1)) = w;
  }
}

// ---- QKV split (no RoPE): 8-column groups never straddle a head (head_dim % 8 == 0) ----------
// acc * rs + bias on packed pairs (rs: folded-norm row scale, 1 without; pf.bias is zero without a bias): the epilogue
// of the K = 1280 QKV projection paces its mainloop (ncu r2n: 2207 instructions per warp and tile), so the scale, the
// bias and the conversion are one FFMA2 + one pack per pair, and the transposed V stores share one packed conversion.
__device__ __forceinline__ void epi_qkv(const GemmDev& p, const uint32_t* r, float rs, int row, int n0,
                                        const EpiPrefetch& pf) {
  const int b = row / p.seq_in;
  const int s = row - b * p.seq_in;
  const int hd = p.head_dim;
  const int hw = p.heads * hd;
  // (q|k|v, head, offset in head) of the chunk's first column by division, then advanced by 8 per group: the
  // runtime divisions (head_dim 80 is no power of two) were more instructions than the stores they steer
  int which = n0 / hw;
  const int rem0 = n0 - which * hw;
  int h = rem0 / hd;
  int d = rem0 - h * hd;
  const float2 rs2 = make_float2(rs, rs);
  const uint32_t sp = (uint32_t)p.seq_pad;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int col = n0 + j;
    if (col >= p.N) break;
    const uint4 bb = pf.bias[j >> 3];
    const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v = ffma2(make_float2(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1])), rs2,
                             unpack_bf16(bw[e]));
      o[e] = pack_bf16(v.x, v.y);
    }
    const size_t bh = (size_t)b * p.heads + h;
    if (which < 2) {
      bf16* dst = (which == 0 ? p.q : p.k) + (bh * p.seq_pad + s) * hd + d;
      *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.vt) + (bh * hd + d) * p.seq_pad + s;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dst[(2 * e) * sp] = (uint16_t)(o[e] & 0xffffu);
        dst[(2 * e + 1) * sp] = (uint16_t)(o[e] >> 16);
      }
    }
    d += 8;   // 8-column groups never straddle a head (head_dim % 8 == 0)
    if (d >= hd) {
      d = 0;
      if (++h == p.heads) {
        h = 0;
        ++which;
      }
    }
  }
}

// ---- rel-pos gather: QR[row, i] -> qext / row_bias (see include/llmseg_b200.h) --------------
// The QR tile (128 table columns) of the warp pair that shares a TMEM lane quarter is staged in shared
// memory as stage[col][lane] (fp32, conflict-free both ways); every thread then assembles its own
// output row with a per-row dynamic column index and writes it with 16-byte stores, so a warp writes
// one contiguous 2-4 KB block instead of 32-way scattered 2-byte stores.
//   table layout (ops.make_rel_hw): windows  rel_h rows [0,2G-1), rel_w rows [32,32+2G-1)   (1 N-tile)
//                                   global   rel_h rows [0,127),  rel_w rows [128,255)       (2 N-tiles)
__device__ __forceinline__ void relpos_tile(const GemmDev& p, float* stage, uint32_t taddr, int row,
                                            int n_blk, int quarter, int chalf, int lane, bool row_ok) {
  // 1. stage this warp's 64 accumulator columns (rounded to bf16 like the reference's einsum output)
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(taddr + chalf * 64 + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e)
      stage[(chalf * 64 + c * 32 + e) * 32 + lane] = bf16_round(__uint_as_float(r[e]));
  }
  asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
  // 2. assemble + store
  const int s = row % p.rp_seq_pad;
  const bool live = row_ok && s < p.rp_seq;
  const int G = p.rp_grid;
  const int qh = s / G, qw = s - qh * G;
  if (p.rp_rh == nullptr) {
    // windows: out[j<G] = QR[qh+G-1-j], out[G+j] = QR[32+qw+G-1-j], j<G; 16 outputs per warp half
    const int j0 = chalf * 16;
    float v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int j = j0 + t;
      float x = 0.f;
      if (j < G) x = stage[(qh + G - 1 - j) * 32 + lane];
      else if (j < 2 * G) x = stage[(32 + qw + 2 * G - 1 - j) * 32 + lane];
      v[t] = x * p.rp_inv_scale;
    }
    if (live) {
      uint4 o0, o1;
      o0.x = pack_bf16(v[0], v[1]); o0.y = pack_bf16(v[2], v[3]); o0.z = pack_bf16(v[4], v[5]); o0.w = pack_bf16(v[6], v[7]);
      o1.x = pack_bf16(v[8], v[9]); o1.y = pack_bf16(v[10], v[11]); o1.z = pack_bf16(v[12], v[13]); o1.w = pack_bf16(v[14], v[15]);
      uint4* dst = reinterpret_cast<uint4*>(p.rp_qext + (size_t)row * 32 + j0);
      dst[0] = o0;
      dst[1] = o1;
    }
  } else {
    // global: N-tile 0 -> row_bias[kh] = QR[qh+63-kh]; N-tile 1 -> qext[kw] = QR[qw+63-kw]/scale
    const int pos = n_blk == 0 ? qh : qw;
    const float sc = n_blk == 0 ? 1.0f : p.rp_inv_scale;
    bf16* dstp = (n_blk == 0 ? p.rp_rh : p.rp_qext) + (size_t)row * 64 + chalf * 32;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      float v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int j = chalf * 32 + g8 * 8 + t;
        const int col = pos + G - 1 - j;
        v[t] = (col >= 0 && col < 128 ? stage[col * 32 + lane] : 0.f) * sc;
      }
      if (live) {
        uint4 o;
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        reinterpret_cast<uint4*>(dstp)[g8] = o;
      }
    }
  }
  asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");  // staging buffer is reused by the next tile
}

// ---- QKV split with rotate-half RoPE, head_dim == 128: lo/hi are columns [d0,d0+32) and
//      [d0+64,d0+96) of the same head; d0 in {0,32}.  cos/sin tables are bf16 [seq, 64]. --------
__device__ __forceinline__ void epi_qkv_rope(const GemmDev& p, const uint32_t* lo, const uint32_t* hi,
                                             int row, int col_lo) {
  const int b = row / p.seq_in;
  const int s = row - b * p.seq_in;
  const int hw = p.heads * 128;
  const int which = col_lo / hw;
  const int rem = col_lo - which * hw;
  const int h = rem >> 7;
  const int d0 = rem & 127;  // 0 or 32
  const size_t bh = (size_t)b * p.heads + h;
  if (which == 2) {
    bf16* dst = p.vt + (bh * 128 + d0) * p.seq_pad + s;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      dst[(size_t)e * p.seq_pad] = __float2bfloat16_rn(__uint_as_float(lo[e]));
      dst[(size_t)(e + 64) * p.seq_pad] = __float2bfloat16_rn(__uint_as_float(hi[e]));
    }
    return;
  }
  bf16* dst = (which == 0 ? p.q : p.k) + (bh * p.seq_pad + s) * 128 + d0;
  const bf16* cs = p.rope_cos + (size_t)s * 64 + d0;
  const bf16* sn = p.rope_sin + (size_t)s * 64 + d0;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 cw = ldg16(cs + j), sw = ldg16(sn + j);
    const uint32_t c4[4] = {cw.x, cw.y, cw.z, cw.w};
    const uint32_t s4[4] = {sw.x, sw.y, sw.z, sw.w};
    float ol[8], oh[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 c = unpack_bf16(c4[e]);
      const float2 sn2 = unpack_bf16(s4[e]);
      const float cc[2] = {c.x, c.y}, ss[2] = {sn2.x, sn2.y};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float x1 = __uint_as_float(lo[j + 2 * e + t]);
        const float x2 = __uint_as_float(hi[j + 2 * e + t]);
        // q*cos + rotate_half(q)*sin in fp32, rounded once at the store
        ol[2 * e + t] = fmaf(x1, cc[t], -x2 * ss[t]);
        oh[2 * e + t] = fmaf(x2, cc[t], x1 * ss[t]);
      }
    }
    uint4 o;
    o.x = pack_bf16(ol[0], ol[1]);
    o.y = pack_bf16(ol[2], ol[3]);
    o.z = pack_bf16(ol[4], ol[5]);
    o.w = pack_bf16(ol[6], ol[7]);
    *reinterpret_cast<uint4*>(dst + j) = o;
    o.x = pack_bf16(oh[0], oh[1]);
    o.y = pack_bf16(oh[2], oh[3]);
    o.z = pack_bf16(oh[4], oh[5]);
    o.w = pack_bf16(oh[6], oh[7]);
    *reinterpret_cast<uint4*>(dst + 64 + j) = o;
  }
}

// ---- stream-K tail helpers (CTA-pair kernel) --------------------------------------------------
// Work of one CTA pair = `full` whole tiles (grp = pair, pair + n_pairs, ...) followed, when sk_w > 0, by
// the k-block range [pair*sk_w, (pair+1)*sk_w) of the linearised (tail tile, k-block) space.  A range
// touches at most two tiles (sk_w < num_k_blocks): the end of one tile (a PARTIAL segment: its fp32
// accumulators go to the workspace) and the start of the next (the OWNER segment: it adds the partials of
// the pairs that follow and runs the real epilogue).  Producers never wait, owners wait only on
// higher-numbered pairs, so the scheme cannot deadlock however the grid is scheduled.
struct SkSeg {
  int grp, k0, k1;
  int n_peers;  // owner: number of partial segments to fold in (pairs pair+1 .. pair+n_peers)
};
struct SkWalk {
  int full_end, pos, end, kb, w;
  __device__ SkWalk(const GemmDev& p, int num_groups, int pair, int n_pairs) {
    kb = p.num_k_blocks;
    w = p.sk_w;
    if (w > 0) {
      full_end = (num_groups / n_pairs) * n_pairs;
      const int total = (num_groups - full_end) * kb;
      pos = pair * w < total ? pair * w : total;
      end = pos + w < total ? pos + w : total;
    } else {
      full_end = num_groups;
      pos = end = 0;
    }
  }
  // next tail segment of this pair, false when done
  __device__ bool next(SkSeg& sgm) {
    if (pos >= end) return false;
    const int t = pos / kb;
    sgm.grp = full_end + t;
    sgm.k0 = pos - t * kb;
    const int left = end - pos;
    sgm.k1 = sgm.k0 + left < kb ? sgm.k0 + left : kb;
    sgm.n_peers = 0;
    if (sgm.k0 == 0 && sgm.k1 < kb) sgm.n_peers = ((t + 1) * kb - 1) / w - pos / w;
    pos += sgm.k1 - sgm.k0;
    return true;
  }
};
__device__ __forceinline__ float* sk_slot(const GemmDev& p, int bn, int pair, int cta_rank) {
  return p.sk_ws + ((size_t)pair * 2 + cta_rank) * (size_t)(128 * bn);
}
// acc[0..32) += partial columns [c0, c0+32) of this thread's row, for every peer
__device__ __forceinline__ void sk_accumulate(const GemmDev& p, int bn, uint32_t* r, int c0, int row_in_cta,
                                              int pair, int cta_rank, int n_peers) {
  for (int pr = 1; pr <= n_peers; ++pr) {
    const float4* ws = reinterpret_cast<const float4*>(sk_slot(p, bn, pair + pr, cta_rank)) +
                       (size_t)(c0 >> 2) * 128 + row_in_cta;
    float4 f[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) f[v] = __ldcg(ws + v * 128);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      r[4 * v + 0] = __float_as_uint(__uint_as_float(r[4 * v + 0]) + f[v].x);
      r[4 * v + 1] = __float_as_uint(__uint_as_float(r[4 * v + 1]) + f[v].y);
      r[4 * v + 2] = __float_as_uint(__uint_as_float(r[4 * v + 2]) + f[v].z);
      r[4 * v + 3] = __float_as_uint(__uint_as_float(r[4 * v + 3]) + f[v].w);
    }
  }
}
// partial segment: this warp's share (32 rows x half the columns) of the fp32 accumulators -> workspace
template <int BN, int NP>
__device__ __forceinline__ void sk_dump(const GemmDev& p, uint32_t taddr, int quarter, int chalf, int lane,
                                        int pair, int cta_rank) {
  float4* ws = reinterpret_cast<float4*>(sk_slot(p, BN, pair, cta_rank)) + quarter * 32 + lane;
#pragma unroll 1
  for (int c = chalf * (BN / 32 / NP); c < (chalf + 1) * (BN / 32 / NP); ++c) {
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int v = 0; v < 8; ++v)
      ws[(size_t)(c * 8 + v) * 128] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                  __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
  }
}
// Row normalisation of A folded into the GEMM: acc is x·W''ᵀ for the un-normalised rows x, and
// Norm(x)·Wᵀ = rstd · acc (+ bias') — see include/llmseg_b200.h (row_stats).
__device__ __forceinline__ void row_scale(uint32_t* r, float rs) {
#pragma unroll
  for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * rs);
}

struct SkOwner {
  int pair, cta_rank, n_peers;  // n_peers == 0: ordinary tile
};

// rstd of A row `row` for the folded norm: either stored directly or reduced from the (sum, sum of squares)
// partials the producing GEMM's epilogue wrote (stats_out).  Called BEFORE the wait for the accumulators.
__device__ __forceinline__ float row_rstd(const GemmDev& p, int row) {
  if (p.row_stats == nullptr || row >= p.M) return 1.f;
  if (p.stats_parts_in == 0) return __ldg(p.row_stats + row).y;
  const float2* st = p.row_stats + (size_t)row * p.stats_parts_in;
  float s = 0.f, ss = 0.f;
  for (int i0 = 0; i0 < p.stats_parts_in; i0 += 8) {  // 8 independent loads per round trip
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = i0 + j < p.stats_parts_in ? __ldcg(st + i0 + j) : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s += v[j].x;
      ss += v[j].y;
    }
  }
  const float mean = p.norm_rms ? 0.f : s * p.norm_inv_dim;
  const float var = fmaxf(fmaf(-mean, mean, ss * p.norm_inv_dim), 0.f);
  return rsqrtf(var + p.norm_eps);
}

// After a tile's epilogue (all 8 epilogue warps of the CTA): count the tile in for its 128-row block; the
// CTA that brings the count to num_n_tiles reduces the block's partials — in index order, so the result
// does not depend on which CTA that is — into (mean, rstd) and re-arms the counter.
__device__ __forceinline__ void stats_finish(const GemmDev& p, int m_blk, int warp, int lane, volatile int* s_flag) {
  __threadfence();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (warp == 4 && lane == 0) {
    const int old = atomicAdd(p.stats_counters + m_blk, 1);
    const int last = old == p.num_n_tiles - 1;
    if (last) p.stats_counters[m_blk] = 0;
    *s_flag = last;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (*s_flag) {
    __threadfence();
    const int t = (warp - 4) * 32 + lane;
    const int row = m_blk * BM + t;
    if (t < BM && row < p.M) {
      const float2* st = p.stats_out + (size_t)row * p.stats_parts_out;
      float s = 0.f, ss = 0.f;
      for (int i0 = 0; i0 < p.stats_parts_out; i0 += 8) {
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = i0 + j < p.stats_parts_out ? __ldcg(st + i0 + j) : make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s += v[j].x;
          ss += v[j].y;
        }
      }
      const float mean = p.so_rms ? 0.f : s * p.so_inv_dim;
      const float var = fmaxf(fmaf(-mean, mean, ss * p.so_inv_dim), 0.f);
      p.stats_final[row] = make_float2(mean, rsqrtf(var + p.so_eps));
    }
  }
}

// first residual row of an epilogue warp's 32-row block (res_mod: the residual is a res_mod-row table that
// repeats down the output; the launcher admits tma_res only when res_mod % 32 == 0, so a block never wraps)
__device__ __forceinline__ int res_row0(const GemmDev& p, int m_blk, int quarter) {
  const int r = m_blk * BM + quarter * 32;
  return p.res_mod > 0 ? r % p.res_mod : r;
}

// tma_res: start the loads of a tile's first two residual blocks of this warp (called BEFORE the wait for the
// tile's accumulators; the landing buffers are free since the previous tile's last reads).
template <int BN, int NP>
__device__ __forceinline__ void res_prefetch_first(const GemmDev& p, uint8_t* res_stage, uint64_t* res_bar,
                                                   const CUtensorMap* tmR, int m_blk, int n_blk, int quarter,
                                                   int chalf, int lane) {
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int n0b = n_blk * BN + (chalf * (BN / 32 / NP) + b) * 32;
      if (n0b < p.N) {
        mbar_expect_tx(&res_bar[b], 2048);
        tma_load_2d(res_stage + b * 2048, tmR, &res_bar[b], n0b, res_row0(p, m_blk, quarter));
      }
    }
  }
  __syncwarp();
}

// One warp's share of a finished 128 x BN accumulator tile: quarter = TMEM lane quarter (32 rows),
// chalf = which of the NP column parts of the tile (NP warps share a lane quarter).
template <int BN, int MODE, bool ROPE, int NP = 2>
__device__ __forceinline__ void epilogue_tile(const GemmDev& p, float* rp_stage, uint32_t taddr, int m_blk,
                                              int n_blk, int quarter, int chalf, int lane, float rs,
                                              const SkOwner sk = SkOwner{0, 0, 0}, uint8_t* epi_stage = nullptr,
                                              const CUtensorMap* tmC = nullptr, uint8_t* res_stage = nullptr,
                                              uint64_t* res_bar = nullptr, const CUtensorMap* tmR = nullptr,
                                              uint32_t* res_ph = nullptr, int out_row_pre = INT_MIN) {
  const int row = m_blk * BM + quarter * 32 + lane;
  int out_row = row;
  if (out_row_pre != INT_MIN)
    out_row = out_row_pre;         // the caller fetched the row map before it waited for the accumulators
  else if ((MODE == LLMSEG_GEMM_PLAIN || MODE == LLMSEG_GEMM_QKV) && p.out_row_map != nullptr && row < p.M)
    out_row = p.out_row_map[row];  // QKV: position of this token in the (sequence, slot) index space
  const bool live = row < p.M && out_row >= 0;
  const bool fold = MODE != MODE_RELPOS && p.row_stats != nullptr;
  float2 st_s = make_float2(0.f, 0.f), st_ss = make_float2(0.f, 0.f);
  if (MODE == MODE_RELPOS) {
    relpos_tile(p, rp_stage + quarter * (128 * 32), taddr, row, n_blk, quarter, chalf, lane, row < p.M);
  } else if (MODE == LLMSEG_GEMM_QKV && ROPE) {
    // (128-column group c, 32-column offset `half`) pairs: BN/64 of them, dealt out to the NP warps
#pragma unroll 1
    for (int u = chalf; u < BN / 64; u += NP) {
      const int c = u >> 1;
      const int half = u & 1;
      uint32_t lo[32], hi[32];
      tmem_ld32(taddr + c * 128 + half * 32, lo);
      tmem_ld32(taddr + c * 128 + half * 32 + 64, hi);
      tmem_ld_wait();
      if (sk.n_peers > 0) {
        sk_accumulate(p, BN, lo, c * 128 + half * 32, quarter * 32 + lane, sk.pair, sk.cta_rank, sk.n_peers);
        sk_accumulate(p, BN, hi, c * 128 + half * 32 + 64, quarter * 32 + lane, sk.pair, sk.cta_rank, sk.n_peers);
      }
      const int col = n_blk * BN + c * 128 + half * 32;
      if (fold) {
        row_scale(lo, rs);
        row_scale(hi, rs);
      }
      if (live && col < p.N) epi_qkv_rope(p, lo, hi, out_row, col);
    }
  } else if (MODE == LLMSEG_GEMM_SWIGLU || MODE == LLMSEG_GEMM_QKV) {
    // SwiGLU / Q-K-V split: straight-line over the warp's chunks, TMEM read one chunk ahead (the next 32 columns are
    // in flight while this chunk is transformed and stored) — same restructuring as epilogue_plain_fast
    constexpr int NC = BN / 32 / NP;
    const int c0 = chalf * NC;
    uint32_t r[2][32];
    tmem_ld32(taddr + c0 * 32, r[0]);
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
      const int c = c0 + ci;
      const int n0 = n_blk * BN + c * 32;
      EpiPrefetch pf;
      if (MODE == LLMSEG_GEMM_QKV) epi_prefetch(p, pf, 0, n0, live);
      tmem_ld_wait();
      if (ci + 1 < NC) tmem_ld32(taddr + (c + 1) * 32, r[(ci + 1) & 1]);
      if (sk.n_peers > 0) sk_accumulate(p, BN, r[ci & 1], c * 32, quarter * 32 + lane, sk.pair, sk.cta_rank, sk.n_peers);
      if (fold && MODE == LLMSEG_GEMM_SWIGLU) row_scale(r[ci & 1], rs);   // QKV folds the scale into its bias FFMA2
      if (live && n0 < p.N) {
        if (MODE == LLMSEG_GEMM_SWIGLU) epi_swiglu(p, r[ci & 1], out_row, n0);
        else epi_qkv(p, r[ci & 1], fold ? rs : 1.f, out_row, n0, pf);
      }
    }
  } else {
    // Residual through TMA (pair kernel): the warp's 32-row x 32-column residual blocks land in two 2 KB
    // 64B-swizzled buffers, two blocks ahead of their use; ncu had the epilogue warps parked ~30 % of their
    // time on the first use of per-lane residual loads (32 rows per request), profiles/r02c.
    constexpr int C_FIRST_STRIDE = BN / 32 / NP;
    const bool tres = MODE == LLMSEG_GEMM_PLAIN && res_stage != nullptr && p.tma_res;
#pragma unroll 1
    for (int c = chalf * (BN / 32 / NP); c < (chalf + 1) * (BN / 32 / NP); ++c) {
      uint32_t r[32];
      const int n0 = n_blk * BN + c * 32;
      const int ci = c - chalf * C_FIRST_STRIDE;
      tmem_ld32(taddr + c * 32, r);
      EpiPrefetch pf;
      if (MODE != LLMSEG_GEMM_SWIGLU) epi_prefetch(p, pf, MODE == LLMSEG_GEMM_PLAIN ? out_row : 0, n0, live);
      tmem_ld_wait();
      if (sk.n_peers > 0) sk_accumulate(p, BN, r, c * 32, quarter * 32 + lane, sk.pair, sk.cta_rank, sk.n_peers);
      if (fold && MODE == LLMSEG_GEMM_SWIGLU) row_scale(r, rs);  // PLAIN / QKV fold the scale into their bias FFMA2
      const bool staged = MODE == LLMSEG_GEMM_PLAIN && epi_stage != nullptr && p.tma_store;
      if (tres && n0 < p.N) {
        mbar_wait(&res_bar[ci & 1], (*res_ph >> (ci & 1)) & 1u);
        *res_ph ^= 1u << (ci & 1);
      }
      if (staged && (c & 1) == 0) {
        // the previous tile store of this warp must have drained the staging tile before it is rewritten
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
      if (live && n0 < p.N) {
        if (MODE == LLMSEG_GEMM_PLAIN)
          epi_plain(p, r, fold ? rs : 1.f, out_row, n0, pf, st_s, st_ss, staged ? epi_stage + lane * 128 : nullptr, (c & 1) * 4,
                    lane & 7, tres ? res_stage + (ci & 1) * 2048 + lane * 64 : nullptr, (lane >> 1) & 3);
        else if (MODE == LLMSEG_GEMM_SWIGLU) epi_swiglu(p, r, out_row, n0);
        else epi_qkv(p, r, fold ? rs : 1.f, out_row, n0, pf);
      }
      if (tres) {
        // every lane has consumed its row of this landing buffer: refill it with the block two steps ahead
        __syncwarp();
        const int n0n = n0 + 64;
        if (lane == 0 && ci + 2 < C_FIRST_STRIDE && n0n < p.N) {
          mbar_expect_tx(&res_bar[ci & 1], 2048);
          tma_load_2d(res_stage + (ci & 1) * 2048, tmR, &res_bar[ci & 1], n0n, res_row0(p, m_blk, quarter));
        }
      }
      if (staged && (c & 1) == 1) {
        // 64 columns x 32 rows staged: hand them to the TMA (full 128-byte lines; rows >= M and
        // columns >= N are clipped by the tensor map)
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmC, epi_stage, n_blk * BN + (c - 1) * 32, m_blk * BM + quarter * 32);
          bulk_commit();
        }
      }
    }
    if (MODE == LLMSEG_GEMM_PLAIN && p.stats_out != nullptr && live)
      p.stats_out[(size_t)out_row * p.stats_parts_out + n_blk * NP + chalf] = make_float2(st_s.x + st_s.y, st_ss.x + st_ss.y);
  }
}

// ---------------------------------------------------------------------------------------------
// Compile-time variants of the CTA-pair kernel's PLAIN epilogue.  ncu on the SAM lin1 GEMM (profiles/r02_summary.md,
// r02e) had 8.6 % of all samples on instruction-cache misses and 7.9 % on branch resolution inside the one epilogue
// body that serves every (bias, activation, residual, statistics, staging) combination, with the MMA warp waiting for
// `tmem_empty` 13 % of its time.  The combinations that carry the step get their own straight-line body:
//   1  bias + GELU, folded norm, staged TMA store                      SAM / DINOv2 MLP lin1
//   2  bias + TMA residual + row statistics, staged TMA store          SAM / DINOv2 proj and lin2 (folded norms next)
//   3  TMA residual, no bias, staged TMA store                         LLaMA o_proj / down_proj, CLIP-less text branch
//   4  bias + TMA residual, staged TMA store                           SAM mask-decoder projections (proposal generation)
//   5  bias alone, staged TMA store                                    mask-decoder ConvTranspose #1 as a GEMM
//   0  generic (every option a runtime test) — everything else
// The variants also read TMEM one chunk ahead (the next 32 columns are in flight while this chunk is worked on).
// ---------------------------------------------------------------------------------------------
template <int EPI>
struct EpiX {
  static constexpr bool bias = EPI == 1 || EPI == 2 || EPI == 4 || EPI == 5;
  static constexpr bool gelu = EPI == 1;
  static constexpr bool res = EPI == 2 || EPI == 3 || EPI == 4;
  static constexpr bool stats = EPI == 2;
};

template <int EPI>
__device__ __forceinline__ void epi_plain_x(const uint32_t* r, float rs, const uint4* bias4, float2& st_s, float2& st_ss,
                                            uint8_t* stage_row, int chunk0, int swz, const uint8_t* res_row, int res_swz) {
  using X = EpiX<EPI>;
  const float2 rs2 = make_float2(rs, rs);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    float2 v[4];
    if (X::bias) {
      const uint4 b = bias4[j >> 3];
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        v[e] = ffma2(make_float2(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1])), rs2, unpack_bf16(bw[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = make_float2(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1]));
    }
    if (X::gelu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = gelu2(v[e]);
    }
    if (X::res) {
      const uint4 b = *reinterpret_cast<const uint4*>(res_row + (((j >> 3) ^ res_swz) << 4));
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = fadd2(v[e], unpack_bf16(bw[e]));
    }
    if (X::stats) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        st_s = fadd2(st_s, v[e]);
        st_ss = ffma2(v[e], v[e], st_ss);
      }
    }
    uint4 o;
    o.x = pack_bf16(v[0].x, v[0].y);
    o.y = pack_bf16(v[1].x, v[1].y);
    o.z = pack_bf16(v[2].x, v[2].y);
    o.w = pack_bf16(v[3].x, v[3].y);
    *reinterpret_cast<uint4*>(stage_row + (((chunk0 + (j >> 3)) ^ swz) << 4)) = o;
  }
}

// one epilogue warp's share (32 rows x BN/2 columns) of a finished tile, variants 1-5 (pair kernel, BN = 256,
// N % 64 == 0, rows not scattered, C through staged TMA stores, residual through the TMA landing buffers)
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_plain_fast(const GemmDev& p, uint32_t taddr, int m_blk, int n_blk, int quarter,
                                                    int chalf, int lane, float rs, const SkOwner sk, uint8_t* epi_stage,
                                                    const CUtensorMap* tmC, uint8_t* res_stage, uint64_t* res_bar,
                                                    const CUtensorMap* tmR, uint32_t* res_ph) {
  using X = EpiX<EPI>;
  constexpr int NC = BN / 32 / 2;   // 32-column chunks per warp
  const int row = m_blk * BM + quarter * 32 + lane;
  const bool live = row < p.M;
  const int c0 = chalf * NC;
  float2 st_s = make_float2(0.f, 0.f), st_ss = make_float2(0.f, 0.f);
  uint32_t r[2][32];
  tmem_ld32(taddr + c0 * 32, r[0]);
#pragma unroll
  for (int ci = 0; ci < NC; ++ci) {
    const int c = c0 + ci;
    const int n0 = n_blk * BN + c * 32;
    uint4 bias4[4];
    if (X::bias) {
#pragma unroll
      for (int g = 0; g < 4; ++g) bias4[g] = n0 + g * 8 < p.N ? ldg16(p.bias + n0 + g * 8) : make_uint4(0, 0, 0, 0);
    }
    tmem_ld_wait();
    if (ci + 1 < NC) tmem_ld32(taddr + (c + 1) * 32, r[(ci + 1) & 1]);   // next chunk in flight under this one
    if (sk.n_peers > 0) sk_accumulate(p, BN, r[ci & 1], c * 32, quarter * 32 + lane, sk.pair, sk.cta_rank, sk.n_peers);
    if (X::res && n0 < p.N) {
      mbar_wait(&res_bar[ci & 1], (*res_ph >> (ci & 1)) & 1u);
      *res_ph ^= 1u << (ci & 1);
    }
    if ((c & 1) == 0) {
      // the previous tile store of this warp must have drained the staging tile before it is rewritten
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
    }
    if (live && n0 < p.N)
      epi_plain_x<EPI>(r[ci & 1], rs, bias4, st_s, st_ss, epi_stage + lane * 128, (c & 1) * 4, lane & 7,
                       res_stage + (ci & 1) * 2048 + lane * 64, (lane >> 1) & 3);
    if (X::res) {
      // every lane has consumed its row of this landing buffer: refill it with the block two steps ahead
      __syncwarp();
      const int n0n = n0 + 64;
      if (lane == 0 && ci + 2 < NC && n0n < p.N) {
        mbar_expect_tx(&res_bar[ci & 1], 2048);
        tma_load_2d(res_stage + (ci & 1) * 2048, tmR, &res_bar[ci & 1], n0n, res_row0(p, m_blk, quarter));
      }
    }
    if ((c & 1) == 1) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, epi_stage, n_blk * BN + (c - 1) * 32, m_blk * BM + quarter * 32);
        bulk_commit();
      }
    }
  }
  if (X::stats && live)
    p.stats_out[(size_t)row * p.stats_parts_out + n_blk * 2 + chalf] = make_float2(st_s.x + st_s.y, st_ss.x + st_ss.y);
}

template <int BN, int MODE, bool ROPE>
__global__ void __launch_bounds__(384, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmDev p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int NST = MODE == MODE_RELPOS ? C::RP_STAGES : C::STAGES;
  float* rp_stage = reinterpret_cast<float*>(smem + NST * C::STAGE_BYTES);  // RELPOS only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NST * C::STAGE_BYTES +
                                                   (MODE == MODE_RELPOS ? C::RP_STAGING : 0));
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tmem_full = empty_bar + NST;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  // Thread-block cluster of `cl` CTAs along M: the CTAs of a cluster work on `cl` vertically adjacent
  // output tiles that share the same W tile; each CTA TMA-loads 1/cl of it and multicasts the slice to
  // every CTA of the cluster, cutting L2->SM operand traffic from (BM+BN) to (BM+BN/cl) rows per k-block.
  const int cl = p.cluster;
  const int cta_rank = cl > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cl_mask = (uint16_t)((1u << cl) - 1u);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], cl);  // one tcgen05.commit arrival from every CTA that reads the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();  // peers' barriers must be initialised before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // prologue done; everything below may read what the previous kernel wrote

  // work items are groups of `cl` M-tiles x one N-tile, M-groups fastest; cluster c takes groups
  // c, c + num_clusters, ...  (all CTAs of a cluster iterate the same groups in lockstep)
  // group order keeps the smaller operand L2-resident: N fastest when A is the big (streamed-once)
  // operand (SAM: M >> N), M fastest when the weights are (LLaMA: N >> M).
  const int m_groups = (p.num_m_tiles + cl - 1) / cl;
  const int num_groups = m_groups * p.num_n_tiles;
  const bool n_fastest = p.n_fastest != 0;
  const int cluster_id = blockIdx.x / cl;
  const int num_clusters = gridDim.x / cl;
  const int slice_rows = BN / cl;

  // producer / MMA warps: all 32 lanes walk the schedule, one elected lane issues (see gemm2_kernel)
  if (warp == 0) {
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
        const int m_blk = (n_fastest ? grp / p.num_n_tiles : grp % m_groups) * cl + cta_rank;
        const int n_blk = n_fastest ? grp % p.num_n_tiles : grp / m_groups;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (issuer) {
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
            if (cl == 1) {
              tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
              tma_load_2d_mcast(sb + cta_rank * slice_rows * 128, &tmB, &full_bar[stage], kb * BK,
                                n_blk * BN + cta_rank * slice_rows, cl_mask);
            }
          }
          __syncwarp();
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint64_t da0 = umma_smem_desc(sa, 1024, UMMA_SW128);
          const uint64_t db0 = umma_smem_desc(sa + A_STAGE_BYTES, 1024, UMMA_SW128);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            if (cl == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_mcast(&empty_bar[stage], cl_mask);  // frees the slot in every CTA of the cluster
            if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: two per TMEM lane quarter, each taking half of the tile's columns
    const int quarter = warp & 3;
    const int chalf = (warp - 4) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
      const int m_blk = (n_fastest ? grp / p.num_n_tiles : grp % m_groups) * cl + cta_rank;
      const int n_blk = n_fastest ? grp % p.num_n_tiles : grp / m_groups;
      const float rs = MODE == MODE_RELPOS ? 1.f : row_rstd(p, m_blk * BM + quarter * 32 + lane);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
      epilogue_tile<BN, MODE, ROPE>(p, rp_stage, taddr, m_blk, n_blk, quarter, chalf, lane, rs);
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
      if (MODE == LLMSEG_GEMM_PLAIN && p.stats_final != nullptr)
        stats_finish(p, m_blk, warp, lane, reinterpret_cast<volatile int*>(tmem_ptr + 1));
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  Each CTA
// TMA-loads its own 128 A rows and HALF of the W tile; one thread of the leader CTA issues
// M=256 MMAs that read both CTAs' shared memory and write 128 accumulator rows into each CTA's TMEM.
// Per CTA a stage is 16 KB (A) + BN/2*128 B (W half) = 32 KB at BN=256: 6 stages instead of 4 in the
// same 192 KB, and the shared-memory operand traffic per MMA drops from 12 KB to 8 KB per CTA — the
// single-CTA kernel sits at 50 % tensor-pipe with its smem operand path and L2->SM path both half
// used (profiles/r01a_ncu_summary.md): it is latency-bound on a 4-stage ring.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct Cfg2 {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
  static constexpr int TMEM_COLS = 2 * BN;
  // epilogue staging: one [32 rows x 64 cols] bf16 tile (4 KB, 128B-swizzled) per epilogue warp
  static constexpr int EPI_TILE_BYTES = 32 * 128;
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + 8 * EPI_TILE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 1024 + 512;
  // tma_res: the ring runs STAGES-1 deep and the last stage's 32 KB hold the residual landing buffers
  // (two 2 KB [32 x 32] tiles per epilogue warp)
  static constexpr int OFF_RES = (STAGES - 1) * STAGE_BYTES;
};

// Epilogue warps per CTA (G2_EPI_WARPS/4 per TMEM lane quarter, each BN*4/G2_EPI_WARPS columns).  16 warps
// (640 threads, 96 registers each) measured 8 % SLOWER than 8 on the SAM lin1 shape: the epilogue spills
// and the extra warps do not buy latency hiding that matters.
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = 128 + G2_EPI_WARPS * 32;
template <int BN, int MODE, bool ROPE, int EPI = 0>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmDev p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* stats_bar = tmem_empty + 3;  // [4]: epilogue warps -> statistics warp, one phase per tile
  uint64_t* res_bar = stats_bar + 4;     // [2 per epilogue warp]: residual landing buffers (tma_res)
  const int nst = p.tma_res ? C::STAGES - 1 : C::STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int cta_rank = (int)cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's copy is the one in use: 1 arrive (expect_tx) + both CTAs' bytes
      mbar_init(&empty_bar[i], 1);  // one multicast commit per consumed stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * G2_EPI_WARPS);  // leader's copy: epilogue warps x 2 CTAs
    }
    for (int i = 0; i < 4; ++i) mbar_init(&stats_bar[i], G2_EPI_WARPS);
    for (int i = 0; i < 2 * G2_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // prologue done; everything below may read what the previous kernel wrote

  const int m_groups = (p.num_m_tiles + 1) / 2;
  const int num_groups = m_groups * p.num_n_tiles;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const bool n_fastest = p.n_fastest != 0;

  // Producer and MMA warps walk their schedules with all 32 lanes (uniform control flow: addresses,
  // descriptors and loop state live in uniform registers) and only the tcgen05 / TMA instructions sit
  // under an elected-lane predicate.  With the whole loop inside `if (lane == 0)` ptxas treated every
  // operand as divergent and wrapped each UTCHMMA in a ~20-instruction ELECT / R2UR loop: ~125
  // instructions per k-block kept the tensor pipe waiting on its own issuer (ncu: 843 clk per k-block for
  // 512 clk of MMA work, profiles/r01k_gemm_issue.md).
  if (warp == 0) {
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      SkWalk walk(p, num_groups, cluster_id, num_clusters);
      SkSeg sg;
      for (int grp = cluster_id;; grp += num_clusters) {
        int k0 = 0, k1 = p.num_k_blocks;
        int g = grp;
        if (grp >= walk.full_end) {
          if (!walk.next(sg)) break;
          g = sg.grp; k0 = sg.k0; k1 = sg.k1;
        }
        const int m_blk = (n_fastest ? g / p.num_n_tiles : g % m_groups) * 2 + cta_rank;
        const int n_blk = n_fastest ? g % p.num_n_tiles : g / m_groups;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (issuer) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmA, lbar, kb * BK, m_blk * BM);
            tma_load_2d_pair(sb, &tmB, lbar, kb * BK, n_blk * BN + cta_rank * (BN / 2));
          }
          __syncwarp();
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      const bool issuer = elect_one();  // one lane issues every MMA and commit (commits track their issuer)
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      SkWalk walk(p, num_groups, cluster_id, num_clusters);
      SkSeg sg;
      const uint32_t smem_base = smem_u32(smem);
      for (int grp = cluster_id;; grp += num_clusters) {
        int k0 = 0, k1 = p.num_k_blocks;
        if (grp >= walk.full_end) {
          if (!walk.next(sg)) break;
          k0 = sg.k0; k1 = sg.k1;
        }
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          // descriptors of the stage's first 16-column slice; the next slices are +32 bytes = +2 in the
          // (address >> 4) field
          const uint64_t da0 = umma_smem_desc(sa, 1024, UMMA_SW128);
          const uint64_t db0 = umma_smem_desc(sa + A_STAGE_BYTES, 1024, UMMA_SW128);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma2_ss(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, (kb != k0 || k != 0) ? 1u : 0u);
            umma2_commit_mcast(&empty_bar[stage], 3);  // frees the slot in both CTAs
            if (kb == k1 - 1) umma2_commit_mcast(&tmem_full[as], 3);
          }
          __syncwarp();
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ---- statistics warp: finishes the per-row (sum, sum of squares) partials off the critical path.  It
    //      walks the same schedule; for every finished tile it counts the tile in for its 128-row block and,
    //      when it was the last of the block's N-tiles (any CTA), reduces the partials — in index order, so
    //      the result does not depend on who is last — into (mean, rstd) and re-arms the counter. ----
    if (MODE == LLMSEG_GEMM_PLAIN && p.stats_final != nullptr) {
      int st_it = 0;
      SkWalk walk(p, num_groups, cluster_id, num_clusters);
      SkSeg sg;
      for (int grp = cluster_id;; grp += num_clusters) {
        int g = grp;
        if (grp >= walk.full_end) {
          if (!walk.next(sg)) break;
          if (sg.k0 > 0) continue;  // partial segment: no epilogue, no statistics
          g = sg.grp;
        }
        const int m_blk = (n_fastest ? g / p.num_n_tiles : g % m_groups) * 2 + cta_rank;
        mbar_wait(&stats_bar[st_it & 3], (st_it >> 2) & 1);
        ++st_it;
        __threadfence();
        int last = 0;
        if (lane == 0) {
          const int old = atomicAdd(p.stats_counters + m_blk, 1);
          last = old == p.num_n_tiles - 1;
          if (last) p.stats_counters[m_blk] = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          __threadfence();
          // lane handles rows lane, lane+32, lane+64, lane+96 of the block; all 4 x 8 loads of a round are in
          // flight together (this reduction is the kernel's tail when the block finishes last)
          float sm[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
          for (int i0 = 0; i0 < p.stats_parts_out; i0 += 8) {
            float2 v[4][8];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const int row = m_blk * BM + lane + 32 * rr;
              const float2* st = p.stats_out + (size_t)(row < p.M ? row : 0) * p.stats_parts_out;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[rr][j] = (row < p.M && i0 + j < p.stats_parts_out) ? __ldcg(st + i0 + j) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                sm[rr] += v[rr][j].x;
                ss[rr] += v[rr][j].y;
              }
          }
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int row = m_blk * BM + lane + 32 * rr;
            if (row < p.M) {
              const float mean = p.so_rms ? 0.f : sm[rr] * p.so_inv_dim;
              const float var = fmaxf(fmaf(-mean, mean, ss[rr] * p.so_inv_dim), 0.f);
              p.stats_final[row] = make_float2(mean, rsqrtf(var + p.so_eps));
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int chalf = (warp - 4) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    int st_it = 0;
    uint32_t res_ph = 0;  // parity of this warp's two residual landing barriers
    SkWalk walk(p, num_groups, cluster_id, num_clusters);
    SkSeg sg;
    for (int grp = cluster_id;; grp += num_clusters) {
      int g = grp;
      bool partial = false;
      SkOwner own{cluster_id, cta_rank, 0};
      if (grp >= walk.full_end) {
        if (!walk.next(sg)) break;
        g = sg.grp;
        partial = sg.k0 > 0;
        own.n_peers = sg.n_peers;
      }
      const int m_blk = (n_fastest ? g / p.num_n_tiles : g % m_groups) * 2 + cta_rank;
      const int n_blk = n_fastest ? g % p.num_n_tiles : g / m_groups;
      if (MODE == LLMSEG_GEMM_PLAIN && p.tma_res && !partial)
        res_prefetch_first<BN, G2_EPI_WARPS / 4>(p, smem + C::OFF_RES + (warp - 4) * 4096, res_bar + (warp - 4) * 2, &tmR,
                                                 m_blk, n_blk, quarter, chalf, lane);
      const float rs = partial ? 1.f : row_rstd(p, m_blk * BM + quarter * 32 + lane);
      // row map entry of this lane's row, in flight under the wait (ncu r2n: the QKV epilogue sat ~4 % of the kernel's
      // samples on this load at the top of every tile)
      int out_row_pre = m_blk * BM + quarter * 32 + lane;
      if ((MODE == LLMSEG_GEMM_PLAIN || MODE == LLMSEG_GEMM_QKV) && p.out_row_map != nullptr && out_row_pre < p.M)
        out_row_pre = __ldg(p.out_row_map + out_row_pre);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
      if (partial) {
        // fp32 accumulators -> workspace, then publish (all 256 epilogue threads have fenced their stores)
        sk_dump<BN, G2_EPI_WARPS / 4>(p, taddr, quarter, chalf, lane, cluster_id, cta_rank);
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(G2_EPI_WARPS * 32) : "memory");
        if (warp == 4 && lane == 0) {
          int* flag = p.sk_flags + cluster_id * 2 + cta_rank;
          asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
        }
      } else {
        if (own.n_peers > 0) {
          if (warp == 4 && lane == 0) {
            for (int pr = 1; pr <= own.n_peers; ++pr) {
              const int* flag = p.sk_flags + (cluster_id + pr) * 2 + cta_rank;
              int v = 0;
              long long spins = 0;
              do {
                asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (++spins > (1ll << 22)) __trap();  // a lost producer must not hang the device
              } while (v == 0);
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(G2_EPI_WARPS * 32) : "memory");
        }
        if constexpr (EPI != 0)
          epilogue_plain_fast<BN, EPI>(p, taddr, m_blk, n_blk, quarter, chalf, lane, rs, own,
                                       smem + C::OFF_EPI + (warp - 4) * C::EPI_TILE_BYTES, &tmC,
                                       smem + C::OFF_RES + (warp - 4) * 4096, res_bar + (warp - 4) * 2, &tmR, &res_ph);
        else
          epilogue_tile<BN, MODE, ROPE, G2_EPI_WARPS / 4>(p, nullptr, taddr, m_blk, n_blk, quarter, chalf, lane, rs, own,
                                                          smem + C::OFF_EPI + (warp - 4) * C::EPI_TILE_BYTES, &tmC,
                                                          smem + C::OFF_RES + (warp - 4) * 4096, res_bar + (warp - 4) * 2,
                                                          &tmR, &res_ph, out_row_pre);
        if (own.n_peers > 0) {
          asm volatile("bar.sync 1, %0;" ::"n"(G2_EPI_WARPS * 32) : "memory");  // every reader of the partials is done
          if (warp == 4 && lane == 0)
            for (int pr = 1; pr <= own.n_peers; ++pr) p.sk_flags[(cluster_id + pr) * 2 + cta_rank] = 0;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
      if (MODE == LLMSEG_GEMM_PLAIN && p.stats_final != nullptr && !partial) {
        // partials of this tile are written: tell the statistics warp (release at CTA scope; it adds the
        // GPU-scope fence) and move on — nothing here waits on other CTAs
        if (lane == 0) mbar_arrive(&stats_bar[st_it & 3]);
        ++st_it;
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) bulk_wait0();  // this warp's TMA stores have landed before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int MODE, bool ROPE, int EPI = 0>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
            const GemmDev& d, int grid, cudaStream_t stream) {
  auto kern = gemm2_kernel<BN, MODE, ROPE, EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    LLMSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::SMEM_BYTES));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = Cfg2<BN>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr, 1);
  LLMSEG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, d));
  g_launches.fetch_add(1);
  return 0;
}

// The CTA-pair kernel is the default for problems with >= 8 row tiles (measured +4..15 % over the
// single-CTA kernel there, profiles/r01e_gemm_2cta.md; small-M problems keep the single-CTA kernel:
// a pair wastes half its tile on an odd tile count).  LLMSEG_GEMM_2CTA=0 disables, =1 forces (>= 2 tiles).
bool use_pair_kernel(int m_tiles) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LLMSEG_GEMM_2CTA");
    mode = e ? atoi(e) : 2;
  }
  if (mode == 0) return false;
  if (mode == 1) return m_tiles >= 2;
  return m_tiles >= 8;
}

template <int BN, int MODE, bool ROPE>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& d, int grid,
           cudaStream_t stream) {
  auto kern = gemm_kernel<BN, MODE, ROPE>;
  constexpr int smem_bytes = MODE == MODE_RELPOS ? Cfg<BN>::RP_SMEM_BYTES : Cfg<BN>::SMEM_BYTES;
  static bool attr_done = false;  // idempotent; a race only repeats the call
  if (!attr_done) {
    LLMSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = d.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr, 1);
  LLMSEG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, d));
  g_launches.fetch_add(1);
  return 0;
}

// cluster size along M for a problem with `m_tiles` row tiles: share the W tile between up to 4 CTAs.
// LLMSEG_GEMM_CLUSTER=1|2|4 overrides (tuning / debugging).
int pick_cluster(int m_tiles) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("LLMSEG_GEMM_CLUSTER");
    forced = e ? atoi(e) : 0;
  }
  // measured on B200 (profiles/r01b_gemm_cluster.md): pairs give +3..7 % on M >= 4096, 4-CTA clusters
  // lose it again (stranded SMs), and an odd tile count wastes a whole CTA-tile per group on small M.
  int cl = (m_tiles >= 2 && (m_tiles % 2 == 0 || m_tiles >= 16)) ? 2 : 1;
  if (forced == 1) cl = 1;
  if ((forced == 2 || forced == 4) && m_tiles >= forced) cl = forced;
  return cl;
}

// persistent grid: as many whole clusters as fit the machine (<= one CTA per SM), never more than the work
int pick_grid(int groups, int cl, int sms) {
  int clusters = sms / cl;
  if (cl == 4) clusters = clusters > 33 ? 33 : clusters;  // 4-CTA clusters strand a few SMs (GPC packing)
  if (clusters > groups) clusters = groups;
  if (clusters < 1) clusters = 1;
  return clusters * cl;
}

// LLMSEG_GEMM_TMA_STORE=0: the pair kernel's PLAIN epilogue writes C with per-lane 16-byte stores again
bool tma_store_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LLMSEG_GEMM_TMA_STORE");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0;
}
// LLMSEG_GEMM_TMA_RES=0: residual tiles are read with per-lane global loads again (and the ring keeps all stages)
int streamk_min_kb() {
  const char* e = getenv("LLMSEG_GEMM_SK_MIN_KB");  // read per call (A/B runs)
  return e == nullptr ? 32 : atoi(e);
}

bool tma_res_enabled() {
  const char* e = getenv("LLMSEG_GEMM_TMA_RES");  // read per call: scripts/gpu_gemm_ab.py flips it between launches
  return e == nullptr || atoi(e) != 0;
}
// LLMSEG_GEMM_EPI=0 routes every PLAIN problem through the generic epilogue again (read per call: A/B runs)
bool epi_variants_enabled() {
  const char* e = getenv("LLMSEG_GEMM_EPI");
  return e == nullptr || atoi(e) != 0;
}
// LLMSEG_GEMM_STREAMK=0 disables the stream-K tail (A/B measurements)
bool streamk_enabled() {
  const char* e = getenv("LLMSEG_GEMM_STREAMK");  // read per call (scripts/gpu_gemm_ab.py flips it between launches)
  return e == nullptr || atoi(e) != 0;
}
// LLMSEG_RELPOS_WIN=0 routes the 14x14-window rel-pos prep through the generic GEMM kernel again
bool relpos_win_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LLMSEG_RELPOS_WIN");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0;
}
constexpr size_t SK_FLAG_BYTES = 16384;   // [0, 2048): stream-K flags; [2048, 16384): stats row-block counters
constexpr size_t SK_COUNTER_OFF = 2048;
constexpr int SK_MAX_COUNTERS = (16384 - 2048) / 4;
constexpr size_t SK_MAX_PAIRS = 128;
constexpr size_t SK_WS_BYTES = SK_FLAG_BYTES + SK_MAX_PAIRS * 2 * 128 * 256 * sizeof(float);

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace
}  // namespace llmseg

using namespace llmseg;

extern "C" int llmseg_gemm(const llmseg_gemm_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LLMSEG_REQUIRE(p != nullptr, LLMSEG_EARG, "llmseg_gemm: null params");
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, LLMSEG_ESHAPE, "llmseg_gemm: empty problem %dx%dx%d",
                 p->M, p->N, p->K);
  LLMSEG_REQUIRE(p->K % 8 == 0 && p->lda % 8 == 0 && p->ldw % 8 == 0, LLMSEG_ESHAPE,
                 "llmseg_gemm: K=%d lda=%d ldw=%d must be multiples of 8", p->K, p->lda, p->ldw);
  LLMSEG_REQUIRE(p->N % 8 == 0, LLMSEG_ESHAPE, "llmseg_gemm: N=%d must be a multiple of 8", p->N);
  LLMSEG_REQUIRE(p->A && p->W, LLMSEG_EARG, "llmseg_gemm: null A/W");
  LLMSEG_REQUIRE(p->mode >= 0 && p->mode <= 2, LLMSEG_EARG, "llmseg_gemm: bad mode %d", p->mode);
  LLMSEG_REQUIRE(p->act >= 0 && p->act <= 3, LLMSEG_EARG, "llmseg_gemm: bad act %d", p->act);

  GemmDev d{};
  d.M = p->M; d.N = p->N; d.K = p->K;
  d.C = static_cast<bf16*>(p->C); d.ldc = p->ldc;
  d.bias = static_cast<const bf16*>(p->bias);
  d.residual = static_cast<const bf16*>(p->residual); d.ldr = p->ldr; d.res_mod = p->res_mod;
  d.act = p->act; d.out_row_map = p->out_row_map;
  d.q = static_cast<bf16*>(p->q); d.k = static_cast<bf16*>(p->k); d.vt = static_cast<bf16*>(p->vt);
  d.heads = p->heads; d.head_dim = p->head_dim; d.seq_in = p->seq_in; d.seq_pad = p->seq_pad;
  d.rope_cos = static_cast<const bf16*>(p->rope_cos);
  d.rope_sin = static_cast<const bf16*>(p->rope_sin);
  d.row_stats = static_cast<const float2*>(p->row_stats);
  d.stats_parts_in = p->row_stats_parts;
  d.norm_inv_dim = p->norm_dim > 0 ? 1.0f / (float)p->norm_dim : 0.f;
  d.norm_eps = p->norm_eps;
  d.norm_rms = p->norm_rms;
  d.stats_out = static_cast<float2*>(p->stats_out);
  if (p->row_stats != nullptr)
    LLMSEG_REQUIRE((reinterpret_cast<uintptr_t>(p->row_stats) & 7) == 0 && p->row_stats_parts >= 0 &&
                       (p->row_stats_parts == 0 || p->norm_dim > 0),
                   LLMSEG_EARG, "llmseg_gemm: row_stats must be 8-byte aligned; partials need norm_dim > 0");
  if (p->stats_out != nullptr)
    LLMSEG_REQUIRE(p->mode == LLMSEG_GEMM_PLAIN && (reinterpret_cast<uintptr_t>(p->stats_out) & 7) == 0,
                   LLMSEG_EARG, "llmseg_gemm: stats_out needs PLAIN mode and 8-byte alignment");
  if (p->stats_final != nullptr) {
    LLMSEG_REQUIRE(p->stats_out != nullptr && p->out_row_map == nullptr && p->stats_dim > 0 &&
                       (reinterpret_cast<uintptr_t>(p->stats_final) & 7) == 0,
                   LLMSEG_EARG, "llmseg_gemm: stats_final needs stats_out, no out_row_map and stats_dim > 0");
    LLMSEG_REQUIRE(p->workspace != nullptr && p->workspace_bytes >= SK_WS_BYTES &&
                       (p->M + BM - 1) / BM <= SK_MAX_COUNTERS,
                   LLMSEG_EARG, "llmseg_gemm: stats_final needs the workspace and M <= %d", SK_MAX_COUNTERS * BM);
    d.stats_final = static_cast<float2*>(p->stats_final);
    d.stats_counters = reinterpret_cast<int*>(static_cast<uint8_t*>(p->workspace) + SK_COUNTER_OFF);
    d.so_inv_dim = 1.0f / (float)p->stats_dim;
    d.so_eps = p->stats_eps;
    d.so_rms = p->stats_rms;
  }

  bool rope = false;
  if (p->mode == LLMSEG_GEMM_QKV) {
    LLMSEG_REQUIRE(p->q && p->k && p->vt, LLMSEG_EARG, "llmseg_gemm(QKV): null q/k/vt");
    LLMSEG_REQUIRE(p->heads > 0 && p->head_dim % 8 == 0 && p->N == 3 * p->heads * p->head_dim,
                   LLMSEG_ESHAPE, "llmseg_gemm(QKV): N=%d != 3*%d*%d", p->N, p->heads, p->head_dim);
    LLMSEG_REQUIRE(p->seq_in > 0 && p->seq_pad >= p->seq_in && p->seq_pad % 8 == 0 &&
                       (p->out_row_map != nullptr || p->M % p->seq_in == 0),
                   LLMSEG_ESHAPE, "llmseg_gemm(QKV): M=%d seq_in=%d seq_pad=%d inconsistent", p->M,
                   p->seq_in, p->seq_pad);
    rope = p->rope_cos != nullptr;
    if (rope)
      LLMSEG_REQUIRE(p->head_dim == 128 && p->rope_sin != nullptr, LLMSEG_ESHAPE,
                     "llmseg_gemm(QKV+RoPE): head_dim must be 128 (got %d)", p->head_dim);
  } else {
    LLMSEG_REQUIRE(p->C != nullptr, LLMSEG_EARG, "llmseg_gemm: null C");
    LLMSEG_REQUIRE(p->ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(p->C) & 15) == 0, LLMSEG_EALIGN,
                   "llmseg_gemm: C / ldc not 16-byte aligned");
    if (p->residual)
      LLMSEG_REQUIRE(p->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p->residual) & 15) == 0,
                     LLMSEG_EALIGN, "llmseg_gemm: residual / ldr not 16-byte aligned");
    if (p->mode == LLMSEG_GEMM_SWIGLU)
      LLMSEG_REQUIRE(p->N % 16 == 0, LLMSEG_ESHAPE, "llmseg_gemm(SWIGLU): N=%d %% 16 != 0", p->N);
  }
  if (p->bias)
    LLMSEG_REQUIRE((reinterpret_cast<uintptr_t>(p->bias) & 15) == 0, LLMSEG_EALIGN,
                   "llmseg_gemm: bias not 16-byte aligned");

  // tile shape: BN=256 when it still fills the machine, else BN=128 for more CTAs
  const int m_tiles = (p->M + BM - 1) / BM;
  const int sms = num_sms();
  int bn = 256;
  if (p->N < 256 || m_tiles * ((p->N + 255) / 256) < sms) bn = 128;
  // (N = 384 — the mask decoder's image-side projection, a full and a half-empty 256-wide tile per row block — measured
  //  407 us with 256-wide tiles, 736 us with 128-wide ones, 470 us as a 256 + 128 pair of GEMMs: the wide tile stays)
  // (tried: 128-wide pair tiles when 256-wide ones quantise badly, e.g. LLaMA o_proj/down at batch 8 with
  //  2.16 waves — slower: a 256x128 pair tile needs twice the L2->SM bytes per flop and becomes L2-bound.)
  d.num_m_tiles = m_tiles;
  d.num_n_tiles = (p->N + bn - 1) / bn;
  d.stats_parts_out = d.num_n_tiles * 2;
  d.num_k_blocks = (p->K + BK - 1) / BK;
  d.cluster = pick_cluster(m_tiles);
  d.n_fastest = p->M > p->N ? 1 : 0;
  const int grid = pick_grid(((m_tiles + d.cluster - 1) / d.cluster) * d.num_n_tiles, d.cluster, sms);

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->M};
    uint64_t str[1] = {(uint64_t)p->lda * 2};
    uint32_t box[2] = {BK, BM};
    if (int e = make_tmap_bf16(&tmA, p->A, 2, dims, str, box, 128)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
    uint64_t str[1] = {(uint64_t)p->ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)(bn / d.cluster)};  // each CTA loads (and multicasts) 1/cluster of the W tile
    if (int e = make_tmap_bf16(&tmB, p->W, 2, dims, str, box, 128)) return e;
  }

  if (use_pair_kernel(m_tiles)) {
    d.cluster = 2;
    const int groups = ((m_tiles + 1) / 2) * d.num_n_tiles;
    int pgrid = pick_grid(groups, 2, sms);
    // stream-K tail: when the last wave holds `rem` < n_pairs tiles, cut it along K into one range of
    // sk_w k-blocks per pair — worth it when that shortens the wave by >= 12 k-blocks (the fp32 partial
    // round trip through L2 costs about 6-8) and the ranges stay >= 4 k-blocks long.
    const int n_pairs = sms / 2;
    if (streamk_enabled() && p->workspace != nullptr && p->workspace_bytes >= SK_WS_BYTES &&
        (size_t)n_pairs <= SK_MAX_PAIRS && (reinterpret_cast<uintptr_t>(p->workspace) & 255) == 0) {
      const int rem = groups % n_pairs;
      const int kb = d.num_k_blocks;
      // at most ~4 segments per tile: every extra partial is another 128 KB round trip for the owner
      int w = (rem * kb + n_pairs - 1) / n_pairs;
      if (w < (kb + 3) / 4) w = (kb + 3) / 4;
      // (round 2 tried the tail for K = 1280 too — LLMSEG_GEMM_SK_MIN_KB=16: at batch 1 the SAM GEMMs run 1.08 / 3.2 / 4.3
      //  waves — and measured no gain, 14.97 vs 15.23 ms per step: the text branch's kernels on the second stream already
      //  fill the idle SMs of those tail waves.  The floor stays at 32 k-blocks.)
      if (rem > 0 && kb >= streamk_min_kb() && kb - w >= 12 && w >= 4) {
        d.sk_w = w;
        d.sk_flags = static_cast<int*>(p->workspace);
        d.sk_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(p->workspace) + SK_FLAG_BYTES);
        pgrid = n_pairs * 2;
      }
    }
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
    uint64_t str[1] = {(uint64_t)p->ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)(bn / 2)};
    if (int e = make_tmap_bf16(&tmB, p->W, 2, dims, str, box, 128)) return e;
    // C through TMA tile stores: plain row-major output, rows not scattered (LLMSEG_GEMM_TMA_STORE=0: off)
    CUtensorMap tmC = tmA;
    d.tma_store = 0;
    if (p->mode == LLMSEG_GEMM_PLAIN && p->out_row_map == nullptr && tma_store_enabled()) {
      uint64_t cdims[2] = {(uint64_t)p->N, (uint64_t)p->M};
      uint64_t cstr[1] = {(uint64_t)p->ldc * 2};
      uint32_t cbox[2] = {64, 32};
      if (int e = make_tmap_bf16(&tmC, p->C, 2, cdims, cstr, cbox, 128)) return e;
      d.tma_store = 1;
    }
    // residual through TMA landing buffers (LLMSEG_GEMM_TMA_RES=0: per-lane loads again): same tiling as C
    CUtensorMap tmR = tmA;
    d.tma_res = 0;
    if (d.tma_store && bn == 256 && p->residual != nullptr && p->res_mod % 32 == 0 && tma_res_enabled()) {
      uint64_t rdims[2] = {(uint64_t)p->N, (uint64_t)(p->res_mod > 0 ? p->res_mod : p->M)};
      uint64_t rstr[1] = {(uint64_t)p->ldr * 2};
      uint32_t rbox[2] = {32, 32};
      if (int e = make_tmap_bf16(&tmR, p->residual, 2, rdims, rstr, rbox, 64)) return e;
      d.tma_res = 1;
    }
#define LLMSEG_GEMM2_DISPATCH(BN_)                                                              \
  switch (p->mode) {                                                                            \
    case LLMSEG_GEMM_PLAIN: return launch2<BN_, LLMSEG_GEMM_PLAIN, false>(tmA, tmB, tmC, tmR, d, pgrid, stream);   \
    case LLMSEG_GEMM_SWIGLU: return launch2<BN_, LLMSEG_GEMM_SWIGLU, false>(tmA, tmB, tmC, tmR, d, pgrid, stream); \
    default:                                                                                    \
      return rope ? launch2<BN_, LLMSEG_GEMM_QKV, true>(tmA, tmB, tmC, tmR, d, pgrid, stream)              \
                  : launch2<BN_, LLMSEG_GEMM_QKV, false>(tmA, tmB, tmC, tmR, d, pgrid, stream);            \
  }
    // straight-line epilogue variants (see EpiX): BN = 256, PLAIN, staged TMA store, whole 64-column groups
    if (bn == 256 && p->mode == LLMSEG_GEMM_PLAIN && d.tma_store && p->N % 64 == 0 && epi_variants_enabled()) {
      const bool res = p->residual != nullptr;
      if (p->act == LLMSEG_ACT_GELU && !res && p->stats_out == nullptr && p->bias != nullptr)
        return launch2<256, LLMSEG_GEMM_PLAIN, false, 1>(tmA, tmB, tmC, tmR, d, pgrid, stream);
      if (p->act == LLMSEG_ACT_NONE && res && d.tma_res && p->stats_out != nullptr && p->bias != nullptr &&
          p->row_stats == nullptr)
        return launch2<256, LLMSEG_GEMM_PLAIN, false, 2>(tmA, tmB, tmC, tmR, d, pgrid, stream);
      if (p->act == LLMSEG_ACT_NONE && res && d.tma_res && p->stats_out == nullptr && p->bias == nullptr &&
          p->row_stats == nullptr)
        return launch2<256, LLMSEG_GEMM_PLAIN, false, 3>(tmA, tmB, tmC, tmR, d, pgrid, stream);
      if (p->act == LLMSEG_ACT_NONE && res && d.tma_res && p->stats_out == nullptr && p->bias != nullptr &&
          p->row_stats == nullptr)
        return launch2<256, LLMSEG_GEMM_PLAIN, false, 4>(tmA, tmB, tmC, tmR, d, pgrid, stream);
      if (p->act == LLMSEG_ACT_NONE && !res && p->stats_out == nullptr && p->bias != nullptr && p->row_stats == nullptr)
        return launch2<256, LLMSEG_GEMM_PLAIN, false, 5>(tmA, tmB, tmC, tmR, d, pgrid, stream);
    }
    if (bn == 256) {
      LLMSEG_GEMM2_DISPATCH(256)
    } else {
      LLMSEG_GEMM2_DISPATCH(128)
    }
#undef LLMSEG_GEMM2_DISPATCH
  }

#define LLMSEG_GEMM_DISPATCH(BN_)                                                            \
  switch (p->mode) {                                                                         \
    case LLMSEG_GEMM_PLAIN: return launch<BN_, LLMSEG_GEMM_PLAIN, false>(tmA, tmB, d, grid, stream); \
    case LLMSEG_GEMM_SWIGLU: return launch<BN_, LLMSEG_GEMM_SWIGLU, false>(tmA, tmB, d, grid, stream); \
    default:                                                                                 \
      return rope ? launch<BN_, LLMSEG_GEMM_QKV, true>(tmA, tmB, d, grid, stream)             \
                  : launch<BN_, LLMSEG_GEMM_QKV, false>(tmA, tmB, d, grid, stream);           \
  }
  if (bn == 256) {
    LLMSEG_GEMM_DISPATCH(256)
  } else {
    LLMSEG_GEMM_DISPATCH(128)
  }
#undef LLMSEG_GEMM_DISPATCH
}

extern "C" size_t llmseg_gemm_workspace_bytes(void) { return SK_WS_BYTES; }

// partials per output row that llmseg_gemm writes to stats_out for an (M, N) PLAIN problem — must mirror the
// tile-shape choice in llmseg_gemm (two epilogue warps share a row of every N-tile)
extern "C" int llmseg_gemm_stats_parts(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  const int m_tiles = (M + BM - 1) / BM;
  int bn = 256;
  if (N < 256 || m_tiles * ((N + 255) / 256) < num_sms()) bn = 128;
  return ((N + bn - 1) / bn) * 2;
}

extern "C" int llmseg_relpos_prep(const void* q, const void* rel_hw, int n_pad, int bh, int seq,
                                  int seq_pad, int head_dim, int grid, float inv_scale, void* qext,
                                  int ext_cols, void* row_bias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(q && rel_hw && qext, LLMSEG_EARG, "llmseg_relpos_prep: null pointer");
  LLMSEG_REQUIRE(grid > 0 && seq == grid * grid && seq_pad >= seq && head_dim % 8 == 0 &&
                     head_dim <= 128 && n_pad == (row_bias ? 256 : 64),
                 LLMSEG_ESHAPE, "llmseg_relpos_prep: grid=%d seq=%d seq_pad=%d head_dim=%d n_pad=%d", grid,
                 seq, seq_pad, head_dim, n_pad);
  LLMSEG_REQUIRE((row_bias == nullptr && ext_cols == 32 && 2 * grid <= 32) ||
                     (row_bias != nullptr && ext_cols == 64 && grid <= 64),
                 LLMSEG_ESHAPE, "llmseg_relpos_prep: ext_cols=%d inconsistent with grid=%d", ext_cols, grid);
  if (row_bias == nullptr && grid == 14 && head_dim == 80 && relpos_win_enabled())
    return launch_relpos_win(q, rel_hw, bh * seq_pad, seq, seq_pad, inv_scale, qext, stream);
  GemmDev d{};
  d.M = bh * seq_pad; d.N = n_pad; d.K = head_dim;
  d.rp_grid = grid; d.rp_seq = seq; d.rp_seq_pad = seq_pad; d.rp_ext = ext_cols;
  d.rp_inv_scale = inv_scale;
  d.rp_qext = static_cast<bf16*>(qext);
  d.rp_rh = static_cast<bf16*>(row_bias);
  const int sms = num_sms();
  const int bn = 128;
  d.num_m_tiles = (d.M + BM - 1) / BM;
  d.num_n_tiles = (d.N + bn - 1) / bn;
  d.num_k_blocks = (d.K + BK - 1) / BK;
  d.cluster = pick_cluster(d.num_m_tiles);
  d.n_fastest = 1;
  const int grid_x = pick_grid(((d.num_m_tiles + d.cluster - 1) / d.cluster) * d.num_n_tiles, d.cluster, sms);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)head_dim, (uint64_t)d.M};
    uint64_t str[1] = {(uint64_t)head_dim * 2};
    uint32_t box[2] = {BK, BM};
    if (int e = make_tmap_bf16(&tmA, q, 2, dims, str, box, 128)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)head_dim, (uint64_t)n_pad};
    uint64_t str[1] = {(uint64_t)head_dim * 2};
    uint32_t box[2] = {BK, (uint32_t)(bn / d.cluster)};
    if (int e = make_tmap_bf16(&tmB, rel_hw, 2, dims, str, box, 128)) return e;
  }
  return launch<128, MODE_RELPOS, false>(tmA, tmB, d, grid_x, stream);
}
