// llmseg_b200 — SAM-Everything proposal generation: the kernels around the tcgen05 GEMMs that turn
// SAM ViT-H image features into the K soft mask proposals LLM-Seg's selector consumes.
//
//   point_tokens     prompt encoder for one foreground point per prompt + output tokens
//                    (reference segment_anything/modeling/prompt_encoder.py:73-90,186-195; mask_decoder.py:123-131)
//   tok2img_attn_mma 7 tokens x 4096 image keys, 8 heads x 16 (reference modeling/transformer.py:222-242 as called at
//                    :169-172 and :101-103): block per prompt, warp per head, K / V through a warp-private cp.async ring,
//                    both products on mma.sync, online softmax (tok2img_attn: the CUDA-core kernel it replaced, kept
//                    behind LLMSEG_T2I_V1 and compared with it in the tests)
//   img2tok_attn     4096 image queries x 7 token keys (transformer.py:179-182): warp per head, token per lane, K / V as
//                    warp-uniform shared-memory broadcasts, query / output tiles staged with whole-row copies
//   upscale_logits   LayerNorm2d(64) + GELU -> second ConvTranspose (mma.sync) + GELU -> hyper-network product in one
//                    kernel (mask_decoder.py:56-64,139-157); the three un-fused steps stay as the comparison path:
//   ln64_gelu        LayerNorm2d(64) + GELU of the first up-scaling step on the un-shuffled ConvTranspose output
//                    (mask_decoder.py:56-62): every group of 64 columns is one output pixel
//   mask_logits      hyper-network product on the un-shuffled second ConvTranspose output -> low-res mask logits
//                    [P,3,256,256] (mask_decoder.py:143-157, multimask slice :101-104)
//   mask_stats       per candidate, on the 4x bilinear up-sampling of its logits evaluated on the fly (sam.py:155-166):
//                    area, the two stability counts (utils/amg.py:156-176) and the box (utils/amg.py:303-346)
//   box_nms          greedy box NMS over score-sorted boxes (automatic_mask_generator.py:256-262 -> torchvision nms)
//   mask_soft        antialiased bilinear 1024 -> 256 resize of the binarised up-sampled mask (utils/dataset.py:620-622),
//                    again straight from the low-res logits: the 1024 x 1024 masks never exist unless asked for
//   mask_binarize    the 1024 x 1024 binary masks themselves (what `origin_segs_list` holds), on request
//
// Most of these are HBM / latency bound integer-and-compare work; the dense algebra of the decoder (projections, MLPs,
// the first ConvTranspose as a GEMM) runs on llmseg_gemm.
#include <atomic>

#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
namespace {

constexpr int E = 256;      // transformer width
constexpr int NTOK = 7;     // iou token + 4 mask tokens + point + padding point
constexpr int IMG_TOK = 4096;
constexpr int LOW = 256;    // low-res mask side
constexpr int HI = 1024;    // image side

// ---------------------------------------------------------------------------------------------
// tokens[p, 0..4] = iou / mask tokens; [5] = PE(point) + point_embeddings[1]; [6] = not_a_point_embed
// PE(c) = sin / cos(2*pi * ((2u-1) G[0,c] + (2v-1) G[1,c])), (u, v) = (x + 0.5, y + 0.5) / img
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
point_tokens_kernel(const float* __restrict__ pts, const float* __restrict__ gauss, const bf16* __restrict__ out_tok,
                    const bf16* __restrict__ point_emb, const bf16* __restrict__ not_a_point, bf16* __restrict__ tokens,
                    float inv_img) {
  const int p = blockIdx.x, c = threadIdx.x;
  bf16* t = tokens + (size_t)p * NTOK * E;
#pragma unroll
  for (int i = 0; i < 5; ++i) t[i * E + c] = out_tok[i * E + c];
  const float u = 2.f * ((pts[2 * p] + 0.5f) * inv_img) - 1.f;
  const float v = 2.f * ((pts[2 * p + 1] + 0.5f) * inv_img) - 1.f;
  const int cc = c & 127;
  const float arg = 6.283185307179586f * (u * gauss[cc] + v * gauss[128 + cc]);
  const float pe = c < 128 ? sinf(arg) : cosf(arg);
  t[5 * E + c] = __float2bfloat16_rn(pe + __bfloat162float(point_emb[c]));
  t[6 * E + c] = not_a_point[c];
}

// ---------------------------------------------------------------------------------------------
// token -> image attention.  grid (P, 8); 256 threads; thread t owns keys t, t+256, ...
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load16(const bf16* p, float* f) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 v = unpack_bf16(w[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}

__global__ void __launch_bounds__(256, 1)
tok2img_attn_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, int ldk, long long k_bs,
                    const bf16* __restrict__ v, int ldv, long long v_bs, bf16* __restrict__ out, int ldo) {
  constexpr int KPT = IMG_TOK / 256;  // keys per thread
  const int p = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ float qs[NTOK][16];
  __shared__ float red[8][NTOK];
  __shared__ float accs[8][NTOK * 16];
  if (tid < NTOK * 16) {
    const int qi = tid >> 4, d = tid & 15;
    qs[qi][d] = __bfloat162float(q[(size_t)(p * NTOK + qi) * ldq + h * 16 + d]) * 0.25f;  // 1/sqrt(16)
  }
  __syncthreads();
  const bf16* kp = k + (size_t)p * k_bs + h * 16;
  const bf16* vp = v + (size_t)p * v_bs + h * 16;
  // two passes over the keys (scores are recomputed in the second one: K comes back from L1/L2, and keeping the
  // 16 x 7 scores next to the 7 x 16 accumulators would not fit the register file)
  float mx[NTOK];
#pragma unroll
  for (int i = 0; i < NTOK; ++i) mx[i] = -INFINITY;
#pragma unroll 4
  for (int j = 0; j < KPT; ++j) {
    float kf[16];
    load16(kp + (size_t)(tid + j * 256) * ldk, kf);
#pragma unroll
    for (int i = 0; i < NTOK; ++i) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) a = fmaf(qs[i][d], kf[d], a);
      mx[i] = fmaxf(mx[i], a);
    }
  }
#pragma unroll
  for (int i = 0; i < NTOK; ++i) mx[i] = warp_max(mx[i]);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NTOK; ++i) red[warp][i] = mx[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NTOK; ++i) {
    float m = red[0][i];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w][i]);
    mx[i] = m;
  }
  __syncthreads();
  float sum[NTOK];
  float acc[NTOK][16];
#pragma unroll
  for (int i = 0; i < NTOK; ++i) {
    sum[i] = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[i][d] = 0.f;
  }
#pragma unroll 2
  for (int j = 0; j < KPT; ++j) {
    float kf[16], vf[16];
    load16(kp + (size_t)(tid + j * 256) * ldk, kf);
    load16(vp + (size_t)(tid + j * 256) * ldv, vf);
#pragma unroll
    for (int i = 0; i < NTOK; ++i) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) a = fmaf(qs[i][d], kf[d], a);
      const float pr = __expf(a - mx[i]);
      sum[i] += pr;
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[i][d] = fmaf(pr, vf[d], acc[i][d]);
    }
  }
  // block reduction: shuffles inside a warp, fixed-order sum over the 8 warps (deterministic)
#pragma unroll
  for (int i = 0; i < NTOK; ++i) {
    sum[i] = warp_sum(sum[i]);
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[i][d] = warp_sum(acc[i][d]);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NTOK; ++i) {
      red[warp][i] = sum[i];
#pragma unroll
      for (int d = 0; d < 16; ++d) accs[warp][i * 16 + d] = acc[i][d];
    }
  }
  __syncthreads();
  if (tid < NTOK * 16) {
    const int qi = tid >> 4;
    float a = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += accs[w][tid];
      l += red[w][qi];
    }
    out[(size_t)(p * NTOK + qi) * ldo + h * 16 + (tid & 15)] = __float2bfloat16_rn(a / l);
  }
}

// ---------------------------------------------------------------------------------------------
// token -> image attention, tensor-core version.  grid (P); 256 threads = 8 warps, ONE HEAD PER WARP: the warp walks the
// 4096 keys of its head in 32-key chunks that arrive through a warp-private cp.async ring (no block barrier in the
// loop), S = Q K^T and O += P V are mma.sync m16n8k16 (7 live query rows of 16; head_dim 16 = one k-step), softmax is
// the usual online form on the 8 live scores a lane holds per chunk.  ~60 registers instead of 234, so 24 warps per SM
// keep ~100 KB of K / V in flight per SM — the thread-per-key kernel above ran 8 warps per SM at 1.2 TB/s.
// ---------------------------------------------------------------------------------------------
constexpr int T2I_CHUNK = 32;                       // keys per chunk
constexpr int T2I_STAGES = 4;
constexpr int T2I_STAGE_BYTES = T2I_CHUNK * 32 * 2;  // K rows then V rows, 32 B (16 bf16) each
constexpr int T2I_SMEM = 8 * T2I_STAGES * T2I_STAGE_BYTES;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col); rows 8..15 of A are zero here (a1 = a3 = 0)
__device__ __forceinline__ void mma16816(float* d, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 3)
tok2img_attn_mma_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, int ldk, long long k_bs,
                        const bf16* __restrict__ v, int ldv, long long v_bs, bf16* __restrict__ out, int ldo) {
  extern __shared__ __align__(128) uint8_t t2i_smem[];
  const int p = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qr = lane >> 2, qc = (lane & 3) * 2;     // fragment row (query / key-in-block / dim) and column pair
  const uint32_t ring = smem_u32(t2i_smem) + h * (T2I_STAGES * T2I_STAGE_BYTES);
  // A fragment of Q (rows 7..15 are zero padding): a0 = Q[qr][qc, qc+1], a2 = Q[qr][qc+8, qc+9]
  uint32_t qa0 = 0, qa2 = 0;
  if (qr < NTOK) {
    const bf16* qp = q + (size_t)(p * NTOK + qr) * ldq + h * 16 + qc;
    qa0 = *reinterpret_cast<const uint32_t*>(qp);
    qa2 = *reinterpret_cast<const uint32_t*>(qp + 8);
  }
  // a lane pair fetches one key's 32-byte head slice: 16 keys per instruction, whole sectors
  const int lk = lane >> 1, lh = lane & 1;
  const bf16* kp = k + (size_t)p * k_bs + (size_t)lk * ldk + h * 16 + lh * 8;
  const bf16* vp = v + (size_t)p * v_bs + (size_t)lk * ldv + h * 16 + lh * 8;
  const uint32_t st_off = lk * 32 + lh * 16;
  auto fetch = [&](int chunk) {
    const uint32_t dst = ring + (chunk % T2I_STAGES) * T2I_STAGE_BYTES + st_off;
    const size_t key0 = (size_t)chunk * T2I_CHUNK;
    cp_async16(dst, kp + key0 * ldk);
    cp_async16(dst + 16 * 32, kp + (key0 + 16) * ldk);
    cp_async16(dst + T2I_CHUNK * 32, vp + key0 * ldv);
    cp_async16(dst + T2I_CHUNK * 32 + 16 * 32, vp + (key0 + 16) * ldv);
  };
  constexpr int NCH = IMG_TOK / T2I_CHUNK;
#pragma unroll
  for (int c = 0; c < T2I_STAGES - 1; ++c) {
    fetch(c);
    cp_async_commit();
  }
  // ldmatrix lane addresses inside a stage.  K (non-transposed; matrices: keys 0-7 x dims 0-7 | keys 0-7 x dims 8-15 |
  // keys 8-15 x dims 0-7 | keys 8-15 x dims 8-15 -> B fragments of two 8-key blocks), V (transposed; matrices: keys 0-7 x
  // dims 0-7 | keys 8-15 x dims 0-7 | keys 0-7 x dims 8-15 | keys 8-15 x dims 8-15 -> B fragments of the two dim blocks)
  const int mi = lane >> 3, mr = lane & 7;
  const uint32_t k_lm = ((mi >> 1) * 8 + mr) * 32 + (mi & 1) * 16;
  const uint32_t v_lm = T2I_CHUNK * 32 + ((mi & 1) * 8 + mr) * 32 + (mi >> 1) * 16;
  constexpr float C1 = 0.25f * 1.4426950408889634f;   // 1/sqrt(16) * log2(e)
  float m = -INFINITY, l = 0.f;
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};   // O[qr][qc..] for dims 0-7 / 8-15 (d[2], d[3]: padding rows)
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    cp_async_wait<T2I_STAGES - 2>();
    __syncwarp();
    if (c + T2I_STAGES - 1 < NCH) fetch(c + T2I_STAGES - 1);
    cp_async_commit();
    const uint32_t st = ring + (c % T2I_STAGES) * T2I_STAGE_BYTES;
    float s[4][4];
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {   // 16 keys per ldmatrix
      uint32_t kf[4];
      ldmatrix_x4(st + kb * 16 * 32 + k_lm, kf);
#pragma unroll
      for (int e = 0; e < 4; ++e) s[kb * 2][e] = s[kb * 2 + 1][e] = 0.f;
      mma16816(s[kb * 2], qa0, qa2, kf[0], kf[1]);
      mma16816(s[kb * 2 + 1], qa0, qa2, kf[2], kf[3]);
    }
    // live scores of this lane: s[nb][0], s[nb][1] = query qr x keys nb*8 + qc, qc+1
    float mx = fmaxf(fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1])), fmaxf(fmaxf(s[2][0], s[2][1]), fmaxf(s[3][0], s[3][1])));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m, mx * C1);
    const float alpha = ex2(m - m_new);   // m = -inf on the first chunk: exp2(-inf) = 0
    m = m_new;
    uint32_t pa[4];
    float ls = 0.f;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const float p0 = ex2(fmaf(s[nb][0], C1, -m_new)), p1 = ex2(fmaf(s[nb][1], C1, -m_new));
      ls += p0 + p1;
      pa[nb] = pack_bf16(p0, p1);
    }
    l = fmaf(l, alpha, ls);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      o0[e] *= alpha;
      o1[e] *= alpha;
    }
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      uint32_t vf[4];
      ldmatrix_x4_trans(st + kb * 16 * 32 + v_lm, vf);
      mma16816(o0, pa[kb * 2], pa[kb * 2 + 1], vf[0], vf[1]);
      mma16816(o1, pa[kb * 2], pa[kb * 2 + 1], vf[2], vf[3]);
    }
  }
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  if (qr < NTOK) {
    const float inv = 1.f / l;
    bf16* op = out + (size_t)(p * NTOK + qr) * ldo + h * 16 + qc;
    *reinterpret_cast<uint32_t*>(op) = pack_bf16(o0[0] * inv, o0[1] * inv);
    *reinterpret_cast<uint32_t*>(op + 8) = pack_bf16(o1[0] * inv, o1[1] * inv);
  }
}

// ---------------------------------------------------------------------------------------------
// image -> token attention.  grid (4096 / I2T_TOK, P); 256 threads = 8 warps, ONE HEAD PER WARP and one image token
// per lane, so every K / V read from shared memory is a warp-uniform 128-bit broadcast (a thread-per-(token, head)
// layout put 4 heads of a warp on the same bank: 4-way conflicts on 224 loads per thread made the kernel
// shared-memory bound at 0.6 TB/s).  Each block walks I2T_TOK tokens so the 7 KB K/V staging is paid once per 256 rows.
// ---------------------------------------------------------------------------------------------
constexpr int I2T_TOK = 256;
constexpr int I2T_ROW = 272;   // staged row pitch (256 B + 16): a lane's 16-byte reads at pitch 272 hit distinct banks
__global__ void __launch_bounds__(256)
img2tok_attn_kernel(const bf16* __restrict__ q, int ldq, long long q_bs, const bf16* __restrict__ k, int ldk,
                    const bf16* __restrict__ v, int ldv, bf16* __restrict__ out, int ldo) {
  const int p = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ __align__(16) float ks[NTOK][128];
  __shared__ __align__(16) float vs[NTOK][128];
  // Query rows and output rows pass through shared memory (ncu r2r: with every lane reading its own 32-byte head slice
  // straight from global memory each load / store instruction touched 32 cache lines and the kernel sat at 91 % L1
  // throughput, 2.0 TB/s): 32 tokens x 256 B are fetched with whole-row 16-byte copies (cp.async, one tile ahead).
  __shared__ __align__(16) uint8_t qs[2][32 * I2T_ROW];
  __shared__ __align__(16) uint8_t os[32 * I2T_ROW];
  for (int i = threadIdx.x; i < NTOK * 128; i += 256) {
    const int j = i >> 7, c = i & 127;
    ks[j][c] = __bfloat162float(k[(size_t)(p * NTOK + j) * ldk + c]) * 0.25f;   // 1/sqrt(16) folded into K
    vs[j][c] = __bfloat162float(v[(size_t)(p * NTOK + j) * ldv + c]);
  }
  const bf16* qp = q + (size_t)p * q_bs + (size_t)blockIdx.x * I2T_TOK * ldq;
  bf16* op = out + ((size_t)p * IMG_TOK + (size_t)blockIdx.x * I2T_TOK) * ldo;
  // chunk c of a 32-token tile: token c / 16, 16-byte piece c % 16; two chunks per thread
  auto fetch = [&](int it, int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = threadIdx.x + i * 256, tok = c >> 4, piece = c & 15;
      cp_async16(smem_u32(qs[buf]) + tok * I2T_ROW + piece * 16, qp + (size_t)(it * 32 + tok) * ldq + piece * 8);
    }
    cp_async_commit();
  };
  fetch(0, 0);
  for (int it = 0; it < I2T_TOK / 32; ++it) {
    if (it + 1 < I2T_TOK / 32) fetch(it + 1, (it + 1) & 1);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();   // tile `it` (and, first time, K / V) visible; the previous tile's output rows have been stored
    float qf[16];
    {
      const uint4* src = reinterpret_cast<const uint4*>(qs[it & 1] + lane * I2T_ROW + h * 32);
      const uint4 a = src[0], b2 = src[1];
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 t = unpack_bf16(w[i]);
        qf[2 * i] = t.x;
        qf[2 * i + 1] = t.y;
      }
    }
    float s[NTOK], m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        const float4 kk = *reinterpret_cast<const float4*>(&ks[j][h * 16 + d4 * 4]);
        a = fmaf(qf[d4 * 4 + 0], kk.x, a);
        a = fmaf(qf[d4 * 4 + 1], kk.y, a);
        a = fmaf(qf[d4 * 4 + 2], kk.z, a);
        a = fmaf(qf[d4 * 4 + 3], kk.w, a);
      }
      s[j] = a;
      m = fmaxf(m, a);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      s[j] = __expf(s[j] - m);
      l += s[j];
    }
    const float inv = 1.f / l;
    float o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        const float4 vv = *reinterpret_cast<const float4*>(&vs[j][h * 16 + d4 * 4]);
        o[d4 * 4 + 0] = fmaf(s[j], vv.x, o[d4 * 4 + 0]);
        o[d4 * 4 + 1] = fmaf(s[j], vv.y, o[d4 * 4 + 1]);
        o[d4 * 4 + 2] = fmaf(s[j], vv.z, o[d4 * 4 + 2]);
        o[d4 * 4 + 3] = fmaf(s[j], vv.w, o[d4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] *= inv;
    uint4 w0, w1;
    w0.x = pack_bf16(o[0], o[1]); w0.y = pack_bf16(o[2], o[3]); w0.z = pack_bf16(o[4], o[5]); w0.w = pack_bf16(o[6], o[7]);
    w1.x = pack_bf16(o[8], o[9]); w1.y = pack_bf16(o[10], o[11]); w1.z = pack_bf16(o[12], o[13]); w1.w = pack_bf16(o[14], o[15]);
    uint4* dst = reinterpret_cast<uint4*>(os + lane * I2T_ROW + h * 32);
    dst[0] = w0;
    dst[1] = w1;
    __syncthreads();   // output tile complete; every warp is done with query tile `it` (its buffer is refilled next)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = threadIdx.x + i * 256, tok = c >> 4, piece = c & 15;
      *reinterpret_cast<uint4*>(op + (size_t)(it * 32 + tok) * ldo + piece * 8) =
          *reinterpret_cast<const uint4*>(os + tok * I2T_ROW + piece * 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rows of 64 bf16: y = GELU(LayerNorm(x) * gamma + beta), in place allowed.  8 lanes per row, 8 values per lane.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln64_gelu_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const bf16* __restrict__ gamma,
                 const bf16* __restrict__ beta, long long rows, float eps) {
  const long long row = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int part = threadIdx.x & 7;
  if (row >= rows) return;   // whole 8-lane groups leave together: the xor-shuffles below stay inside a group
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in + row * 64 + part * 8));
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  float x[8], s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = unpack_bf16(w[i]);
    x[2 * i] = v.x;
    x[2 * i + 1] = v.y;
    s += v.x + v.y;
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 64.f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] -= mean;
    ss = fmaf(x[i], x[i], ss);
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss * (1.f / 64.f) + eps);
  const uint4 gr = __ldg(reinterpret_cast<const uint4*>(gamma + part * 8));
  const uint4 br = __ldg(reinterpret_cast<const uint4*>(beta + part * 8));
  const uint32_t gw[4] = {gr.x, gr.y, gr.z, gr.w}, bw[4] = {br.x, br.y, br.z, br.w};
  uint32_t o4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 g = unpack_bf16(gw[i]), b = unpack_bf16(bw[i]);
    const float y0 = fmaf(x[2 * i] * rstd, g.x, b.x), y1 = fmaf(x[2 * i + 1] * rstd, g.y, b.y);
    o4[i] = pack_bf16(0.5f * y0 * (1.f + erff(y0 * 0.70710678118654752f)),
                      0.5f * y1 * (1.f + erff(y1 * 0.70710678118654752f)));
  }
  *reinterpret_cast<uint4*>(out + row * 64 + part * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
}

// ---------------------------------------------------------------------------------------------
// low-res mask logits.  up2 rows are (prompt, ty, tx, dy, dx), its 128 columns (dy2, dx2, c32); output pixel
// (4 ty + 2 dy + dy2, 4 tx + 2 dx + dx2).  One block per (ty, prompt): 64 tokens x 16 sub-pixels x 3 masks staged
// in shared memory and written as 4 full image rows per mask.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_logits_kernel(const bf16* __restrict__ up2, const bf16* __restrict__ hyper, float* __restrict__ low) {
  const int ty = blockIdx.x, p = blockIdx.y;
  __shared__ float hy[3][32];
  __shared__ float stage[3][4][LOW];
  if (threadIdx.x < 96) hy[threadIdx.x >> 5][threadIdx.x & 31] =
      __bfloat162float(hyper[((size_t)p * 4 + 1 + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31)]);
  __syncthreads();
  // 64 tokens x 4 (dy,dx) x 4 (dy2,dx2) = 1024 items, 4 per thread; item = (tx, dydx, s): consecutive threads read
  // consecutive 64-byte quarter rows
  for (int it = threadIdx.x; it < 1024; it += 256) {
    const int s = it & 3, dydx = (it >> 2) & 3, tx = it >> 4;
    const size_t row = (((size_t)p * IMG_TOK + ty * 64 + tx) << 2) + dydx;
    const uint4* src = reinterpret_cast<const uint4*>(up2 + row * 128 + s * 32);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 raw = __ldg(src + g);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = unpack_bf16(w[i]);
        const int c = g * 8 + 2 * i;
        a0 = fmaf(v.x, hy[0][c], a0); a0 = fmaf(v.y, hy[0][c + 1], a0);
        a1 = fmaf(v.x, hy[1][c], a1); a1 = fmaf(v.y, hy[1][c + 1], a1);
        a2 = fmaf(v.x, hy[2][c], a2); a2 = fmaf(v.y, hy[2][c + 1], a2);
      }
    }
    const int yy = 2 * (dydx >> 1) + (s >> 1), xx = 4 * tx + 2 * (dydx & 1) + (s & 1);
    stage[0][yy][xx] = a0;
    stage[1][yy][xx] = a1;
    stage[2][yy][xx] = a2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 4 * LOW; i += 256) {
    const int m = i / (4 * LOW), r = (i / LOW) & 3, x = i & (LOW - 1);
    low[(((size_t)p * 3 + m) * LOW + 4 * ty + r) * LOW + x] = stage[m][r][x];
  }
}

// ---------------------------------------------------------------------------------------------
// Fused tail of the mask decoder's output up-scaling (mask_decoder.py:139-164): for one (prompt, token row ty) the block
// takes the 256 un-shuffled rows of ConvTranspose #1 (64 tokens x 4 sub-pixels, 64 channels; bias already added by the
// GEMM), applies LayerNorm2d(64) + GELU, multiplies by ConvTranspose #2 as a [256 x 64] x [64 x 128] product on
// mma.sync (columns = 4 sub-sub-pixels x 32 channels), adds its bias, GELU, and contracts the 32 channels with the three
// hyper-network vectors of the prompt — 4 rows x 256 pixels of low-res logits per mask leave the block.  Replaces
// ln64_gelu + a 4 M x 128 x 64 GEMM + mask_logits: 0.27 + 0.54 + 1.07 + 1.07 GB of HBM traffic per 256 prompts become
// one 0.27 GB read, and the ConvTranspose #2 activations are never rounded to bf16.
// Shared memory: X [256 x 64] and W2 [128 x 64] bf16 with the 16-byte chunks of a row XOR-swizzled by (row & 7)
// (ldmatrix conflict-free), the output staging tile, and the small vectors.
// ---------------------------------------------------------------------------------------------
constexpr int UPF_X_BYTES = 256 * 128, UPF_W_BYTES = 128 * 128, UPF_STAGE_BYTES = 3 * 4 * LOW * 4;
constexpr int UPF_VEC_FLOATS = 64 + 64 + 128 + 96;
constexpr int UPF_SMEM = UPF_X_BYTES + UPF_W_BYTES + UPF_STAGE_BYTES + UPF_VEC_FLOATS * 4;

__device__ __forceinline__ void mma16816_full(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2)
upscale_logits_kernel(const bf16* __restrict__ u1, const bf16* __restrict__ gamma, const bf16* __restrict__ beta, float eps,
                      const bf16* __restrict__ w2, const bf16* __restrict__ b2, const bf16* __restrict__ hyper,
                      float* __restrict__ low) {
  extern __shared__ __align__(128) uint8_t upf_smem[];
  const int ty = blockIdx.x, p = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* xs = upf_smem;
  uint8_t* ws = upf_smem + UPF_X_BYTES;
  float (*stage)[4][LOW] = reinterpret_cast<float (*)[4][LOW]>(upf_smem + UPF_X_BYTES + UPF_W_BYTES);
  float* gam = reinterpret_cast<float*>(upf_smem + UPF_X_BYTES + UPF_W_BYTES + UPF_STAGE_BYTES);
  float* bet = gam + 64;
  float* bias2 = bet + 64;
  float* hy = bias2 + 128;   // [3][32]
  const uint32_t xs_u = smem_u32(xs), ws_u = smem_u32(ws);
  // ---- loads: the block's 256 u1 rows are contiguous (32 KB); W2 is 16 KB (L2-resident across blocks)
  const bf16* src = u1 + ((size_t)p * IMG_TOK + (size_t)ty * 64) * 256;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = tid + i * 256, row = q >> 3, c = q & 7;
    cp_async16(xs_u + row * 128 + ((c ^ (row & 7)) << 4), src + (size_t)q * 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = tid + i * 256, row = q >> 3, c = q & 7;
    cp_async16(ws_u + row * 128 + ((c ^ (row & 7)) << 4), w2 + (size_t)q * 8);
  }
  cp_async_commit();
  if (tid < 64) {
    gam[tid] = __bfloat162float(gamma[tid]);
    bet[tid] = __bfloat162float(beta[tid]);
  }
  if (tid < 128) bias2[tid] = __bfloat162float(b2[tid]);
  if (tid >= 128 && tid < 224) {
    const int i = tid - 128;
    hy[i] = __bfloat162float(hyper[((size_t)p * 4 + 1 + (i >> 5)) * 32 + (i & 31)]);
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- LayerNorm2d(64) + GELU in place: 8 lanes per row, 8 channels per lane (arithmetic of ln64_gelu_kernel)
  {
    const int part = tid & 7;
#pragma unroll 2
    for (int pass = 0; pass < 8; ++pass) {
      const int row = pass * 32 + (tid >> 3);
      uint4* cell = reinterpret_cast<uint4*>(xs + row * 128 + ((part ^ (row & 7)) << 4));
      const uint4 raw = *cell;
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      float x[8], sm = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = unpack_bf16(w[i]);
        x[2 * i] = v.x;
        x[2 * i + 1] = v.y;
        sm += v.x + v.y;
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
      const float mean = sm * (1.f / 64.f);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] -= mean;
        ss = fmaf(x[i], x[i], ss);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss * (1.f / 64.f) + eps);
      uint32_t o4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = part * 8 + 2 * i;
        const float2 y = gelu2(make_float2(fmaf(x[2 * i] * rstd, gam[c], bet[c]), fmaf(x[2 * i + 1] * rstd, gam[c + 1], bet[c + 1])));
        o4[i] = pack_bf16(y.x, y.y);
      }
      *cell = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
  }
  __syncthreads();
  // ---- [32 rows of this warp] x [64 ch] x W2^T, two 64-column halves; epilogue per half
  const int mi = lane >> 3, mr = lane & 7;
  const int qr = lane >> 2, qc = (lane & 3) * 2;
#pragma unroll 1
  for (int nh = 0; nh < 2; ++nh) {
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        // matrices: rows 0-7 / k 0-7, rows 8-15 / k 0-7, rows 0-7 / k 8-15, rows 8-15 / k 8-15  ->  a0..a3
        const int row = warp * 32 + mt * 16 + (mi & 1) * 8 + mr, kc = ks * 2 + (mi >> 1);
        ldmatrix_x4(xs_u + row * 128 + ((kc ^ (row & 7)) << 4), a[mt]);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        // matrices: n-tile 2np / k 0-7, 2np / k 8-15, 2np+1 / k 0-7, 2np+1 / k 8-15  ->  (b0, b1) of two n-tiles
        const int n = nh * 64 + (np * 2 + (mi >> 1)) * 8 + mr, kc = ks * 2 + (mi & 1);
        uint32_t b[4];
        ldmatrix_x4(ws_u + n * 128 + ((kc ^ (n & 7)) << 4), b);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816_full(acc[mt][np * 2], a[mt], b[0], b[1]);
          mma16816_full(acc[mt][np * 2 + 1], a[mt], b[2], b[3]);
        }
      }
    }
    // lane holds, per (mt, row half rh, n-tile nt): columns nh*64 + nt*8 + qc, +1  ->  sub-sub-pixel s = nh*2 + nt/4,
    // channels (nt & 3)*8 + qc, +1.  Channel group t outermost: the six hyper-network weights of a lane's two
    // channels are read once per t and serve all eight (mt, rh, pl) dot products.
    float d[2][2][2][3];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) d[mt][rh][pl][0] = d[mt][rh][pl][1] = d[mt][rh][pl][2] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = t * 8 + qc;
      const float2 h0 = *reinterpret_cast<const float2*>(hy + c), h1 = *reinterpret_cast<const float2*>(hy + 32 + c),
                   h2 = *reinterpret_cast<const float2*>(hy + 64 + c);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        const int nt = pl * 4 + t;
        const float2 bb = *reinterpret_cast<const float2*>(bias2 + nh * 64 + nt * 8 + qc);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const float2 g = gelu2(make_float2(acc[mt][nt][rh * 2] + bb.x, acc[mt][nt][rh * 2 + 1] + bb.y));
            float* dd = d[mt][rh][pl];
            dd[0] = fmaf(g.x, h0.x, dd[0]); dd[0] = fmaf(g.y, h0.y, dd[0]);
            dd[1] = fmaf(g.x, h1.x, dd[1]); dd[1] = fmaf(g.y, h1.y, dd[1]);
            dd[2] = fmaf(g.x, h2.x, dd[2]); dd[2] = fmaf(g.y, h2.y, dd[2]);
          }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          float d0 = d[mt][rh][pl][0], d1 = d[mt][rh][pl][1], d2 = d[mt][rh][pl][2];
#pragma unroll
          for (int o = 1; o < 4; o <<= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, o);
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
          }
          const int R = warp * 32 + mt * 16 + rh * 8 + qr;   // row of the block: token tx = R / 4, sub-pixel R % 4
          const int s2 = nh * 2 + pl;
          const int yy = 2 * ((R >> 1) & 1) + (s2 >> 1), xx = 4 * (R >> 2) + 2 * (R & 1) + (s2 & 1);
          const int m = lane & 3;
          if (m < 3) stage[m][yy][xx] = m == 0 ? d0 : (m == 1 ? d1 : d2);
        }
  }
  __syncthreads();
  for (int i = tid; i < 3 * 4 * LOW; i += 256) {
    const int m = i / (4 * LOW), r = (i / LOW) & 3, x = i & (LOW - 1);
    low[(((size_t)p * 3 + m) * LOW + 4 * ty + r) * LOW + x] = stage[m][r][x];
  }
}

// ---------------------------------------------------------------------------------------------
// 4x bilinear up-sampling (align_corners = False) of a 256 x 256 logit map, evaluated per output pixel the way
// ATen's upsample_bilinear2d does: src = 0.25 (dst + 0.5) - 0.5 clamped at 0, neighbours i0, i0 + (i0 < 255).
// ---------------------------------------------------------------------------------------------
struct Tap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Tap tap_of(int d) {
  float s = 0.25f * ((float)d + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  Tap t;
  t.i0 = (int)s;
  t.i1 = t.i0 + (t.i0 < LOW - 1 ? 1 : 0);
  t.l1 = s - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}
__device__ __forceinline__ float up_at(const float* r0, const float* r1, const Tap& ty, const Tap& tx) {
  return ty.l0 * (tx.l0 * r0[tx.i0] + tx.l1 * r0[tx.i1]) + ty.l1 * (tx.l0 * r1[tx.i0] + tx.l1 * r1[tx.i1]);
}

// stats[c] = {area, count(> thr + off), count(> thr - off), max(1023 - x), max(1023 - y), max x, max y} over logits > thr
// (zero-initialised by the launcher; an empty mask is area == 0).  grid (8 row chunks, n_cand); 256 threads.
__global__ void __launch_bounds__(256)
mask_stats_kernel(const float* __restrict__ low, const int* __restrict__ cand, int* __restrict__ stats, float thr,
                  float off) {
  const int c = blockIdx.y, chunk = blockIdx.x;
  const float* src = low + (size_t)(cand ? cand[c] : c) * LOW * LOW;
  constexpr int ROWS = HI / 8;             // 128 hi-res rows per block
  constexpr int LROWS = ROWS / 4 + 2;      // low-res rows they touch
  __shared__ float tile[LROWS][LOW];
  const int lr0 = max(chunk * (ROWS / 4) - 1, 0);
  for (int i = threadIdx.x; i < LROWS * LOW; i += 256) {
    const int r = min(lr0 + i / LOW, LOW - 1);
    tile[i / LOW][i & (LOW - 1)] = src[(size_t)r * LOW + (i & (LOW - 1))];
  }
  __syncthreads();
  int area = 0, hi_c = 0, lo_c = 0, mnx = 0, mny = 0, mxx = 0, mxy = 0;
  // thread owns 4 hi-res columns (the taps of a column are computed once).  Four consecutive hi-res rows share their two
  // low-res rows, so the horizontal half of the interpolation is kept per low-res row (h0 / h1, same expression as
  // up_at) and a hi-res pixel costs one vertical blend and three compares; the box is tracked as "any pixel in this
  // column / row" flags instead of four max updates per pixel.
  const int x0 = threadIdx.x * 4;
  Tap tx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tx[j] = tap_of(x0 + j);
  const float thr_hi = thr + off, thr_lo = thr - off;
  unsigned colany = 0;
  // Hi-res rows 4k+2 .. 4k+5 blend low-res rows k and min(k+1, 255) with weights l1 = 1/8, 3/8, 5/8, 7/8 (exact in fp32;
  // rows 0, 1 clamp to l1 = 0 on row 0 = "group -1").  The 128 rows of a chunk are 33 groups, the first and the last
  // one half inside the chunk; hA / hB are the horizontally interpolated low-res rows (the expression of up_at).
  auto hrow = [&](int lr, float* h) {
    const float* rr = tile[lr - lr0];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = tx[j].l0 * rr[tx[j].i0] + tx[j].l1 * rr[tx[j].i1];
  };
  auto rows = [&](const float* hA, const float* hB, int k, int i_lo, int i_hi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < i_lo || i >= i_hi) continue;
      const float l1 = k < 0 ? 0.f : 0.125f + 0.25f * (float)i, l0 = 1.f - l1;
      const int y = 4 * k + 2 + i;
      bool rowany = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = l0 * hA[j] + l1 * hB[j];
        hi_c += v > thr_hi;
        lo_c += v > thr_lo;
        const bool in = v > thr;
        area += in;
        colany |= (unsigned)in << j;
        rowany |= in;
      }
      if (rowany) {
        mny = max(mny, HI - 1 - y);
        mxy = max(mxy, y);
      }
    }
  };
  float hA[4], hB[4];
  int k = chunk * (ROWS / 4) - 1;
  hrow(max(k, 0), hA);
  hrow(min(k + 1, LOW - 1), hB);
  rows(hA, hB, k, 2, 4);
#pragma unroll 1
  for (int g = 1; g < ROWS / 4; ++g) {
    ++k;
#pragma unroll
    for (int j = 0; j < 4; ++j) hA[j] = hB[j];
    hrow(min(k + 1, LOW - 1), hB);
    rows(hA, hB, k, 0, 4);
  }
  ++k;
#pragma unroll
  for (int j = 0; j < 4; ++j) hA[j] = hB[j];
  hrow(min(k + 1, LOW - 1), hB);
  rows(hA, hB, k, 0, 2);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if ((colany >> j) & 1u) {
      mnx = max(mnx, HI - 1 - (x0 + j));
      mxx = max(mxx, x0 + j);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    area += __shfl_xor_sync(0xffffffffu, area, o);
    hi_c += __shfl_xor_sync(0xffffffffu, hi_c, o);
    lo_c += __shfl_xor_sync(0xffffffffu, lo_c, o);
    mnx = max(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mny = max(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if ((threadIdx.x & 31) == 0) {
    int* st = stats + (size_t)c * 8;
    if (lo_c) atomicAdd(st + 2, lo_c);
    if (hi_c) atomicAdd(st + 1, hi_c);
    if (area) {
      atomicAdd(st + 0, area);
      atomicMax(st + 3, mnx);
      atomicMax(st + 4, mny);
      atomicMax(st + 5, mxx);
      atomicMax(st + 6, mxy);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// greedy NMS: boxes sorted by score (descending, stable); keep[i] = 1 unless an earlier kept box overlaps it with
// IoU > thr.  One block; the kept boxes are visited in order, every thread clears its share of the later ones.
// ---------------------------------------------------------------------------------------------
constexpr int NMS_MAX = 4096;
__global__ void __launch_bounds__(1024)
box_nms_kernel(const float* __restrict__ boxes, int n, float thr, int* __restrict__ keep) {
  __shared__ unsigned char dead[NMS_MAX];
  for (int i = threadIdx.x; i < n; i += blockDim.x) dead[i] = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (dead[i]) continue;   // block-uniform: `dead` only changes between the barriers below
    const float ax0 = boxes[4 * i], ay0 = boxes[4 * i + 1], ax1 = boxes[4 * i + 2], ay1 = boxes[4 * i + 3];
    const float aa = (ax1 - ax0) * (ay1 - ay0);
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      const float bx0 = boxes[4 * j], by0 = boxes[4 * j + 1], bx1 = boxes[4 * j + 2], by1 = boxes[4 * j + 3];
      const float w = fmaxf(fminf(ax1, bx1) - fmaxf(ax0, bx0), 0.f), h = fmaxf(fminf(ay1, by1) - fmaxf(ay0, by0), 0.f);
      const float inter = w * h;
      const float iou = inter / (aa + (bx1 - bx0) * (by1 - by0) - inter);
      if (iou > thr) dead[j] = 1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) keep[i] = dead[i] ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// soft proposal: out[i, j] = sum over the antialias window of w_y w_x [up(y, x) > thr], window of output i:
// inputs [max(4i - 2, 0), min(4i + 6, 1024)), triangle weights 1 - |(y + 0.5 - (4i + 2)) / 4| normalised by their sum
// (ATen _compute_indices_weights_aa for scale 4).  grid (256 output rows, K); 256 threads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void aa_window(int i, int& lo, int& n, float* w) {
  lo = max(4 * i - 2, 0);
  const int hi = min(4 * i + 6, HI);
  n = hi - lo;
  const float center = 4.f * ((float)i + 0.5f);
  float tot = 0.f;
  for (int j = 0; j < 8; ++j) {
    float v = 0.f;
    if (j < n) v = fmaxf(0.f, 1.f - fabsf(((float)(j + lo) - center + 0.5f) * 0.25f));
    w[j] = v;
    tot += v;
  }
  const float inv = 1.f / tot;
  for (int j = 0; j < 8; ++j) w[j] *= inv;
}

__global__ void __launch_bounds__(256)
mask_soft_kernel(const float* __restrict__ low, const int* __restrict__ cand, bf16* __restrict__ out, float thr) {
  const int i = blockIdx.x, c = blockIdx.y;
  const float* src = low + (size_t)cand[c] * LOW * LOW;
  __shared__ float tile[4][LOW];
  __shared__ float colw[HI];
  int ylo, yn;
  float wy[8];
  aa_window(i, ylo, yn, wy);
  const int lr0 = tap_of(ylo).i0;
  for (int t = threadIdx.x; t < 4 * LOW; t += 256) {
    const int r = min(lr0 + t / LOW, LOW - 1);
    tile[t / LOW][t & (LOW - 1)] = src[(size_t)r * LOW + (t & (LOW - 1))];
  }
  __syncthreads();
  for (int x = threadIdx.x; x < HI; x += 256) {
    const Tap tx = tap_of(x);
    float a = 0.f;
    for (int j = 0; j < yn; ++j) {
      const Tap ty = tap_of(ylo + j);
      const float v = up_at(tile[ty.i0 - lr0], tile[ty.i1 - lr0], ty, tx);
      a += v > thr ? wy[j] : 0.f;
    }
    colw[x] = a;
  }
  __syncthreads();
  {
    const int j = threadIdx.x;
    int xlo, xn;
    float wx[8];
    aa_window(j, xlo, xn, wx);
    float a = 0.f;
    for (int t = 0; t < xn; ++t) a = fmaf(wx[t], colw[xlo + t], a);
    out[((size_t)c * LOW + i) * LOW + j] = __float2bfloat16_rn(a);
  }
}

// binary 1024 x 1024 masks (uint8 0/1) of the listed candidates.  grid (1024 rows / 4, K); 256 threads x 4 columns.
__global__ void __launch_bounds__(256)
mask_binarize_kernel(const float* __restrict__ low, const int* __restrict__ cand, unsigned char* __restrict__ out,
                     float thr) {
  const int c = blockIdx.y;
  const float* src = low + (size_t)cand[c] * LOW * LOW;
  __shared__ float tile[3][LOW];
  const int y0 = blockIdx.x * 4;
  const int lr0 = tap_of(y0).i0;
  for (int t = threadIdx.x; t < 3 * LOW; t += 256) {
    const int r = min(lr0 + t / LOW, LOW - 1);
    tile[t / LOW][t & (LOW - 1)] = src[(size_t)r * LOW + (t & (LOW - 1))];
  }
  __syncthreads();
  const int x0 = threadIdx.x * 4;
  Tap tx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tx[j] = tap_of(x0 + j);
  for (int r = 0; r < 4; ++r) {
    const Tap ty = tap_of(y0 + r);
    uchar4 o;
    o.x = up_at(tile[ty.i0 - lr0], tile[ty.i1 - lr0], ty, tx[0]) > thr;
    o.y = up_at(tile[ty.i0 - lr0], tile[ty.i1 - lr0], ty, tx[1]) > thr;
    o.z = up_at(tile[ty.i0 - lr0], tile[ty.i1 - lr0], ty, tx[2]) > thr;
    o.w = up_at(tile[ty.i0 - lr0], tile[ty.i1 - lr0], ty, tx[3]) > thr;
    *reinterpret_cast<uchar4*>(out + ((size_t)c * HI + y0 + r) * HI + x0) = o;
  }
}

}  // namespace
}  // namespace llmseg

using namespace llmseg;

#define AMG_LAUNCHED()            \
  LLMSEG_CUDA(cudaGetLastError()); \
  g_launches.fetch_add(1);         \
  return 0

extern "C" int llmseg_point_tokens(const float* points, int n_prompts, const float* gauss, const void* out_tokens,
                                   const void* point_embed, const void* not_a_point, float img_size, void* tokens,
                                   void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(points && gauss && out_tokens && point_embed && not_a_point && tokens && n_prompts > 0 && img_size > 0.f,
                 LLMSEG_EARG, "llmseg_point_tokens: bad arguments");
  point_tokens_kernel<<<n_prompts, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      points, gauss, static_cast<const bf16*>(out_tokens), static_cast<const bf16*>(point_embed),
      static_cast<const bf16*>(not_a_point), static_cast<bf16*>(tokens), 1.f / img_size);
  AMG_LAUNCHED();
}

extern "C" int llmseg_tok2img_attention(const void* q, int ldq, const void* k, int ldk, long long k_batch_stride,
                                        const void* v, int ldv, long long v_batch_stride, void* out, int ldo,
                                        int n_prompts, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(q && k && v && out && n_prompts > 0, LLMSEG_EARG, "llmseg_tok2img_attention: bad arguments");
  LLMSEG_REQUIRE(ldk % 8 == 0 && ldv % 8 == 0 && k_batch_stride % 8 == 0 && v_batch_stride % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                 LLMSEG_EALIGN, "llmseg_tok2img_attention: k / v rows must be 16-byte aligned");
  LLMSEG_REQUIRE(ldq % 2 == 0 && ldo % 2 == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out)) & 3) == 0,
                 LLMSEG_EALIGN, "llmseg_tok2img_attention: q / out rows must be 4-byte aligned");
  const char* v1 = getenv("LLMSEG_T2I_V1");   // thread-per-key CUDA-core kernel (A/B runs and tests)
  if (v1 != nullptr && atoi(v1) != 0) {
    tok2img_attn_kernel<<<dim3(n_prompts, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, k_batch_stride, static_cast<const bf16*>(v),
        ldv, v_batch_stride, static_cast<bf16*>(out), ldo);
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      LLMSEG_CUDA(cudaFuncSetAttribute(tok2img_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2I_SMEM));
      attr_set = true;
    }
    tok2img_attn_mma_kernel<<<n_prompts, 256, T2I_SMEM, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, k_batch_stride, static_cast<const bf16*>(v),
        ldv, v_batch_stride, static_cast<bf16*>(out), ldo);
  }
  AMG_LAUNCHED();
}

extern "C" int llmseg_img2tok_attention(const void* q, int ldq, long long q_batch_stride, const void* k, int ldk,
                                        const void* v, int ldv, void* out, int ldo, int n_prompts, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(q && k && v && out && n_prompts > 0, LLMSEG_EARG, "llmseg_img2tok_attention: bad arguments");
  LLMSEG_REQUIRE(ldq % 8 == 0 && ldo % 8 == 0 && q_batch_stride % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                 LLMSEG_EALIGN, "llmseg_img2tok_attention: q / out rows must be 16-byte aligned");
  img2tok_attn_kernel<<<dim3(IMG_TOK / I2T_TOK, n_prompts), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(q), ldq, q_batch_stride, static_cast<const bf16*>(k), ldk, static_cast<const bf16*>(v),
      ldv, static_cast<bf16*>(out), ldo);
  AMG_LAUNCHED();
}

extern "C" int llmseg_ln64_gelu(const void* in, void* out, const void* gamma, const void* beta, long long rows,
                                float eps, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(in && out && gamma && beta && rows > 0, LLMSEG_EARG, "llmseg_ln64_gelu: bad arguments");
  LLMSEG_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, LLMSEG_EALIGN,
                 "llmseg_ln64_gelu: buffers must be 16-byte aligned");
  const long long blocks = (rows + 31) / 32;
  ln64_gelu_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), static_cast<bf16*>(out), static_cast<const bf16*>(gamma),
      static_cast<const bf16*>(beta), rows, eps);
  AMG_LAUNCHED();
}

extern "C" int llmseg_mask_logits(const void* up2, const void* hyper, int n_prompts, float* low_res, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(up2 && hyper && low_res && n_prompts > 0, LLMSEG_EARG, "llmseg_mask_logits: bad arguments");
  mask_logits_kernel<<<dim3(64, n_prompts), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(up2), static_cast<const bf16*>(hyper), low_res);
  AMG_LAUNCHED();
}

extern "C" int llmseg_upscale_logits(const void* up1, const void* gamma, const void* beta, float eps, const void* w2,
                                     const void* b2, const void* hyper, int n_prompts, float* low_res, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(up1 && gamma && beta && w2 && b2 && hyper && low_res && n_prompts > 0, LLMSEG_EARG,
                 "llmseg_upscale_logits: bad arguments");
  LLMSEG_REQUIRE(((reinterpret_cast<uintptr_t>(up1) | reinterpret_cast<uintptr_t>(w2)) & 15) == 0, LLMSEG_EALIGN,
                 "llmseg_upscale_logits: up1 / w2 must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    LLMSEG_CUDA(cudaFuncSetAttribute(upscale_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UPF_SMEM));
    attr_set = true;
  }
  upscale_logits_kernel<<<dim3(64, n_prompts), 256, UPF_SMEM, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(up1), static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), eps,
      static_cast<const bf16*>(w2), static_cast<const bf16*>(b2), static_cast<const bf16*>(hyper), low_res);
  AMG_LAUNCHED();
}

extern "C" int llmseg_mask_stats(const float* low_res, const int32_t* cand, int n_cand, float threshold, float offset,
                                 int32_t* stats, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(low_res && stats && n_cand > 0, LLMSEG_EARG, "llmseg_mask_stats: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LLMSEG_CUDA(cudaMemsetAsync(stats, 0, (size_t)n_cand * 8 * sizeof(int32_t), s));
  mask_stats_kernel<<<dim3(8, n_cand), 256, 0, s>>>(low_res, cand, stats, threshold, offset);
  AMG_LAUNCHED();
}

extern "C" int llmseg_box_nms(const float* boxes_sorted, int n, float iou_threshold, int32_t* keep, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(boxes_sorted && keep && n > 0 && n <= NMS_MAX, LLMSEG_ESHAPE, "llmseg_box_nms: n=%d (1..%d)", n, NMS_MAX);
  box_nms_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(boxes_sorted, n, iou_threshold, keep);
  AMG_LAUNCHED();
}

extern "C" int llmseg_mask_soft(const float* low_res, const int32_t* cand, int n_cand, float threshold, void* out,
                                void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(low_res && cand && out && n_cand > 0, LLMSEG_EARG, "llmseg_mask_soft: bad arguments");
  mask_soft_kernel<<<dim3(LOW, n_cand), 256, 0, static_cast<cudaStream_t>(stream)>>>(low_res, cand,
                                                                                       static_cast<bf16*>(out), threshold);
  AMG_LAUNCHED();
}

extern "C" int llmseg_mask_binarize(const float* low_res, const int32_t* cand, int n_cand, float threshold,
                                    uint8_t* out, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(low_res && cand && out && n_cand > 0, LLMSEG_EARG, "llmseg_mask_binarize: bad arguments");
  mask_binarize_kernel<<<dim3(HI / 4, n_cand), 256, 0, static_cast<cudaStream_t>(stream)>>>(low_res, cand, out, threshold);
  AMG_LAUNCHED();
}
