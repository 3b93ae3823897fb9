// llmseg_b200 — mask-proposal selector kernels (HBM / latency bound; warp-shuffle reductions).
//
//   maskpool      fused  bilinear-upsample(64²→256²) ∘ mask-pooling   (reference LISA.py:201-218,350-361)
//                 via the exact adjoint form  W·(U·E) = (Uᵀ·W)·E  (SURVEY §A.4): each soft mask is
//                 read ONCE (K·65536·2 B), resampled to 64² with the transposed bilinear taps, then
//                 contracted with the 64² embedding.  The 256-channel 256² upsampled tensor
//                 (67 MB fp32 + 33 MB bf16 per image in the reference) never exists.
//   small_attn    8-head × 32-dim attention among ≤128 mask tokens / from the single text token
//                 (reference transformer.py:319-341) with the reference's bf16 rounding points
//   select        IoU head output layer + sigmoid, cosine similarity, argmax
//                 (reference LISA.py:387-408, training.py:627-629)
//   losses        KL-align / weighted-MSE (reference loss.py:50-94) and dice / BCE (loss.py:4-47)
#include <atomic>

#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
namespace {

// ---------------------------------------------------------------------------------------------
// maskpool, stage 1: adjoint resampling.  wt[k, jy, jx] = Σ_{y,x} a[jy,y]·a[jx,x]·w[k,y,x]
// a[j, 4j-2+t] = {.125,.375,.625,.875,.875,.625,.375,.125}[t] with the align_corners=False edge
// clamps a[0,0]=a[0,1]=1 and a[63,254]=a[63,255]=1 (F.interpolate bilinear, scale 4).
// grid (8 row-groups, n_masks); block 256 threads (one per hi-res column).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tap_w(int j, int y) {
  // weight of low-res cell j for hi-res pixel y (0 when out of the 8-tap support)
  const int t = y - (4 * j - 2);
  if (t < 0 || t > 7 || y < 0 || y > 255) return 0.f;
  if ((j == 0 && y < 2) || (j == 63 && y > 253)) return 1.f;
  return t < 4 ? 0.125f + 0.25f * (float)t : 0.125f + 0.25f * (float)(7 - t);
}

__global__ void __launch_bounds__(256)
maskpool_adjoint_kernel(const bf16* __restrict__ segs, float* __restrict__ wt,
                        float* __restrict__ part_sum) {
  const int k = blockIdx.y;
  const int grp = blockIdx.x;  // low-res rows [8*grp, 8*grp+8)
  const int x = threadIdx.x;
  const bf16* seg = segs + (size_t)k * 65536;
  __shared__ float tmp[256 + 8];
  __shared__ float red[8];
  float own_sum = 0.f;
  for (int jj = 0; jj < 8; ++jj) {
    const int jy = grp * 8 + jj;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int y = 4 * jy - 2 + t;
      if (y >= 0 && y < 256) {
        const float v = __bfloat162float(seg[(size_t)y * 256 + x]);
        acc += tap_w(jy, y) * v;
        if (t >= 2 && t < 6) own_sum += v;  // rows 4jy..4jy+3 are owned by this low-res row
      }
    }
    __syncthreads();
    tmp[x + 2] = acc;
    if (x < 2) { tmp[x] = 0.f; tmp[258 + x] = 0.f; }
    __syncthreads();
    if (x < 64) {
      float o = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) o += tap_w(x, 4 * x - 2 + t) * tmp[4 * x + t];
      wt[(size_t)k * 4096 + jy * 64 + x] = o;
    }
  }
  own_sum = warp_sum(own_sum);
  if ((x & 31) == 0) red[x >> 5] = own_sum;
  __syncthreads();
  if (x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    part_sum[k * 8 + grp] = s;
  }
}

// stage 2: out[k, c] = bf16( (Σ_cell wt[k,cell]·E[img(k), cell, c]) / (Σ_p w[k,p] + 1e-8) )
// E is token-major [B, 4096, 256] bf16 (the SAM neck output in NHWC).  The 4096-cell contraction is
// split 8 ways over blockIdx.y (512 cells each) into fp32 partials so 8x more blocks stream E;
// stage 3 sums the partials in a fixed order (deterministic) and normalises.
// grid (ceil(n_masks/8), 8); block 256 threads (one per channel).
__global__ void __launch_bounds__(256)
maskpool_apply_kernel(const float* __restrict__ wt, const bf16* __restrict__ emb,
                      const int* __restrict__ mask_image, float* __restrict__ partial, int n_masks) {
  const int c = threadIdx.x;
  const int k0 = blockIdx.x * 8;
  const int cell0 = blockIdx.y * 512;
  __shared__ float ws[8][512];
  float acc[8];
  int img[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = 0.f;
    img[i] = (k0 + i < n_masks) ? mask_image[k0 + i] : -1;
  }
  const bool same = (img[7] == img[0] || img[7] < 0);  // fast path: one image per block
  for (int i = threadIdx.x; i < 8 * 512; i += 256) {
    const int m = i >> 9, cc = i & 511;
    ws[m][cc] = (k0 + m < n_masks) ? wt[(size_t)(k0 + m) * 4096 + cell0 + cc] : 0.f;
  }
  __syncthreads();
  if (same) {
    const bf16* e = emb + ((size_t)img[0] * 4096 + cell0) * 256 + c;
#pragma unroll 4
    for (int cc = 0; cc < 512; ++cc) {
      const float ev = __bfloat162float(e[(size_t)cc * 256]);
#pragma unroll
      for (int m = 0; m < 8; ++m) acc[m] = fmaf(ws[m][cc], ev, acc[m]);
    }
  } else {
    for (int m = 0; m < 8; ++m) {
      if (img[m] < 0) continue;
      const bf16* e = emb + ((size_t)img[m] * 4096 + cell0) * 256 + c;
      for (int cc = 0; cc < 512; ++cc) acc[m] = fmaf(ws[m][cc], __bfloat162float(e[(size_t)cc * 256]), acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m)
    if (k0 + m < n_masks) partial[((size_t)blockIdx.y * n_masks + k0 + m) * 256 + c] = acc[m];
}

// stage 3: fixed-order sum of the 8 partials, normalise (LISA.py:211-213), round once
__global__ void __launch_bounds__(256)
maskpool_final_kernel(const float* __restrict__ partial, const float* __restrict__ part_sum,
                      bf16* __restrict__ out, int n_masks) {
  const int k = blockIdx.x, c = threadIdx.x;
  float acc = 0.f, s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc += partial[((size_t)i * n_masks + k) * 256 + c];
    s += part_sum[k * 8 + i];
  }
  out[(size_t)k * 256 + c] = __float2bfloat16_rn(acc / (s + 1e-8f));  // fp32 up to the one bf16 store
}

// ---------------------------------------------------------------------------------------------
// small attention: per (image, head) block; queries q_off[b]..q_off[b+1], keys kv_off[b]..kv_off[b+1].
// fp32 arithmetic on the bf16 q/k/v (scores, softmax and P·V never round); one bf16 rounding at the store.
// ---------------------------------------------------------------------------------------------
constexpr int SA_MAX = 128;  // max keys per image
constexpr int SA_HD = 32;

__global__ void __launch_bounds__(128)
small_attn_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, int ldk,
                  const bf16* __restrict__ v, int ldv, bf16* __restrict__ out, int ldo,
                  const int* __restrict__ q_off, const int* __restrict__ kv_off, float inv_sqrt_d) {
  const int b = blockIdx.x, h = blockIdx.y;
  const int q0 = q_off[b], nq = q_off[b + 1] - q0;
  const int k0 = kv_off[b], nk = kv_off[b + 1] - k0;
  __shared__ float ks[SA_MAX][SA_HD + 1];
  __shared__ float vs[SA_MAX][SA_HD + 1];
  for (int i = threadIdx.x; i < nk * SA_HD; i += blockDim.x) {
    const int r = i / SA_HD, d = i % SA_HD;
    ks[r][d] = __bfloat162float(k[(size_t)(k0 + r) * ldk + h * SA_HD + d]);
    vs[r][d] = __bfloat162float(v[(size_t)(k0 + r) * ldv + h * SA_HD + d]);
  }
  __syncthreads();
  for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {
    float qv[SA_HD];
#pragma unroll
    for (int d = 0; d < SA_HD; ++d) qv[d] = __bfloat162float(q[(size_t)(q0 + qi) * ldq + h * SA_HD + d]);
    float mx = -INFINITY;
    for (int j = 0; j < nk; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < SA_HD; ++d) s = fmaf(qv[d], ks[j][d], s);
      s *= inv_sqrt_d;
      mx = fmaxf(mx, s);
    }
    float den = 0.f;
    for (int j = 0; j < nk; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < SA_HD; ++d) s = fmaf(qv[d], ks[j][d], s);
      s *= inv_sqrt_d;
      den += __expf(s - mx);
    }
    float o[SA_HD];
#pragma unroll
    for (int d = 0; d < SA_HD; ++d) o[d] = 0.f;
    const float inv = 1.f / den;
    for (int j = 0; j < nk; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < SA_HD; ++d) s = fmaf(qv[d], ks[j][d], s);
      s *= inv_sqrt_d;
      const float p = __expf(s - mx) * inv;
#pragma unroll
      for (int d = 0; d < SA_HD; ++d) o[d] = fmaf(p, vs[j][d], o[d]);
    }
    bf16* orow = out + (size_t)(q0 + qi) * ldo + h * SA_HD;
#pragma unroll
    for (int d = 0; d < SA_HD; d += 2)
      *reinterpret_cast<uint32_t*>(orow + d) = pack_bf16(o[d], o[d + 1]);
  }
}

// ---------------------------------------------------------------------------------------------
// select: one block per CONVERSATION c of group g = conv_group[c] (identity when NULL), one warp per mask
// token (looping).  The mask tokens of a group went through the selector blocks with the group's FIRST
// conversation as their text key (reference LISA.py:359-391: `sam_segs_feature_list[b][0]`), and every
// conversation of the group is scored against them (LISA.py:397-403: `pred_embeddings[b]` is [C,256]):
//   sim[c,k] = (t_c/‖t_c‖) · (e_k/‖e_k‖)                         fp32 arithmetic on the bf16 inputs
//   iou[g,k] = sigmoid(h_iou[k]·w2 + b2)                          h_iou = relu(W1·q+b1) [*,128]
//   best[g]  = first argmax_k bf16(sim[c0(g),k])  (torch.argmax tie-break on the bf16 values the reference
//              returns, training.py:627-629; c0 = first conversation of the group)
// sim / iou leave in fp32, UNROUNDED: the bf16 `pred_similarity` / `pred_iou` the caller hands out are one
// rounding of these.  (Round 1 mimicked the eager path's five bf16 rounding points here — norm, quotients,
// dot, logit, sigmoid; tests/parity_bisect.py showed that chain alone cost up to 4e-3 on pred_iou against
// the fp32 oracle with exact inputs, more than both encoders together.)
// conv_valid[c] < 0 (the conversation holds no [SEG], its text row is meaningless): the row is NaN and, for a
// first conversation, iou is NaN and best = -1 — a sentinel instead of plausible-looking numbers.
// outputs are padded to k_stride (-inf / 0 beyond K_g).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
select_kernel(const bf16* __restrict__ feat, const bf16* __restrict__ text,
              const bf16* __restrict__ h_iou, const bf16* __restrict__ w2, const bf16* __restrict__ b2,
              const int* __restrict__ k_off, const int* __restrict__ conv_group,
              const int* __restrict__ conv_valid, float* __restrict__ sim_out, float* __restrict__ iou_out,
              int* __restrict__ best, int k_stride) {
  const int c = blockIdx.x;
  const int g = conv_group ? conv_group[c] : c;
  const bool first = conv_group == nullptr || c == 0 || conv_group[c - 1] != g;
  const bool valid = conv_valid == nullptr || conv_valid[c] >= 0;
  const int r0 = k_off[g], nk = k_off[g + 1] - r0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float tn[256];
  __shared__ float sims[SA_MAX];
  if (warp == 0) {
    float t[8], ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t[i] = __bfloat162float(text[(size_t)c * 256 + lane * 8 + i]);
      ss += t[i] * t[i];
    }
    const float inv = 1.f / sqrtf(warp_sum(ss));
#pragma unroll
    for (int i = 0; i < 8; ++i) tn[lane * 8 + i] = t[i] * inv;
  }
  __syncthreads();
  const float qnan = __int_as_float(0x7fc00000);
  for (int kk = warp; kk < k_stride; kk += 8) {
    if (kk >= nk) {
      if (lane == 0) {
        sim_out[(size_t)c * k_stride + kk] = -INFINITY;
        if (first) iou_out[(size_t)g * k_stride + kk] = 0.f;
      }
      continue;
    }
    const size_t r = (size_t)(r0 + kk);
    float e[8], ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      e[i] = __bfloat162float(feat[r * 256 + lane * 8 + i]);
      ss += e[i] * e[i];
    }
    const float inv = 1.f / sqrtf(warp_sum(ss));
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dot += tn[lane * 8 + i] * e[i];
    dot = warp_sum(dot) * inv;
    float io = 0.f;
    if (first) {
      float hi = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        hi += __bfloat162float(h_iou[r * 128 + lane * 4 + i]) * __bfloat162float(w2[lane * 4 + i]);
      hi = warp_sum(hi) + __bfloat162float(b2[0]);
      io = 1.f / (1.f + expf(-hi));
    }
    if (lane == 0) {
      sim_out[(size_t)c * k_stride + kk] = valid ? dot : qnan;
      if (first) iou_out[(size_t)g * k_stride + kk] = valid ? io : qnan;
      if (kk < SA_MAX) sims[kk] = bf16_round(dot);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && first) {
    int bi = 0;
    float bv = -INFINITY;
    for (int i = 0; i < nk && i < SA_MAX; ++i)
      if (sims[i] > bv) { bv = sims[i]; bi = i; }
    best[g] = (nk > 0 && valid) ? bi : -1;
  }
}

// ---------------------------------------------------------------------------------------------
// losses (training-side; one block each, fp32 math on bf16/fp32 inputs)
// ---------------------------------------------------------------------------------------------
__device__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}
__device__ float block_max(float v, float* sh) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, sh[i]);
  return t;
}

// kl = KL(softmax(gt_iou/τ) ‖ softmax(sim/τ)) (sum over the K proposals), mse = mean((p-g)²·e^{g-1})·50 with
// g = gt_iop   (loss.py:50-94; the training forward feeds IoU to the first and IoP to the second,
// LISA.py:452-453).  Whole block cooperates; results valid in every thread.
__device__ void align_iou_pair(const float* __restrict__ sim, const float* __restrict__ pred_iou,
                               const float* __restrict__ gt_iou, const float* __restrict__ gt_iop, int K,
                               float it, float* sh, float& kl_out, float& mse_out) {
  float ms = -INFINITY, mg = -INFINITY;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    ms = fmaxf(ms, sim[i] * it);
    mg = fmaxf(mg, gt_iou[i] * it);
  }
  ms = block_max(ms, sh);
  mg = block_max(mg, sh);
  float zs = 0.f, zg = 0.f;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    zs += __expf(sim[i] * it - ms);
    zg += __expf(gt_iou[i] * it - mg);
  }
  zs = block_sum(zs, sh);
  zg = block_sum(zg, sh);
  const float ls = ms + __logf(zs), lg = mg + __logf(zg);
  float kl = 0.f, mse = 0.f;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const float lp_s = sim[i] * it - ls, lp_g = gt_iou[i] * it - lg;
    const float pg = __expf(lp_g);
    kl += pg > 0.f ? pg * (lp_g - lp_s) : 0.f;
    const float d = pred_iou[i] - gt_iop[i];
    mse += d * d * __expf(gt_iop[i] - 1.f);
  }
  kl_out = block_sum(kl, sh);
  mse_out = block_sum(mse, sh) / (float)K * 50.f;
}

__global__ void __launch_bounds__(256)
align_iou_loss_kernel(const float* __restrict__ sim, const float* __restrict__ pred_iou,
                      const float* __restrict__ gt_iou, int K, float temperature, float* __restrict__ out) {
  __shared__ float sh[8];
  float kl, mse;
  align_iou_pair(sim, pred_iou, gt_iou, gt_iou, K, 1.f / temperature, sh, kl, mse);
  if (threadIdx.x == 0) {
    out[0] = kl;
    out[1] = mse;
  }
}

// Training forward, LISA.py:416-474: one block per (image, round) group g; rows of stride k_stride, K_g =
// k_off[g+1]-k_off[g] valid proposals.  per_group[g] = {align_g, regression_g}.
__global__ void __launch_bounds__(256)
selector_losses_kernel(const float* __restrict__ sim, const float* __restrict__ pred_iou,
                       const float* __restrict__ gt_iou, const float* __restrict__ gt_iop,
                       const int32_t* __restrict__ k_off, int k_stride, float temperature,
                       float* __restrict__ per_group) {
  __shared__ float sh[8];
  const int g = blockIdx.x;
  const int K = k_off[g + 1] - k_off[g];
  const size_t o = (size_t)g * k_stride;
  float kl = 0.f, mse = 0.f;
  if (K > 0) align_iou_pair(sim + o, pred_iou + o, gt_iou + o, gt_iop + o, K, 1.f / temperature, sh, kl, mse);
  if (threadIdx.x == 0) {
    per_group[2 * g + 0] = kl;
    per_group[2 * g + 1] = mse;
  }
}
// out4 = {loss, ce_loss, align_loss, regression_loss}: group_weight[g] = 1 / ((rounds of g's image + 1e-8) *
// images with >= 1 round) reproduces the per-image mean over rounds and the mean over images (LISA.py:455-462);
// fixed summation order.
__global__ void selector_losses_final_kernel(const float* __restrict__ per_group, const float* __restrict__ group_weight,
                                             int n_groups, const float* __restrict__ ce2, float w_ce, float w_align,
                                             float w_reg, float* __restrict__ out4) {
  float a = 0.f, r = 0.f;
  for (int g = 0; g < n_groups; ++g) {
    a += per_group[2 * g + 0] * group_weight[g];
    r += per_group[2 * g + 1] * group_weight[g];
  }
  const float ce = (ce2 ? ce2[0] : 0.f) * w_ce;
  a *= w_align;
  r *= w_reg;
  out4[0] = ce + a + r;
  out4[1] = ce;
  out4[2] = a;
  out4[3] = r;
}

// ---------------------------------------------------------------------------------------------
// Language-model cross entropy of the training forward: llava_llama.py:107-118 (shift by one, mean over the
// targets != ignore_index) on the labels of prepare_inputs_labels_for_multimodal (llava_arch.py:185-245:
// IGNORE over the image rows), with the label splice done by index arithmetic instead of a copy.
// grid (T, n_seq): block (t, n) scores logits row n*T+t against the label of spliced position t+1.
// row_loss: CE >= 0, -1 = no target at this row, NaN = target outside [0, vocab).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lm_ce_rows_kernel(const bf16* __restrict__ logits, int ld, const long long* __restrict__ input_ids,
                  const long long* __restrict__ labels, int t_text, int n_img, int vocab, long long image_token,
                  long long ignore_index, float* __restrict__ row_loss) {
  __shared__ int s_img;
  __shared__ float sh[8];
  const int T = t_text + n_img - 1;
  const int t = blockIdx.x, n = blockIdx.y;
  const int pos = t + 1;
  const long long* ids = input_ids + (size_t)n * t_text;
  if (threadIdx.x == 0) s_img = t_text;
  __syncthreads();
  for (int j = threadIdx.x; j < t_text; j += blockDim.x)
    if (ids[j] == image_token) atomicMin(&s_img, j);
  __syncthreads();
  const int i_img = s_img;
  long long lab = ignore_index;
  if (pos < T && i_img < t_text) {
    if (pos < i_img) lab = labels[(size_t)n * t_text + pos];
    else if (pos >= i_img + n_img) lab = labels[(size_t)n * t_text + pos - n_img + 1];
  }
  const size_t row = (size_t)n * T + t;
  if (lab == ignore_index) {
    if (threadIdx.x == 0) row_loss[row] = -1.f;
    return;
  }
  if (lab < 0 || lab >= vocab) {
    if (threadIdx.x == 0) row_loss[row] = __int_as_float(0x7fc00000);
    return;
  }
  const bf16* x = logits + row * (size_t)ld;
  // online (max, sum exp) per thread over 16-byte vectors; ld % 8 == 0 keeps every row 16-byte aligned
  float m = -INFINITY, z = 0.f;
  const int nvec = vocab >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + v);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 p = __bfloat1622float2(h[j]);
      f[2 * j] = p.x;
      f[2 * j + 1] = p.y;
    }
    float vm = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) vm = fmaxf(vm, f[j]);
    if (vm > m) {
      z *= __expf(m - vm);
      m = vm;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) z += __expf(f[j] - m);
  }
  for (int c = (nvec << 3) + threadIdx.x; c < vocab; c += blockDim.x) {
    const float f = __bfloat162float(x[c]);
    if (f > m) {
      z *= __expf(m - f);
      m = f;
    }
    z += __expf(f - m);
  }
  const float bm = block_max(m, sh);
  z = block_sum(z * __expf(m - bm), sh);
  if (threadIdx.x == 0) row_loss[row] = bm + __logf(z) - __bfloat162float(x[lab]);
}
// out2 = {mean CE over the rows that have a target, number of such rows}; fixed summation order.
__global__ void __launch_bounds__(256) lm_ce_final_kernel(const float* __restrict__ row_loss, int rows,
                                                          float* __restrict__ out2) {
  __shared__ float sh[8];
  float s = 0.f, c = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float v = row_loss[r];
    if (!(v < 0.f)) {  // NaN (bad target) propagates on purpose
      s += v;
      c += 1.f;
    }
  }
  s = block_sum(s, sh);
  c = block_sum(c, sh);
  if (threadIdx.x == 0) {
    out2[0] = s / c;
    out2[1] = c;
  }
}

// per-mask partials for dice / BCE: grid (n_masks), out2[m] = {dice_m, bce_mean_m}   (loss.py:4-47)
__global__ void __launch_bounds__(256)
dice_bce_kernel(const float* __restrict__ logits, const float* __restrict__ targets, int hw, float scale,
                float eps, float* __restrict__ out2) {
  __shared__ float sh[8];
  const float* x = logits + (size_t)blockIdx.x * hw;
  const float* t = targets + (size_t)blockIdx.x * hw;
  float num = 0.f, dp = 0.f, dt = 0.f, bce = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float xv = x[i], tv = t[i];
    const float p = 1.f / (1.f + __expf(-xv));
    num += p / scale * tv;
    dp += p / scale;
    dt += tv / scale;
    bce += fmaxf(xv, 0.f) - xv * tv + log1pf(__expf(-fabsf(xv)));
  }
  num = block_sum(num, sh);
  dp = block_sum(dp, sh);
  dt = block_sum(dt, sh);
  bce = block_sum(bce, sh);
  if (threadIdx.x == 0) {
    out2[blockIdx.x * 2 + 0] = 1.f - (2.f * num + eps) / (dp + dt + eps);
    out2[blockIdx.x * 2 + 1] = bce / (float)hw;
  }
}
__global__ void dice_bce_final_kernel(const float* __restrict__ part, int n, float num_masks,
                                      float* __restrict__ out) {
  float d = 0.f, b = 0.f;
  for (int i = 0; i < n; ++i) {
    d += part[2 * i];
    b += part[2 * i + 1];
  }
  out[0] = d / (num_masks + 1e-8f);
  out[1] = b / (num_masks + 1e-8f);
}

}  // namespace
}  // namespace llmseg

using namespace llmseg;

extern "C" size_t llmseg_maskpool_workspace(int n_masks) {
  return (size_t)n_masks * (4096 + 8 + 8 * 256) * sizeof(float);
}

extern "C" int llmseg_maskpool(const void* segs, const void* emb, const int32_t* mask_image,
                               int n_masks, void* out, void* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(segs && emb && mask_image && out && workspace, LLMSEG_EARG, "llmseg_maskpool: null pointer");
  LLMSEG_REQUIRE(n_masks > 0, LLMSEG_ESHAPE, "llmseg_maskpool: n_masks=%d", n_masks);
  float* wt = static_cast<float*>(workspace);
  float* ps = wt + (size_t)n_masks * 4096;
  maskpool_adjoint_kernel<<<dim3(8, n_masks), 256, 0, stream>>>(static_cast<const bf16*>(segs), wt, ps);
  LLMSEG_CUDA(cudaGetLastError());
  float* partial = ps + (size_t)n_masks * 8;
  maskpool_apply_kernel<<<dim3((n_masks + 7) / 8, 8), 256, 0, stream>>>(wt, static_cast<const bf16*>(emb),
                                                                        mask_image, partial, n_masks);
  LLMSEG_CUDA(cudaGetLastError());
  maskpool_final_kernel<<<n_masks, 256, 0, stream>>>(partial, ps, static_cast<bf16*>(out), n_masks);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(3);
  return 0;
}

extern "C" int llmseg_small_attention(const void* q, int ldq, const void* k, int ldk, const void* v,
                                      int ldv, void* out, int ldo, const int32_t* q_off,
                                      const int32_t* kv_off, int batch, int heads, int head_dim,
                                      int max_kv, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(q && k && v && out && q_off && kv_off, LLMSEG_EARG, "llmseg_small_attention: null pointer");
  LLMSEG_REQUIRE(head_dim == SA_HD && batch > 0 && heads > 0 && max_kv <= SA_MAX, LLMSEG_ESHAPE,
                 "llmseg_small_attention: head_dim=%d (need 32) max_kv=%d (<= %d)", head_dim, max_kv, SA_MAX);
  small_attn_kernel<<<dim3(batch, heads), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, static_cast<const bf16*>(v),
      ldv, static_cast<bf16*>(out), ldo, q_off, kv_off, 1.0f / sqrtf((float)head_dim));
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_select(const void* feat, const void* text, const void* h_iou, const void* w2,
                             const void* b2, const int32_t* k_off, int n_groups, int k_stride,
                             const int32_t* conv_group, const int32_t* conv_valid, int n_conv,
                             float* sim_out, float* iou_out, int32_t* best, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(feat && text && h_iou && w2 && b2 && k_off && sim_out && iou_out && best, LLMSEG_EARG,
                 "llmseg_select: null pointer");
  LLMSEG_REQUIRE(n_groups > 0 && k_stride > 0 && k_stride <= SA_MAX, LLMSEG_ESHAPE,
                 "llmseg_select: n_groups=%d k_stride=%d (<= %d)", n_groups, k_stride, SA_MAX);
  LLMSEG_REQUIRE(conv_group ? n_conv >= n_groups : n_conv == n_groups, LLMSEG_ESHAPE,
                 "llmseg_select: n_conv=%d for %d groups (%s conv_group)", n_conv, n_groups,
                 conv_group ? "with" : "without");
  select_kernel<<<n_conv, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(feat), static_cast<const bf16*>(text), static_cast<const bf16*>(h_iou),
      static_cast<const bf16*>(w2), static_cast<const bf16*>(b2), k_off, conv_group, conv_valid, sim_out,
      iou_out, best, k_stride);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_align_iou_loss(const float* sim, const float* pred_iou, const float* gt_iou, int K,
                                     float temperature, float* out2, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(sim && pred_iou && gt_iou && out2 && K > 0 && temperature > 0.f, LLMSEG_EARG,
                 "llmseg_align_iou_loss: bad arguments");
  align_iou_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(sim, pred_iou, gt_iou, K,
                                                                          temperature, out2);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_selector_losses(const float* sim, const float* pred_iou, const float* gt_iou,
                                      const float* gt_iop, const int32_t* k_off, int n_groups, int k_stride,
                                      float temperature, const float* group_weight, const float* ce2, float w_ce,
                                      float w_align, float w_reg, float* per_group, float* out4, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(sim && pred_iou && gt_iou && gt_iop && k_off && group_weight && per_group && out4, LLMSEG_EARG,
                 "llmseg_selector_losses: null pointer");
  LLMSEG_REQUIRE(n_groups > 0 && k_stride > 0 && temperature > 0.f, LLMSEG_ESHAPE,
                 "llmseg_selector_losses: groups=%d k_stride=%d temperature=%f", n_groups, k_stride, temperature);
  selector_losses_kernel<<<n_groups, 256, 0, stream>>>(sim, pred_iou, gt_iou, gt_iop, k_off, k_stride, temperature,
                                                       per_group);
  LLMSEG_CUDA(cudaGetLastError());
  selector_losses_final_kernel<<<1, 1, 0, stream>>>(per_group, group_weight, n_groups, ce2, w_ce, w_align, w_reg,
                                                    out4);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  return 0;
}

extern "C" int llmseg_lm_cross_entropy(const void* logits, int ld, const int64_t* input_ids, const int64_t* labels,
                                       int n_seq, int t_text, int n_img_tokens, int vocab, int64_t image_token_id,
                                       int64_t ignore_index, float* row_loss, float* out2, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(logits && input_ids && labels && row_loss && out2, LLMSEG_EARG, "llmseg_lm_cross_entropy: null pointer");
  LLMSEG_REQUIRE(n_seq > 0 && n_seq <= 65535 && t_text > 0 && n_img_tokens > 0 && vocab > 0 && ld >= vocab,
                 LLMSEG_ESHAPE, "llmseg_lm_cross_entropy: n=%d t=%d F=%d vocab=%d ld=%d", n_seq, t_text, n_img_tokens,
                 vocab, ld);
  LLMSEG_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0, LLMSEG_EALIGN,
                 "llmseg_lm_cross_entropy: logits / ld not 16-byte aligned");
  const int T = t_text + n_img_tokens - 1;
  lm_ce_rows_kernel<<<dim3(T, n_seq), 256, 0, stream>>>(
      static_cast<const bf16*>(logits), ld, reinterpret_cast<const long long*>(input_ids),
      reinterpret_cast<const long long*>(labels), t_text, n_img_tokens, vocab, image_token_id, ignore_index, row_loss);
  LLMSEG_CUDA(cudaGetLastError());
  lm_ce_final_kernel<<<1, 256, 0, stream>>>(row_loss, n_seq * T, out2);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  return 0;
}

extern "C" int llmseg_dice_bce_loss(const float* logits, const float* targets, int n_masks, int hw,
                                    float num_masks, float* workspace, float* out2, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(logits && targets && workspace && out2 && n_masks > 0 && hw > 0, LLMSEG_EARG,
                 "llmseg_dice_bce_loss: bad arguments");
  dice_bce_kernel<<<n_masks, 256, 0, stream>>>(logits, targets, hw, 1000.f, 1e-6f, workspace);
  LLMSEG_CUDA(cudaGetLastError());
  dice_bce_final_kernel<<<1, 1, 0, stream>>>(workspace, n_masks, num_masks, out2);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  return 0;
}
