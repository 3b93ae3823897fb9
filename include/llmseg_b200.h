/* llmseg_b200 — C ABI of the B200-native LLM-Seg forward path.
 *
 * The reference (wangjunchi/LLMSeg) has no FFI / plugin layer: its hot path is pure PyTorch
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds instead of the ATen ops that
 * `LISAForCausalLM.model_forward` (reference model/LISA.py:225-414) issues; each entry point cites
 * the reference call site it replaces.  The binding used by this repo is ctypes
 * (llmseg_b200/_lib.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 or a negative LLMSEG_E* code; text via llmseg_last_error()
 *   - all pointers are DEVICE pointers unless the name says host; bf16 = 2-byte bfloat16
 *   - nothing allocates or frees caller memory; work is enqueued on `stream` (a cudaStream_t
 *     passed as void*) and is asynchronous
 *   - thread-safe for distinct streams.  Process-wide state is limited to: the launch counter behind
 *     llmseg_launch_count() (atomic), the thread-local text behind llmseg_last_error(), one-time function
 *     attributes (dynamic shared-memory opt-in per kernel) and the LLMSEG_* environment switches, which are read
 *     once or per call (DESIGN.md lists them) — nothing a concurrent caller on another stream can observe changing
 *   - sm_100 only: any other device returns LLMSEG_EARCH (there is no CPU / other-arch fallback)
 */
#ifndef LLMSEG_B200_H_
#define LLMSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLMSEG_OK 0
#define LLMSEG_ESHAPE (-1)  /* unsupported / inconsistent shape */
#define LLMSEG_EALIGN (-2)  /* pointer or stride not 16-byte aligned */
#define LLMSEG_EARCH (-3)   /* device is not sm_100 */
#define LLMSEG_ECUDA (-4)   /* CUDA runtime / driver error */
#define LLMSEG_EARG (-5)    /* null pointer or bad enum */

/* activation applied in the GEMM epilogue (after bias, before residual) */
#define LLMSEG_ACT_NONE 0
#define LLMSEG_ACT_GELU 1       /* erf GELU   — SAM MLPBlock, reference common.py:13-26          */
#define LLMSEG_ACT_QUICK_GELU 2 /* x*sigmoid(1.702x) — CLIP ViT-L/14 MLP (transformers CLIPMLP)  */
#define LLMSEG_ACT_RELU 3       /* selector MLPs, reference model/transformer.py:13-26           */

#define LLMSEG_GEMM_PLAIN 0
#define LLMSEG_GEMM_SWIGLU 1 /* W rows interleaved (gate0,up0,gate1,up1,..); C[:, j] = silu(g_j)*u_j */
#define LLMSEG_GEMM_QKV 2    /* split columns into per-head Q, K and transposed V buffers          */

const char* llmseg_last_error(void);
int llmseg_version(void);
/* number of kernel launches issued through this library by the calling process (for bench.py) */
uint64_t llmseg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * C = epilogue(A · Wᵀ)         tcgen05 / TMEM / TMA GEMM, bf16 in, fp32 accumulate, bf16 out.
 * Replaces every nn.Linear / 1×1-conv / patch-embed conv on the path:
 *   SAM  qkv / proj / MLP            reference image_encoder.py:223-224,238-258; common.py:21-26
 *   SAM  patch-embed + pos, neck     image_encoder.py:111-113,418-426,92-108
 *   CLIP / LLaMA linears             clip_encoder.py:53-57, llava_llama.py:93-102 (transformers)
 *   mm_projector, text_hidden_fcs    llava_arch.py:35,95; LISA.py:56-65,317-318
 * The epilogue (bias, activation, residual, RoPE, SwiGLU) runs in fp32 on the fp32 accumulators and rounds
 * to bf16 once, at the store — at least as close to the fp32 reference as its bf16 autocast rounding chain.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int M, N, K;          /* K % 8 == 0 (pad with zeros otherwise)                              */
  const void* A;        /* bf16 [M, lda]                                                       */
  int lda;
  const void* W;        /* bf16 [N, ldw]  (nn.Linear weight layout)                            */
  int ldw;
  void* C;              /* bf16 [*, ldc]  (PLAIN / SWIGLU)                                     */
  int ldc;
  const void* bias;     /* bf16 [N] or NULL                                                    */
  const void* residual; /* bf16 [*, ldr] or NULL; row = out_row % res_mod (res_mod 0: no mod)  */
  int ldr;
  int res_mod;
  int act;
  int mode;
  const int32_t* out_row_map; /* [M] or NULL: C/residual row of GEMM row r; negative = dropped */
  /* LLMSEG_GEMM_QKV: N = 3*heads*head_dim, columns ordered (which, head, d).  GEMM row r is
   * token s = r % seq_in of sequence b = r / seq_in.
   * With out_row_map given, GEMM row r is instead token s = m % seq_in of sequence b = m / seq_in for
   * m = out_row_map[r] (negative = dropped): the SAM window partition (image_encoder.py:263-288)
   * folded into the projection, so the zero padding tokens are never multiplied.
   *   q, k : bf16 [(b*heads+h), seq_pad, head_dim]
   *   vt   : bf16 [(b*heads+h), head_dim, seq_pad]          (transposed for the PV tensor-core tile)
   * rope_cos/sin: bf16 [>=seq_in, head_dim/2] rotate-half RoPE applied to q and k (LLaMA) or NULL */
  void* q;
  void* k;
  void* vt;
  int heads, head_dim, seq_in, seq_pad;
  const void* rope_cos;
  const void* rope_sin;
  /* Optional scratch (NULL / 0 = none): llmseg_gemm_workspace_bytes() bytes of device memory, 256-byte
   * aligned, ZERO-FILLED once by the caller before its first use (the library leaves its flag area zeroed
   * after every call) and never used by two GEMMs that may run concurrently (one buffer per stream).  With
   * it (a) row statistics can be finished in-kernel (stats_final) and (b) problems whose last wave of output tiles would leave most SMs idle split that wave along K across
   * all SMs (stream-K tail: fp32 partial tiles + flags live in the workspace). */
  void* workspace;
  size_t workspace_bytes;
  /* LayerNorm / RMSNorm of the A rows folded into the GEMM (NULL = off).  With Norm(x) = (x - mean)·rstd ⊙
   * gamma + beta and y = Norm(x)·Wᵀ + b:   y = rstd · (x · W''ᵀ) + b'   where
   *   W''[n,:] = W[n,:] ⊙ gamma - mean_k(W[n,k]·gamma[k])   (rows centred: x - mean is orthogonal to 1, so
   *                                                          the mean term drops out; RMSNorm: not centred)
   *   b'[n]    = b[n] + Σ_k beta[k]·W[n,k]
   * Pass A = x (un-normalised), W = bf16(W''), bias = bf16(b') and
   *   row_stats : float2 [M] (mean, rstd) per A row, from llmseg_norm_stats — only rstd is used here.
   * The epilogue multiplies the fp32 accumulators by rstd before bias / activation / RoPE / SwiGLU / the
   * Q-K-V split.  Replaces the separate norm pass over x (reference image_encoder.py:179,191;
   * transformers CLIPEncoderLayer / LlamaDecoderLayer norms) by a statistics-only read. */
  const void* row_stats;
  /* row_stats_parts > 0: row_stats is instead [M][row_stats_parts] float2 partial (sum, sum of squares) of
   * each A row — what a previous llmseg_gemm wrote through stats_out — and the epilogue derives
   * mean = Σsum/norm_dim, rstd = rsqrt(Σsq/norm_dim - mean² + norm_eps)  (norm_rms: rsqrt(Σsq/norm_dim + eps)). */
  int row_stats_parts;
  int norm_dim;
  float norm_eps;
  int norm_rms;
  /* LLMSEG_GEMM_PLAIN producer side: float2 [out rows][llmseg_gemm_stats_parts(M, N)] — every epilogue warp
   * writes the (sum, sum of squares) of its share of each output row (fp32 values before the bf16 store), so
   * the norm that follows needs no pass over C at all.  NULL = off. */
  void* stats_out;
  /* With stats_final (float2 [M], needs stats_out, the workspace and no out_row_map) the partials are also
   * finished inside the kernel: the last tile to complete a 128-row block reduces that block's partials (in
   * index order — deterministic) into (mean, rstd) with stats_dim / stats_eps / stats_rms, ready to be passed
   * as the next GEMM's row_stats with row_stats_parts = 0. */
  void* stats_final;
  int stats_dim;
  float stats_eps;
  int stats_rms;
} llmseg_gemm_params;
int llmseg_gemm(const llmseg_gemm_params* p, void* stream);
size_t llmseg_gemm_workspace_bytes(void);
int llmseg_gemm_stats_parts(int M, int N);

/* ------------------------------------------------------------------------------------------
 * Fused attention  softmax(scale·QKᵀ + bias + mask)·V   (flash-style, tcgen05/TMEM, TMA-fed).
 *   SAM windowed / global with decomposed rel-pos   image_encoder.py:244-257,354-392
 *   CLIP ViT-L/14 self-attention                      (transformers CLIPAttention, eager)
 *   LLaMA causal self-attention (+ right padding)     (transformers LlamaAttention, eager)
 * q,k,vt are the buffers written by LLMSEG_GEMM_QKV.  out is token-major bf16
 * [batch*seq, heads*head_dim] (ldo elements per row) ready for the output projection.
 *
 * Decomposed rel-pos bias rides on the tensor core: the score tile is computed over an extended
 * reduction dimension  S = [q | qext]·[k | kext]ᵀ  where qext (written by llmseg_relpos_prep) holds
 * the per-query bias values divided by `scale` and kext is a constant one-hot key-position matrix:
 *   ext_cols = 32 (14×14 windows): qext[:, kh] = q·rel_h[qh-kh+13], qext[:, 14+kw] = q·rel_w[qw-kw+13];
 *                                  kext bf16 [256, 32], kext[key, key/14] = kext[key, 14+key%14] = 1
 *   ext_cols = 64 (64×64 global) : qext[:, kw] = q·rel_w[qw-kw+63];  kext bf16 [128, 64] = [I;I];
 *                                  the rel_h term is a per-(query, key-row) constant read from
 *                                  row_bias bf16 [(b*heads+h), seq_pad, 64] (index key/64)
 * kv_len: int32 [batch] number of valid (non right-padded) keys, or NULL for all.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q;
  const void* k;
  const void* vt;
  void* out;
  int ldo;
  int batch, heads, head_dim, seq, seq_pad;
  float scale;
  int causal;
  const int32_t* kv_len;
  int ext_cols;         /* 0, 32 or 64 */
  const void* qext;     /* bf16 [(b*heads+h), seq_pad, ext_cols] or NULL */
  const void* kext;     /* bf16 constant, see above, or NULL            */
  const void* row_bias; /* bf16 [(b*heads+h), seq_pad, 64] (ext_cols == 64) or NULL */
  const int32_t* out_row_map; /* int32 [batch*seq] or NULL: output row of query (b, s); negative = not
                                 written (SAM window un-partition + crop, image_encoder.py:291-318) */
} llmseg_attn_params;
int llmseg_attention(const llmseg_attn_params* p, void* stream);

/* Rel-pos prologue: QR = q · [rel_h; rel_w]ᵀ on the tensor core (same GEMM kernel, gather
 * epilogue), scattered into qext (and row_bias for the global case) as described above.
 * Replaces get_rel_pos + the two einsums of add_decomposed_rel_pos (image_encoder.py:321-392);
 * q is the UNSCALED query, values are rounded to bf16 like the reference's einsum outputs.
 *   q      bf16 [bh, seq_pad, head_dim]
 *   rel_hw bf16 [n_pad, head_dim], zero padded:
 *            grid 14 (windows): n_pad 64,  rows [0,27)  = rel_pos_h, rows [32,59)   = rel_pos_w
 *            grid 64 (global) : n_pad 256, rows [0,127) = rel_pos_h, rows [128,255) = rel_pos_w
 *   seq == grid*grid; grid == 14 -> ext_cols 32, row_bias NULL; grid == 64 -> ext_cols 64 + row_bias */
int llmseg_relpos_prep(const void* q, const void* rel_hw, int n_pad, int bh, int seq, int seq_pad,
                       int head_dim, int grid, float inv_scale, void* qext, int ext_cols,
                       void* row_bias, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row norms (HBM-bound, one warp per row, 16-byte vector loads, warp-shuffle reductions).
 * out[r] = norm(in[src_row_map ? src_row_map[r] : r]); a negative source row writes zeros
 * (the SAM 64→70 window padding is zeros *after* LayerNorm, image_encoder.py:179-185,281).
 *   layernorm : SAM norm1/norm2 (eps 1e-6), LayerNorm2d on NHWC rows (common.py:31-43),
 *               CLIP LayerNorms (eps 1e-5), selector LayerNorms (transformer.py:240-250)
 *   rmsnorm   : LLaMA RMSNorm, fp32 variance, eps 1e-6 (transformers LlamaRMSNorm)
 * ------------------------------------------------------------------------------------------ */
int llmseg_layernorm(const void* in, int ld_in, void* out, int ld_out, const void* gamma,
                     const void* beta, int rows_out, int dim, float eps,
                     const int32_t* src_row_map, void* stream);
/* Row statistics only: stats[r] = (mean, rstd) of row r (rms != 0: (0, rsqrt(mean(x^2) + eps))), fp32
 * float2 [rows]; same arithmetic as llmseg_layernorm / llmseg_rmsnorm.  Feeds llmseg_gemm's row_stats. */
int llmseg_norm_stats(const void* in, int ld_in, int rows, int dim, float eps, int rms, void* stats,
                      void* stream);
int llmseg_rmsnorm(const void* in, int ld_in, void* out, int ld_out, const void* gamma,
                   int rows, int dim, float eps, const int32_t* src_row_map, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data movement around the GEMMs (HBM-bound).
 *   patchify      NCHW bf16 image → [batch*(g*g+cls_rows), k_pad] patch rows, column order
 *                 (c,py,px) = flattened conv weight, so PatchEmbed / CLIP patch_embedding become
 *                 one GEMM (image_encoder.py:418-426; transformers CLIPVisionEmbeddings).  With
 *                 cls_rows=1 each image gets a leading row that is zero except column 3*p*p = 1,
 *                 which selects a class-embedding column appended to the weight.
 *   embed_splice  LLaVA `prepare_inputs_labels_for_multimodal` for the inference layout
 *                 (llava_arch.py:185-245,332-345): token embeddings with the single IMAGE token
 *                 (-200) replaced by n_img_tokens feature rows; also emits
 *                 kv_len[n] = n_img_tokens-1 + #true(attention_mask[n]) and the flat row index of
 *                 the hidden state that predicts [SEG] (LISA.py:254-266), -1 if absent.
 *   add_rows_bcast out[r] = x[r] + y[row_group ? row_group[r] : r/group]   (single-key cross
 *                 attention collapses to a broadcast add, transformer.py:264-269)
 * ------------------------------------------------------------------------------------------ */
int llmseg_patchify(const void* images, void* out, int batch, int img_size, int patch, int k_pad,
                    int cls_rows, void* stream);
int llmseg_embed_splice(const int64_t* input_ids, const uint8_t* attention_mask,
                        const void* embed_table, const void* image_feats, void* out,
                        int32_t* kv_len, int32_t* seg_row, int n_seq, int t_text, int n_img_tokens,
                        int dim, int64_t image_token_id, int64_t seg_token_id, int vocab,
                        void* stream);
int llmseg_add_rows_bcast(const void* x, const void* y, void* out, int rows, int dim, int group,
                          const int32_t* row_group, void* stream);
/* out[r] = in[src_row_map[r]] (bf16 rows of `dim`, zero row for a negative index): gathers the [SEG] rows
 * (LISA.py:322-337) so the tail of the last LLaMA layer runs on those rows only. */
int llmseg_gather_rows(const void* in, int ld_in, void* out, int rows, int dim, const int32_t* src_row_map,
                       void* stream);
/* K and Vᵀ entries of window padding tokens: the reference pads with zeros AFTER LayerNorm, so their
 * k, v equal the projection bias (image_encoder.py:179-185,238-242).  pos_map: int32 [*, seq_in],
 * negative = padding position (the same map llmseg_attention takes as out_row_map); seq_ids: int32
 * [n_seqs] sequences (windows) to visit — those that contain padding — or NULL for sequences
 * 0..n_seqs-1; bias_qkv: bf16 [3*heads*head_dim] (q | k | v). */
int llmseg_fill_kv_rows(void* k, void* vt, const void* bias_qkv, const int32_t* pos_map,
                        const int32_t* seq_ids, int n_seqs, int heads, int head_dim, int seq_in,
                        int seq_pad, void* stream);
/* 3x3 / pad-1 im2col on token-major NHWC bf16: out[(b,y,x), (ky,kx,c)]; turns the SAM neck
 * conv3x3 (image_encoder.py:100-106) into one GEMM with the weight permuted to [out,(ky,kx,c)]. */
int llmseg_im2col3x3(const void* in, void* out, int batch, int height, int width, int channels,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * Mask-proposal selector (reference LISA.py:350-408, model/transformer.py:215-341).
 *   maskpool        fused bilinear upsample 64²→256² + mask pooling in the exact adjoint form
 *                   (Uᵀw)·E: segs bf16 [n_masks,256,256] are read once; emb is the token-major
 *                   SAM neck output bf16 [B,4096,256]; mask_image[m] = image of mask m;
 *                   out bf16 [n_masks,256]; workspace = llmseg_maskpool_workspace(n_masks) bytes.
 *   small_attention 32-dim heads among <=128 mask tokens (self-attention) or from the single
 *                   text token to the masks; q rows q_off[b]..q_off[b+1], kv rows kv_off[b]..
 *   select          per conversation c of group g = conv_group[c] (NULL: identity, n_conv == n_groups; a
 *                   group's conversations are contiguous): sim[c,k] = cos(text_c, feat_k) over the group's mask
 *                   tokens (LISA.py:397-403, [C,K] per image); per group: iou[g,k] = sigmoid(h_iou_k·w2+b2),
 *                   best[g] = first argmax_k of the group's FIRST conversation; conv_valid[c] < 0 (no [SEG] in
 *                   that conversation; NULL: all valid) -> NaN row (and NaN iou / best -1 for a first
 *                   conversation).  sim fp32 [n_conv,k_stride], iou fp32 [n_groups,k_stride] (-inf / 0
 *                   padding), best int32 [n_groups]
 *   losses          align/IoU-regression (loss.py:50-94) → out2 = {kl, mse*50};
 *                   dice/BCE (loss.py:4-47) → out2 = {dice, bce}; workspace = 2*n_masks floats
 * ------------------------------------------------------------------------------------------ */
size_t llmseg_maskpool_workspace(int n_masks);
int llmseg_maskpool(const void* segs, const void* emb, const int32_t* mask_image, int n_masks,
                    void* out, void* workspace, void* stream);
int llmseg_small_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                           void* out, int ldo, const int32_t* q_off, const int32_t* kv_off,
                           int batch, int heads, int head_dim, int max_kv, void* stream);
int llmseg_select(const void* feat, const void* text, const void* h_iou, const void* w2,
                  const void* b2, const int32_t* k_off, int n_groups, int k_stride,
                  const int32_t* conv_group, const int32_t* conv_valid, int n_conv, float* sim_out,
                  float* iou_out, int32_t* best, void* stream);
int llmseg_align_iou_loss(const float* sim, const float* pred_iou, const float* gt_iou, int K,
                          float temperature, float* out2, void* stream);
 /* Training forward (LISA.py:416-474): per (image, round) group g, proposals k_off[g]..k_off[g+1] of rows with
 * stride k_stride: align_g = softmax_align_loss(sim_g, gt_iou_g), regression_g = iou_regression_loss(pred_iou_g,
 * gt_iop_g) (loss.py:50-94) -> per_group fp32 [n_groups,2]; then out4 = {loss, ce, align, regression} =
 * {sum, w_ce*ce2[0], w_align*Σ_g weight_g*align_g, w_reg*Σ_g weight_g*regression_g} (ce2 may be NULL). */
int llmseg_selector_losses(const float* sim, const float* pred_iou, const float* gt_iou,
                           const float* gt_iop, const int32_t* k_off, int n_groups, int k_stride,
                           float temperature, const float* group_weight, const float* ce2, float w_ce,
                           float w_align, float w_reg, float* per_group, float* out4, void* stream);
/* LM cross entropy of the training forward (llava_llama.py:107-118): logits bf16 [n_seq*T, ld] (T = t_text +
 * n_img_tokens - 1, first `vocab` columns valid), row (n,t) scored against the label of spliced position t+1,
 * where the spliced labels are labels[n] with the IMAGE token replaced by n_img_tokens IGNORE entries
 * (llava_arch.py:185-245) — computed by index arithmetic from input_ids/labels int64 [n_seq,t_text].
 * row_loss fp32 [n_seq*T] (scratch: CE, -1 = no target); out2 = {mean over targets, number of targets}. */
int llmseg_lm_cross_entropy(const void* logits, int ld, const int64_t* input_ids, const int64_t* labels,
                            int n_seq, int t_text, int n_img_tokens, int vocab, int64_t image_token_id,
                            int64_t ignore_index, float* row_loss, float* out2, void* stream);
int llmseg_dice_bce_loss(const float* logits, const float* targets, int n_masks, int hw,
                         float num_masks, float* workspace, float* out2, void* stream);

/* ------------------------------------------------------------------------------------------
 * SAM-Everything proposal generation: what surrounds llmseg_gemm when SAM's prompt encoder + mask
 * decoder + automatic mask generator run on the image features of llmseg's own SAM encoder, and LLM-Seg's resize of
 * the resulting masks to soft 256 x 256 proposals.  Replaces (reference model/segment_anything/...):
 *   point_tokens        modeling/prompt_encoder.py:73-90,186-195 + the token concatenation of mask_decoder.py:123-131
 *                       points fp32 [P,2] (x,y) input-frame pixels, gauss fp32 [2,128], out_tokens bf16 [5,256]
 *                       (iou_token ; mask_tokens), point_embed = point_embeddings[1], -> tokens bf16 [P,7,256]
 *   tok2img_attention   modeling/transformer.py:222-242 for 7 token queries x 4096 image keys, 8 heads x 16:
 *                       q bf16 [P*7, ldq] (128 cols used), k / v bf16 rows of prompt p at k + p*k_batch_stride
 *                       (0: every prompt shares one image's keys), -> out bf16 [P*7, ldo] (128 cols)
 *   img2tok_attention   same Attention for 4096 image queries x 7 token keys: q rows of prompt p at
 *                       q + p*q_batch_stride (0: shared), k / v bf16 [P*7, ld], -> out bf16 [P*4096, ldo]
 *   ln64_gelu           LayerNorm2d(64) + GELU (mask_decoder.py:56-62, common.py:31-43) on rows of 64 bf16
 *   mask_logits         mask_decoder.py:143-157 on the un-shuffled 2x2 ConvTranspose outputs: up2 bf16 [P*16384,128]
 *                       (row = prompt, token y, token x, dy, dx; col = dy2, dx2, channel), hyper bf16 [P,4,32]
 *                       -> low_res fp32 [P,3,256,256] (mask tokens 1..3: multimask output, mask_decoder.py:101-104)
 *   upscale_logits      the three steps above in one kernel (mask_decoder.py:56-64,139-157): up1 bf16 [P*4096, 256] =
 *                       ConvTranspose #1 as a GEMM with its bias (row = prompt, token; col = dy, dx, 64 channels), gamma /
 *                       beta bf16 [64] of the LayerNorm2d, w2 bf16 [128,64] (row = dy2, dx2, channel) and b2 bf16 [128] of
 *                       ConvTranspose #2, hyper bf16 [P,4,32] -> low_res fp32 [P,3,256,256].  ln64_gelu + llmseg_gemm +
 *                       mask_logits stay exported: they are the un-fused path the tests compare against
 *   mask_stats          modeling/sam.py:155-166 (4x bilinear up-sampling, evaluated on the fly) + utils/amg.py:156-176,
 *                       303-346: stats int32 [n,8] = {area, #(> t+o), #(> t-o), 1023-x0, 1023-y0, x1, y1, 0} of
 *                       candidate cand[i] (NULL: i) — stability = [1]/[2], box valid when area > 0
 *   box_nms             torchvision nms as automatic_mask_generator.py:256-262 calls it: boxes fp32 [n,4] XYXY sorted
 *                       by score (descending, stable), keep[i] = 0 if an earlier kept box has IoU > threshold
 *   mask_soft           utils/dataset.py:620-622 (antialiased bilinear 1024 -> 256) of the binarised up-sampled mask:
 *                       -> out bf16 [n,256,256]
 *   mask_binarize       the binary masks themselves: -> out uint8 [n,1024,1024]
 * ------------------------------------------------------------------------------------------ */
int llmseg_point_tokens(const float* points, int n_prompts, const float* gauss, const void* out_tokens,
                        const void* point_embed, const void* not_a_point, float img_size, void* tokens, void* stream);
int llmseg_tok2img_attention(const void* q, int ldq, const void* k, int ldk, long long k_batch_stride, const void* v,
                             int ldv, long long v_batch_stride, void* out, int ldo, int n_prompts, void* stream);
int llmseg_img2tok_attention(const void* q, int ldq, long long q_batch_stride, const void* k, int ldk, const void* v,
                             int ldv, void* out, int ldo, int n_prompts, void* stream);
int llmseg_ln64_gelu(const void* in, void* out, const void* gamma, const void* beta, long long rows, float eps,
                     void* stream);
int llmseg_mask_logits(const void* up2, const void* hyper, int n_prompts, float* low_res, void* stream);
int llmseg_upscale_logits(const void* up1, const void* gamma, const void* beta, float eps, const void* w2, const void* b2,
                          const void* hyper, int n_prompts, float* low_res, void* stream);
int llmseg_mask_stats(const float* low_res, const int32_t* cand, int n_cand, float threshold, float offset,
                      int32_t* stats, void* stream);
int llmseg_box_nms(const float* boxes_sorted, int n, float iou_threshold, int32_t* keep, void* stream);
int llmseg_mask_soft(const float* low_res, const int32_t* cand, int n_cand, float threshold, void* out, void* stream);
int llmseg_mask_binarize(const float* low_res, const int32_t* cand, int n_cand, float threshold, uint8_t* out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LLMSEG_B200_H_ */
