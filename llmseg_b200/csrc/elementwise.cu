// llmseg_b200 — data-movement kernels around the GEMMs (all HBM-bound, 16-byte vectors where the
// layout allows): patch extraction for the conv-as-GEMM patch embeddings, the LLaVA token/image
// splice, and the row-broadcast add used by the selector's single-key cross attention.
#include <atomic>

#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
namespace {

// One thread per (patch, channel, patch-row): copies `p` contiguous pixels.
// out row layout: [cls_rows zero/one-hot rows][g*g patches] per image, K = 3*p*p (+ zero pad to k_pad);
// column order (c, py, px) == flattened conv weight [out, c, py, px].
__global__ void patchify_kernel(const bf16* __restrict__ img, bf16* __restrict__ out, int B, int S,
                                int p, int g, int k_pad, int cls_rows) {
  const int rows_per_img = g * g + cls_rows;
  const int items_per_row = 3 * p + 1;  // last item: padding + CLS marker columns
  const long long total = (long long)B * rows_per_img * items_per_row;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int item = (int)(idx % items_per_row);
    const long long row = idx / items_per_row;
    const int b = (int)(row / rows_per_img);
    const int t = (int)(row % rows_per_img);
    bf16* orow = out + row * k_pad;
    const int K = 3 * p * p;
    if (item == 3 * p) {
      // pad columns [K, k_pad): zero, except column K of a CLS row = 1 (selects the class-embedding
      // column appended to the weight, so the GEMM emits class_embedding exactly)
      for (int c = K; c < k_pad; ++c)
        orow[c] = __float2bfloat16_rn((t < cls_rows && c == K) ? 1.0f : 0.0f);
      continue;
    }
    const int c = item / p, py = item % p;
    bf16* dst = orow + (c * p + py) * p;
    if (t < cls_rows) {
      for (int x = 0; x < p; ++x) dst[x] = __float2bfloat16_rn(0.f);
      continue;
    }
    const int pi = t - cls_rows;
    const int gy = pi / g, gx = pi % g;
    const bf16* src = img + (((size_t)b * 3 + c) * S + (gy * p + py)) * S + gx * p;
    if (p == 16) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      d4[0] = __ldg(s4);
      d4[1] = __ldg(s4 + 1);
    } else {
      for (int x = 0; x < p; ++x) dst[x] = src[x];
    }
  }
}

// One block per output row of the spliced sequence (reference llava_arch.py:185-245, inference
// layout: exactly one IMAGE token per row).  Row t of sequence n comes from
//   t <  i_img            : embed[ids[t]]
//   i_img <= t < i_img+F  : feats[n, t - i_img]
//   t >= i_img+F          : embed[ids[t - F + 1]]
// Block (n, 0) additionally emits kv_len[n] = (F-1) + #true(mask[n]) and
// seg_row[n] = n*T + (first s with ids[s+1]==seg) + F-1   (reference LISA.py:254-266), -1 if none.
__global__ void __launch_bounds__(128)
embed_splice_kernel(const long long* __restrict__ ids, const unsigned char* __restrict__ mask,
                    const bf16* __restrict__ embed, const bf16* __restrict__ feats,
                    bf16* __restrict__ out, int* __restrict__ kv_len, int* __restrict__ seg_row, int N,
                    int T_text, int F, int D, long long image_token, long long seg_token, int vocab) {
  const int T = T_text + F - 1;
  const int n = blockIdx.y, t = blockIdx.x;
  const long long* row_ids = ids + (size_t)n * T_text;
  __shared__ int s_img;
  if (threadIdx.x == 0) {
    int i_img = -1;
    for (int s = 0; s < T_text; ++s)
      if (row_ids[s] == image_token) { i_img = s; break; }
    s_img = i_img;
    if (t == 0) {
      int cnt = 0;
      for (int s = 0; s < T_text; ++s) cnt += mask ? (mask[(size_t)n * T_text + s] != 0) : 1;
      kv_len[n] = cnt + F - 1;
      int sp = -1;
      for (int s = 0; s + 1 < T_text; ++s)
        if (row_ids[s + 1] == seg_token) { sp = s; break; }
      seg_row[n] = sp < 0 ? -1 : n * T + sp + F - 1;
    }
  }
  __syncthreads();
  const int i_img = s_img;
  const bf16* src;
  if (i_img < 0) {
    src = nullptr;
  } else if (t < i_img) {
    src = embed + (size_t)row_ids[t] * D;
  } else if (t < i_img + F) {
    src = feats + ((size_t)n * F + (t - i_img)) * D;
  } else {
    src = embed + (size_t)row_ids[t - F + 1] * D;
  }
  if (src != nullptr && !(t >= i_img && t < i_img + F)) {
    const long long id = t < i_img ? row_ids[t] : row_ids[t - F + 1];
    if (id < 0 || id >= vocab) src = nullptr;  // malformed id: emit zeros rather than read OOB
  }
  uint4* o4 = reinterpret_cast<uint4*>(out + ((size_t)n * T + t) * D);
  const int nvec = D >> 3;
  if (src == nullptr) {
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) o4[i] = make_uint4(0, 0, 0, 0);
  } else {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) o4[i] = __ldg(s4 + i);
  }
}

// out[r, :] = bf16(x[r, :] + y[r / group, :])   (single-key cross attention: softmax over one key
// is 1, so the attention output is out_proj(v_proj(text)) broadcast over the K mask tokens —
// reference transformer.py:264-269 with keys of length 1)
__global__ void add_rows_bcast_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y,
                                      bf16* __restrict__ out, int rows, int dim, int group,
                                      const int* __restrict__ row_group) {
  const int nvec = dim >> 3;
  const long long total = (long long)rows * nvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / nvec), v = (int)(idx % nvec);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r * dim) + v);
    const int gidx = row_group ? row_group[r] : r / group;
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(y + (size_t)gidx * dim) + v);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = unpack_bf16(aw[e]), fb = unpack_bf16(bw[e]);
      o[e] = pack_bf16(fa.x + fb.x, fa.y + fb.y);
    }
    reinterpret_cast<uint4*>(out + (size_t)r * dim)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[r] = in[map[r]] (zero row when map[r] < 0): the [SEG] rows of the residual stream / attention output feeding
// the row-restricted tail of the last LLaMA layer (row-wise ops commute with the gather of LISA.py:322-337).
__global__ void gather_rows_kernel(const bf16* __restrict__ in, int ld_in, bf16* __restrict__ out, int rows,
                                   int dim, const int* __restrict__ map) {
  const int nvec = dim >> 3;
  const long long total = (long long)rows * nvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / nvec), v = (int)(idx % nvec);
    const int src = map[r];
    uint4 a = make_uint4(0, 0, 0, 0);
    if (src >= 0) a = __ldg(reinterpret_cast<const uint4*>(in + (size_t)src * ld_in) + v);
    reinterpret_cast<uint4*>(out + (size_t)r * dim)[v] = a;
  }
}

// 3x3 / pad 1 im2col on token-major NHWC: out[(b,y,x), (ky,kx,c)] = in[(b,y+ky-1,x+kx-1), c] or 0.
// One thread per 16-byte vector; feeds the SAM neck conv3x3 as a GEMM (image_encoder.py:100-106).
__global__ void im2col3x3_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H,
                                 int W, int Cc) {
  const int vec_per_tap = Cc >> 3;
  const long long total = (long long)B * H * W * 9 * vec_per_tap;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vec_per_tap);
    long long r = idx / vec_per_tap;
    const int tap = (int)(r % 9);
    r /= 9;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      val = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + yy) * W + xx) * Cc) + v);
    reinterpret_cast<uint4*>(out + (((size_t)b * H + y) * W + x) * (size_t)(9 * Cc) + tap * Cc)[v] = val;
  }
}

// One block per (listed sequence b, head h): every position s with pos_map[b*seq_in + s] < 0 (a window
// padding token) gets k[(b*H+h), s, :] = bias_k[h] and vt[(b*H+h), :, s] = bias_v[h].  seq_ids lists the
// sequences that have padding at all (9 of the 25 SAM windows), so no block is launched just to find out
// it has nothing to do; k writes are 16-byte, vt writes run along s.
__global__ void __launch_bounds__(256)
fill_kv_rows_kernel(bf16* __restrict__ k, bf16* __restrict__ vt, const bf16* __restrict__ bias,
                    const int* __restrict__ pos_map, const int* __restrict__ seq_ids, int heads, int hd,
                    int seq_in, int seq_pad) {
  extern __shared__ unsigned char s_pad[];  // [seq_in]
  const int b = seq_ids ? seq_ids[blockIdx.x] : blockIdx.x;
  const int h = blockIdx.y;
  int any = 0;
  for (int s = threadIdx.x; s < seq_in; s += blockDim.x) {
    const unsigned char f = pos_map[(size_t)b * seq_in + s] < 0;
    s_pad[s] = f;
    any |= f;
  }
  if (!__syncthreads_or(any)) return;
  const int hw = heads * hd;
  const size_t bh = (size_t)b * heads + h;
  const int vec = hd >> 3;
  const uint4* bk = reinterpret_cast<const uint4*>(bias + hw + h * hd);
  for (int idx = threadIdx.x; idx < seq_in * vec; idx += blockDim.x) {
    const int s = idx / vec, v = idx - s * vec;
    if (s_pad[s]) reinterpret_cast<uint4*>(k + (bh * seq_pad + s) * hd)[v] = __ldg(bk + v);
  }
  const bf16* bv = bias + 2 * hw + h * hd;
  // pairs of positions (seq_pad is even, so an even s is 4-byte aligned in every vt row): padding comes in
  // runs along s, most pairs are one 4-byte store
  const int half = (seq_in + 1) >> 1;
  for (int idx = threadIdx.x; idx < hd * half; idx += blockDim.x) {
    const int d = idx / half, s = (idx - d * half) * 2;
    const bool p0 = s_pad[s] != 0, p1 = s + 1 < seq_in && s_pad[s + 1] != 0;
    bf16* dst = vt + (bh * hd + d) * seq_pad + s;
    const bf16 v = bv[d];
    if (p0 && p1 && (seq_pad & 1) == 0) {
      *reinterpret_cast<__nv_bfloat162*>(dst) = __halves2bfloat162(v, v);
    } else {
      if (p0) dst[0] = v;
      if (p1) dst[1] = v;
    }
  }
}

}  // namespace
}  // namespace llmseg

using namespace llmseg;

extern "C" int llmseg_fill_kv_rows(void* k, void* vt, const void* bias_qkv, const int32_t* pos_map,
                                   const int32_t* seq_ids, int n_seqs, int heads, int head_dim, int seq_in,
                                   int seq_pad, void* stream) {
  const int batch = n_seqs;
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(k && vt && bias_qkv && pos_map, LLMSEG_EARG, "llmseg_fill_kv_rows: null pointer");
  LLMSEG_REQUIRE(batch > 0 && heads > 0 && heads < 65536 && head_dim > 0 && head_dim % 8 == 0 &&
                     seq_pad >= seq_in && seq_in > 0 && seq_in <= 32768,
                 LLMSEG_ESHAPE, "llmseg_fill_kv_rows: batch=%d heads=%d head_dim=%d seq_in=%d seq_pad=%d", batch,
                 heads, head_dim, seq_in, seq_pad);
  fill_kv_rows_kernel<<<dim3(batch, heads), 256, seq_in, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(k), static_cast<bf16*>(vt), static_cast<const bf16*>(bias_qkv), pos_map, seq_ids,
      heads, head_dim, seq_in, seq_pad);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_im2col3x3(const void* in, void* out, int batch, int height, int width,
                                int channels, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(in && out, LLMSEG_EARG, "llmseg_im2col3x3: null pointer");
  LLMSEG_REQUIRE(batch > 0 && height > 0 && width > 0 && channels % 8 == 0, LLMSEG_ESHAPE,
                 "llmseg_im2col3x3: %dx%dx%dx%d", batch, height, width, channels);
  const long long total = (long long)batch * height * width * 9 * (channels >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  im2col3x3_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), static_cast<bf16*>(out), batch, height, width, channels);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_patchify(const void* images, void* out, int batch, int img_size, int patch,
                               int k_pad, int cls_rows, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(images && out, LLMSEG_EARG, "llmseg_patchify: null pointer");
  LLMSEG_REQUIRE(batch > 0 && patch > 0 && img_size % patch == 0 && k_pad >= 3 * patch * patch &&
                     k_pad % 8 == 0 && (cls_rows == 0 || k_pad > 3 * patch * patch),
                 LLMSEG_ESHAPE, "llmseg_patchify: img=%d patch=%d k_pad=%d cls_rows=%d", img_size, patch,
                 k_pad, cls_rows);
  if (patch == 16)
    LLMSEG_REQUIRE((reinterpret_cast<uintptr_t>(images) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   LLMSEG_EALIGN, "llmseg_patchify: pointers must be 16-byte aligned");
  const int g = img_size / patch;
  const long long total = (long long)batch * (g * g + cls_rows) * (3 * patch + 1);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  patchify_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(images), static_cast<bf16*>(out), batch, img_size, patch, g, k_pad,
      cls_rows);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_embed_splice(const int64_t* input_ids, const uint8_t* attention_mask,
                                   const void* embed_table, const void* image_feats, void* out,
                                   int32_t* kv_len, int32_t* seg_row, int n_seq, int t_text,
                                   int n_img_tokens, int dim, int64_t image_token_id,
                                   int64_t seg_token_id, int vocab, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(input_ids && embed_table && image_feats && out && kv_len && seg_row, LLMSEG_EARG,
                 "llmseg_embed_splice: null pointer");
  LLMSEG_REQUIRE(n_seq > 0 && t_text > 0 && n_img_tokens > 0 && dim % 8 == 0, LLMSEG_ESHAPE,
                 "llmseg_embed_splice: n=%d t=%d F=%d dim=%d", n_seq, t_text, n_img_tokens, dim);
  dim3 grid(t_text + n_img_tokens - 1, n_seq);
  embed_splice_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(input_ids), attention_mask,
      static_cast<const bf16*>(embed_table), static_cast<const bf16*>(image_feats),
      static_cast<bf16*>(out), kv_len, seg_row, n_seq, t_text, n_img_tokens, dim, image_token_id,
      seg_token_id, vocab);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_add_rows_bcast(const void* x, const void* y, void* out, int rows, int dim,
                                     int group, const int32_t* row_group, void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(x && y && out, LLMSEG_EARG, "llmseg_add_rows_bcast: null pointer");
  LLMSEG_REQUIRE(rows > 0 && dim % 8 == 0 && (group > 0 || row_group), LLMSEG_ESHAPE,
                 "llmseg_add_rows_bcast: rows=%d dim=%d group=%d", rows, dim, group);
  const long long total = (long long)rows * (dim >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  add_rows_bcast_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<const bf16*>(y), static_cast<bf16*>(out), rows, dim,
      group, row_group);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" int llmseg_gather_rows(const void* in, int ld_in, void* out, int rows, int dim, const int32_t* src_row_map,
                                  void* stream) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(in && out && src_row_map, LLMSEG_EARG, "llmseg_gather_rows: null pointer");
  LLMSEG_REQUIRE(rows > 0 && dim > 0 && dim % 8 == 0 && ld_in % 8 == 0, LLMSEG_ESHAPE,
                 "llmseg_gather_rows: rows=%d dim=%d ld_in=%d", rows, dim, ld_in);
  LLMSEG_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 LLMSEG_EALIGN, "llmseg_gather_rows: pointers must be 16-byte aligned");
  const long long total = (long long)rows * (dim >> 3);
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  gather_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), ld_in, static_cast<bf16*>(out), rows, dim, src_row_map);
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}
