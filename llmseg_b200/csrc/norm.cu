// llmseg_b200 — LayerNorm / RMSNorm rows (HBM-bound).
//
// One warp per row; the row is read ONCE with 16-byte vector loads into registers (dim <= 4096),
// mean/variance by warp-shuffle reduction in fp32, normalised and written with 16-byte stores.
// Algorithmic traffic = 2 * rows * dim * 2 bytes (+ gamma/beta, L2-resident).
//
// LayerNorm follows torch.nn.LayerNorm on a bf16 tensor (fp32 statistics, biased variance,
// one rounding to bf16 at the end) — SAM norm1/norm2 (reference image_encoder.py:160,170,179,191),
// LayerNorm2d on NHWC rows (common.py:31-43: same formula per pixel), CLIP and selector norms.
// RMSNorm follows transformers LlamaRMSNorm: x32*rsqrt(mean(x32^2)+eps) → bf16, then * weight → bf16.
#include <atomic>

#include "common.cuh"

namespace llmseg {
extern std::atomic<uint64_t> g_launches;
namespace {

constexpr int MAX_VEC = 16;  // 16 * 32 lanes * 8 elems = 4096 max dim

// NV = 16-byte vectors held per lane, sized EXACTLY to the row (1280 -> 5, 1024 -> 4, 4096 -> 16) so
// narrow rows keep few registers and many warps — i.e. many loads — in flight per SM (measured: the
// 8-vector version ran LayerNorm(1280) at 3.1 TB/s against a 6.4 TB/s copy, profiles/r01i_hbm.md).
__device__ __forceinline__ uint4 ld_stream16(const uint4* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
template <bool RMS, int NV>
__global__ void __launch_bounds__(256, (NV <= 5 ? 5 : (NV <= 8 ? 4 : 2)))
norm_kernel(const bf16* __restrict__ in, int ld_in, bf16* __restrict__ out, int ld_out,
            const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int rows, int dim, float eps,
            const int* __restrict__ src_map, float2* __restrict__ stats) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = dim >> 3;  // 8 bf16 per 16-byte vector
  int src = warp;
  if (src_map) src = src_map[warp];
  uint4* orow = reinterpret_cast<uint4*>(out + (size_t)warp * ld_out);
  if (src < 0 && stats == nullptr) {
    for (int i = lane; i < nvec; i += 32) orow[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* irow = reinterpret_cast<const uint4*>(in + (size_t)src * ld_in);
  uint4 v[NV];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int i = lane + t * 32;
    if (i < nvec) {
      v[t] = ld_stream16(irow + i);
      const uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = unpack_bf16(w[e]);
        s += f.x + f.y;
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float inv_n = 1.0f / (float)dim;
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(ss * inv_n + eps);
  } else {
    mean = s * inv_n;
    // two-pass variance from registers for accuracy (the row is already resident)
    float sq = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      const int i = lane + t * 32;
      if (i < nvec) {
        const uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = unpack_bf16(w[e]);
          const float a = f.x - mean, b = f.y - mean;
          sq += a * a + b * b;
        }
      }
    }
    sq = warp_sum(sq);
    rstd = rsqrtf(sq * inv_n + eps);
  }
  if (stats != nullptr) {
    // statistics only: the normalisation itself is folded into the consuming GEMM (llmseg_gemm row_stats)
    if (lane == 0) stats[warp] = make_float2(mean, rstd);
    return;
  }
  const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
  const uint4* b4 = reinterpret_cast<const uint4*>(beta);
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int i = lane + t * 32;
    if (i < nvec) {
      const uint4 g = __ldg(g4 + i);
      uint4 b = make_uint4(0, 0, 0, 0);
      if (!RMS && beta) b = __ldg(b4 + i);
      const uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        const float2 gg = unpack_bf16(gw[e]);
        if (RMS) {
          // HF: (x32 * rsqrt(var+eps)).to(bf16) * weight
          const float a = bf16_round(f.x * rstd), c = bf16_round(f.y * rstd);
          o[e] = pack_bf16(a * gg.x, c * gg.y);
        } else {
          const float2 bb = unpack_bf16(bw[e]);
          o[e] = pack_bf16((f.x - mean) * rstd * gg.x + bb.x, (f.y - mean) * rstd * gg.y + bb.y);
        }
      }
      orow[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Narrow rows (dim <= 256: one 16-byte vector per lane), many of them (the 1 M x 256 token stream of the SAM mask
// decoder, the 256-channel neck): a warp takes R consecutive rows at once, so R loads per lane are in flight instead
// of one — the one-row kernel above keeps 40 warps x 512 B = 20 KB per SM in flight and measured 3.8 TB/s on
// LayerNorm(1 M x 256).  Same arithmetic as norm_kernel (two-pass variance from registers).
// ---------------------------------------------------------------------------------------------
template <bool RMS, int R>
__global__ void __launch_bounds__(256)
norm_narrow_kernel(const bf16* __restrict__ in, int ld_in, bf16* __restrict__ out, int ld_out,
                   const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int rows, int dim, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * R;
  if (row0 >= rows) return;
  const int nvec = dim >> 3;
  const bool on = lane < nvec;
  uint4 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    v[r] = make_uint4(0, 0, 0, 0);
    if (on && row0 + r < rows) v[r] = ld_stream16(reinterpret_cast<const uint4*>(in + (size_t)(row0 + r) * ld_in) + lane);
  }
  uint4 g = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
  if (on) {
    g = __ldg(reinterpret_cast<const uint4*>(gamma) + lane);
    if (!RMS && beta) b = __ldg(reinterpret_cast<const uint4*>(beta) + lane);
  }
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, bw[4] = {b.x, b.y, b.z, b.w};
  const float inv_n = 1.0f / (float)dim;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= rows) break;   // warp-uniform
    const uint32_t w[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
    float f[8], s = 0.f, ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = unpack_bf16(w[e]);
      f[2 * e] = t.x;
      f[2 * e + 1] = t.y;
      s += t.x + t.y;
      ss += t.x * t.x + t.y * t.y;
    }
    float mean = 0.f, rstd;
    if (RMS) {
      rstd = rsqrtf(warp_sum(ss) * inv_n + eps);
    } else {
      mean = warp_sum(s) * inv_n;
      float sq = 0.f;
      if (on) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {   // pairwise like norm_kernel: bit-equal statistics
          const float a = f[2 * e] - mean, c = f[2 * e + 1] - mean;
          sq += a * a + c * c;
        }
      }
      rstd = rsqrtf(warp_sum(sq) * inv_n + eps);
    }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 gg = unpack_bf16(gw[e]);
      if (RMS) {
        const float a = bf16_round(f[2 * e] * rstd), c = bf16_round(f[2 * e + 1] * rstd);
        o[e] = pack_bf16(a * gg.x, c * gg.y);
      } else {
        const float2 bb = unpack_bf16(bw[e]);
        o[e] = pack_bf16((f[2 * e] - mean) * rstd * gg.x + bb.x, (f[2 * e + 1] - mean) * rstd * gg.y + bb.y);
      }
    }
    if (on) reinterpret_cast<uint4*>(out + (size_t)(row0 + r) * ld_out)[lane] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming variant for large problems (>= 1 MB of rows): persistent, one CTA per SM, every warp owns a ring of
// row buffers in shared memory that `cp.async.bulk` (the TMA engine's 1-D copy) keeps full — up to 24 KB per warp,
// 192 KB per SM in flight with no register cost, and loads stay in flight while the warp normalises and stores
// the previous row.  The register-resident kernel above overlaps nothing inside a warp (load, then reduce, then
// store) and measured 46-63 % of the copy bandwidth (profiles/r01i_hbm_after.md).
// ---------------------------------------------------------------------------------------------
// Rows of >= 2 KB only (narrower rows stay on the register kernel: per-row overhead dominates there).
// Warps per CTA follow the row width (8 for >= 4 KB rows, 16 for >= 2 KB, 24 below): the 192 KB of ring space is
// split evenly between them, and narrow rows need more warps to keep the per-row latency chain (barrier wait,
// three shuffle reductions, refill) off the critical path.
constexpr int NS_MAX_WARPS = 24;
constexpr int NS_RING_TOTAL = 196608;
constexpr int NS_MAX_STAGES = 8;
constexpr int NS_SMEM_BYTES = NS_RING_TOTAL + NS_MAX_WARPS * NS_MAX_STAGES * (8 + 4) + 128;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool RMS>
__global__ void __launch_bounds__(NS_MAX_WARPS * 32, 1)
norm_stream_kernel(const bf16* __restrict__ in, int ld_in, bf16* __restrict__ out, int ld_out,
                   const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int rows, int dim, float eps,
                   const int* __restrict__ src_map, float2* __restrict__ stats, int stages) {
  extern __shared__ uint8_t ns_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ns_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)dim * 2u;
  const int stage_bytes = (int)((row_bytes + 127u) & ~127u);
  const int n_warps = blockDim.x >> 5;
  uint8_t* ring = smem + warp * ((NS_RING_TOTAL / n_warps) & ~127);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS_RING_TOTAL) + warp * NS_MAX_STAGES;
  int* absent = reinterpret_cast<int*>(smem + NS_RING_TOTAL + NS_MAX_WARPS * NS_MAX_STAGES * 8) + warp * NS_MAX_STAGES;
  const int total_warps = gridDim.x * n_warps;
  const int w0 = blockIdx.x * n_warps + warp;
  const int nvec = dim >> 3;
  const float inv_n = 1.0f / (float)dim;

  // lane 0: start the copy of output row `row` into stage st (src_map < 0: nothing to copy, a zero row)
  auto issue = [&](int st, int row) {
    const int src = src_map ? src_map[row] : row;
    absent[st] = src < 0;
    if (src >= 0) {
      mbar_expect_tx(&bars[st], row_bytes);
      bulk_g2s(ring + st * stage_bytes, in + (size_t)src * ld_in, row_bytes, &bars[st]);
    }
  };
  if (lane == 0) {
    for (int st = 0; st < stages; ++st) mbar_init(&bars[st], 1);
    fence_barrier_init();
    for (int st = 0; st < stages; ++st) {
      const int row = w0 + st * total_warps;
      if (row < rows) issue(st, row);
    }
  }
  __syncwarp();
  uint32_t phases = 0;
  int st = 0;
  const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
  const uint4* b4 = reinterpret_cast<const uint4*>(beta);
  for (int row = w0; row < rows; row += total_warps) {
    const bool none = absent[st] != 0;
    uint4* orow = reinterpret_cast<uint4*>(out + (size_t)row * ld_out);
    if (none) {
      if (stats == nullptr)
        for (int i = lane; i < nvec; i += 32) orow[i] = make_uint4(0, 0, 0, 0);
      else if (lane == 0) stats[row] = make_float2(0.f, rsqrtf(eps));
    } else {
      mbar_wait(&bars[st], (phases >> st) & 1u);
      phases ^= 1u << st;
      const uint4* v4 = reinterpret_cast<const uint4*>(ring + st * stage_bytes);
      // statistics on packed fp32 pairs (FADD2 / FFMA2): the kernel is close to issue-bound once the loads are
      // hidden, so instructions per element decide the achieved bandwidth
      float2 s2 = make_float2(0.f, 0.f), ss2 = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int i = lane; i < nvec; i += 32) {
        const uint4 v = v4[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w[e]);
          s2 = fadd2(s2, f);
          ss2 = ffma2(f, f, ss2);
        }
      }
      const float s = warp_sum(s2.x + s2.y);
      const float ss = warp_sum(ss2.x + ss2.y);
      float mean = 0.f, rstd;
      if (RMS) {
        rstd = rsqrtf(ss * inv_n + eps);
      } else {
        mean = s * inv_n;
        const float2 nm = make_float2(-mean, -mean);
        float2 sq2 = make_float2(0.f, 0.f);  // two-pass variance (the row is resident in shared memory)
#pragma unroll 4
        for (int i = lane; i < nvec; i += 32) {
          const uint4 v = v4[i];
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = fadd2(unpack_bf16(w[e]), nm);
            sq2 = ffma2(a, a, sq2);
          }
        }
        rstd = rsqrtf(warp_sum(sq2.x + sq2.y) * inv_n + eps);
      }
      if (stats != nullptr) {
        if (lane == 0) stats[row] = make_float2(mean, rstd);
      } else {
        const float2 r2 = make_float2(rstd, rstd);
        const float2 sh = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll 2
        for (int i = lane; i < nvec; i += 32) {
          const uint4 v = v4[i];
          const uint4 g = __ldg(g4 + i);
          uint4 b = make_uint4(0, 0, 0, 0);
          if (!RMS && beta) b = __ldg(b4 + i);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
          const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16(w[e]);
            const float2 gg = unpack_bf16(gw[e]);
            if (RMS) {
              // HF: (x32 * rsqrt(var+eps)).to(bf16) * weight
              const float2 a = fmul2(f, r2);
              const float2 ar = unpack_bf16(pack_bf16(a.x, a.y));
              const float2 y = fmul2(ar, gg);
              o[e] = pack_bf16(y.x, y.y);
            } else {
              // ((x - mean) * rstd) * gamma + beta, the first product as x * rstd + (-mean * rstd)
              const float2 y = ffma2(ffma2(f, r2, sh), gg, unpack_bf16(bw[e]));
              o[e] = pack_bf16(y.x, y.y);
            }
          }
          orow[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    // every lane is done with this stage: refill it with the row `stages` steps ahead
    __syncwarp();
    if (lane == 0) {
      const int next = row + stages * total_warps;
      if (next < rows) {
        fence_async_smem();
        issue(st, next);
      }
    }
    __syncwarp();
    st = st + 1 == stages ? 0 : st + 1;
  }
}

// LLMSEG_NORM_STREAM=0: always the register-resident kernel (A/B; read per call)
bool stream_norm_enabled() {
  const char* e = getenv("LLMSEG_NORM_STREAM");
  return e == nullptr || atoi(e) != 0;
}

template <bool RMS>
int launch_norm(const void* in, int ld_in, void* out, int ld_out, const void* gamma, const void* beta,
                int rows, int dim, float eps, const int32_t* src_map, cudaStream_t stream,
                float2* stats = nullptr) {
  if (int e = check_arch()) return e;
  LLMSEG_REQUIRE(in && (stats || (out && gamma)), LLMSEG_EARG, "norm: null pointer");
  LLMSEG_REQUIRE(rows > 0 && dim > 0 && dim % 8 == 0 && dim <= MAX_VEC * 256, LLMSEG_ESHAPE,
                 "norm: rows=%d dim=%d unsupported (dim %% 8 == 0, dim <= %d)", rows, dim,
                 MAX_VEC * 256);
  LLMSEG_REQUIRE(ld_in % 8 == 0 && ld_out % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                 LLMSEG_EALIGN, "norm: pointers / leading dims must be 16-byte aligned");
  if ((size_t)rows * dim * 2 >= (1u << 20) && dim >= 1024 && stream_norm_enabled()) {
    // large problem: persistent bulk-copy-staged kernel (see norm_stream_kernel)
    auto kern = norm_stream_kernel<RMS>;
    static bool attr_done = false;
    if (!attr_done) {
      LLMSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NS_SMEM_BYTES));
      attr_done = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int stage_bytes = (dim * 2 + 127) & ~127;
    const int n_warps = stage_bytes >= 4096 ? 8 : (stage_bytes >= 2048 ? 16 : NS_MAX_WARPS);
    int stages = ((NS_RING_TOTAL / n_warps) & ~127) / stage_bytes;
    if (stages > NS_MAX_STAGES) stages = NS_MAX_STAGES;
    const int need = (rows + n_warps - 1) / n_warps;
    const int grid = need < sms ? need : sms;
    kern<<<grid, n_warps * 32, NS_SMEM_BYTES, stream>>>(
        static_cast<const bf16*>(in), ld_in, static_cast<bf16*>(out), ld_out, static_cast<const bf16*>(gamma),
        static_cast<const bf16*>(beta), rows, dim, eps, src_map, stats, stages);
    LLMSEG_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
  }
  const int warps_per_block = 8;
  const int per_lane = (dim / 8 + 31) / 32;
  if (per_lane <= 1 && rows >= 8192 && src_map == nullptr && stats == nullptr && out != nullptr) {
    constexpr int R = 4;
    const int nb = ((rows + R - 1) / R + warps_per_block - 1) / warps_per_block;
    norm_narrow_kernel<RMS, R><<<nb, warps_per_block * 32, 0, stream>>>(
        static_cast<const bf16*>(in), ld_in, static_cast<bf16*>(out), ld_out, static_cast<const bf16*>(gamma),
        static_cast<const bf16*>(beta), rows, dim, eps);
    LLMSEG_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
  }
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
#define LLMSEG_NORM_LAUNCH(NV_)                                                              \
  norm_kernel<RMS, NV_><<<blocks, warps_per_block * 32, 0, stream>>>(                        \
      static_cast<const bf16*>(in), ld_in, static_cast<bf16*>(out), ld_out,                  \
      static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), rows, dim, eps, src_map, stats)
  if (per_lane <= 1) LLMSEG_NORM_LAUNCH(1);
  else if (per_lane <= 2) LLMSEG_NORM_LAUNCH(2);
  else if (per_lane <= 3) LLMSEG_NORM_LAUNCH(3);
  else if (per_lane <= 4) LLMSEG_NORM_LAUNCH(4);
  else if (per_lane <= 5) LLMSEG_NORM_LAUNCH(5);
  else if (per_lane <= 8) LLMSEG_NORM_LAUNCH(8);
  else LLMSEG_NORM_LAUNCH(16);
#undef LLMSEG_NORM_LAUNCH
  LLMSEG_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // namespace
}  // namespace llmseg

extern "C" int llmseg_layernorm(const void* in, int ld_in, void* out, int ld_out, const void* gamma,
                                const void* beta, int rows_out, int dim, float eps,
                                const int32_t* src_row_map, void* stream) {
  return llmseg::launch_norm<false>(in, ld_in, out, ld_out, gamma, beta, rows_out, dim, eps,
                                    src_row_map, static_cast<cudaStream_t>(stream));
}

extern "C" int llmseg_rmsnorm(const void* in, int ld_in, void* out, int ld_out, const void* gamma,
                              int rows, int dim, float eps, const int32_t* src_row_map, void* stream) {
  return llmseg::launch_norm<true>(in, ld_in, out, ld_out, gamma, nullptr, rows, dim, eps,
                                   src_row_map, static_cast<cudaStream_t>(stream));
}

extern "C" int llmseg_norm_stats(const void* in, int ld_in, int rows, int dim, float eps, int rms, void* stats,
                                 void* stream) {
  LLMSEG_REQUIRE(stats != nullptr && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, LLMSEG_EARG,
                 "llmseg_norm_stats: stats must be a non-null, 8-byte aligned float2 array");
  float2* st = static_cast<float2*>(stats);
  if (rms)
    return llmseg::launch_norm<true>(in, ld_in, nullptr, 8, nullptr, nullptr, rows, dim, eps, nullptr,
                                     static_cast<cudaStream_t>(stream), st);
  return llmseg::launch_norm<false>(in, ld_in, nullptr, 8, nullptr, nullptr, rows, dim, eps, nullptr,
                                    static_cast<cudaStream_t>(stream), st);
}
