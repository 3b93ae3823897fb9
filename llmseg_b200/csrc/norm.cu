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
  const int warps_per_block = 8;
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
  const int per_lane = (dim / 8 + 31) / 32;
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
