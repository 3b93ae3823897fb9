// llmseg_b200 — host-side runtime glue: error text, arch gate, TMA descriptor encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cudaTypedefs.h>

#include "common.cuh"

namespace llmseg {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_arch() {
  // cached per device id (devices never change capability)
  static std::atomic<int> ok_mask{0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return set_error(LLMSEG_ECUDA, "cudaGetDevice failed: %s (is there a GPU?)",
                     cudaGetErrorString(e));
  if (dev < 31 && (ok_mask.load() >> dev) & 1) return 0;
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess)
    return set_error(LLMSEG_ECUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  if (major != 10)
    return set_error(LLMSEG_EARCH,
                     "device %d is sm_%d%d; llmseg_b200 kernels are sm_100a only (no fallback)", dev,
                     major, minor);
  if (dev < 31) ok_mask.fetch_or(1 << dev);
  return 0;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  auto fn = get_encode_fn();
  LLMSEG_REQUIRE(fn != nullptr, LLMSEG_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  LLMSEG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, LLMSEG_EALIGN,
                 "TMA base pointer %p is not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    LLMSEG_REQUIRE((strides_bytes[i] & 15) == 0, LLMSEG_EALIGN,
                   "TMA stride %d = %llu bytes is not a multiple of 16", i,
                   (unsigned long long)strides_bytes[i]);
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LLMSEG_REQUIRE(r == CUDA_SUCCESS, LLMSEG_ECUDA,
                 "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu box %u,%u)",
                 (int)r, rank, (unsigned long long)dims[0],
                 (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
  return 0;
}

}  // namespace llmseg

extern "C" {
const char* llmseg_last_error(void) { return llmseg::g_err; }
int llmseg_version(void) { return 100; }
uint64_t llmseg_launch_count(void) { return llmseg::g_launches.load(); }
}
