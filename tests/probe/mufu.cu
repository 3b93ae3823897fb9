// Test-only microbenchmark: MUFU.EX2 and FFMA issue throughput per SM (clock64 based).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mufu_kernel(float* out, long long* cycles, int iters, int mode) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
  unsigned long long b0 = threadIdx.x, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3, b4 = b0 + 4, b5 = b0 + 5, b6 = b0 + 6, b7 = b0 + 7;
  const float2 cm2 = make_float2(1.0001f, 0.9999f), ca2 = make_float2(0.5f, 0.25f);
  const unsigned long long cm = *reinterpret_cast<const unsigned long long*>(&cm2), ca = *reinterpret_cast<const unsigned long long*>(&ca2);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
    } else if (mode == 2) {
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b0) : "l"(cm), "l"(ca)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b1) : "l"(cm), "l"(ca));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b2) : "l"(cm), "l"(ca)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b3) : "l"(cm), "l"(ca));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b4) : "l"(cm), "l"(ca)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b5) : "l"(cm), "l"(ca));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b6) : "l"(cm), "l"(ca)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b7) : "l"(cm), "l"(ca));
    } else {
      a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
      a4 = fmaf(a4, 1.0001f, 0.5f); a5 = fmaf(a5, 1.0001f, 0.5f); a6 = fmaf(a6, 1.0001f, 0.5f); a7 = fmaf(a7, 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(b0 ^ b1 ^ b2 ^ b3 ^ b4 ^ b5 ^ b6 ^ b7);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      const int iters = 4096;
      mufu_kernel<<<148, threads>>>(out, cyc, iters, mode); cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double ops = (double)threads * iters * 8 * (mode == 2 ? 2 : 1);
      printf("%s threads/SM=%4d : %.2f ops/clk/SM\n", mode == 0 ? "MUFU.EX2" : mode == 1 ? "FFMA    " : "FFMA2 x2", threads, ops / (double)h[0]);
    }
  return 0;
}
