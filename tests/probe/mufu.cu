// Test-only microbenchmark: MUFU.EX2 and FFMA issue throughput per SM (clock64 based).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mufu_kernel(float* out, long long* cycles, int iters, int mode) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
    } else {
      a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
      a4 = fmaf(a4, 1.0001f, 0.5f); a5 = fmaf(a5, 1.0001f, 0.5f); a6 = fmaf(a6, 1.0001f, 0.5f); a7 = fmaf(a7, 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      const int iters = 4096;
      mufu_kernel<<<148, threads>>>(out, cyc, iters, mode); cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double ops = (double)threads * iters * 8;
      printf("%s threads/SM=%4d : %.2f ops/clk/SM\n", mode == 0 ? "MUFU.EX2" : "FFMA    ", threads, ops / (double)h[0]);
    }
  return 0;
}
