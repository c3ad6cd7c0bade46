// Micro-benchmark: throughput of the f32 -> 16-bit pack conversions used by the epilogues (B200).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int MODE>
__global__ void k(const float* in, uint32_t* out, int iters) {
  float a = in[threadIdx.x], b = in[threadIdx.x + 1];
  uint32_t acc = 0;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      uint32_t r;
      float x = a + (float)j, y = b - (float)j;
      if (MODE == 0) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(y));
      if (MODE == 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(y));
      if (MODE == 2) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(y));
      if (MODE == 3) { x = fminf(fmaxf(x, -65504.f), 65504.f); y = fminf(fmaxf(y, -65504.f), 65504.f);
                       asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(y)); }
      acc ^= r;
    }
    a += 1.f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* in, uint32_t* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(in, out, 100);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(in, out, 2000);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = 148.0 * 8 * 256 * 2000 * 32;
  printf("%-28s %.3f ms  %.1f Gconv-pairs/s  (%.2f pairs/clk/SM @1.9GHz)\n", name, ms, n / ms / 1e6, n / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  float* in; uint32_t* out; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("bf16x2", in, out); run<1>("f16x2", in, out); run<2>("f16x2.satfinite", in, out); run<3>("f16x2 + fmin/fmax clamp", in, out);
  return 0;
}
