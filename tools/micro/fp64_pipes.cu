// Microbenchmark: FP64 FMA pipe vs FP64 tensor (DMMA m8n8k4) vs both interleaved on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE> __global__ void k(double *out, double s) {
  double a[8], c0[8], c1[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-9 + i; c0[i] = i; c1[i] = -i; }
  double x = s, y = s * 0.5;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 2) { a[i] = fma(a[i], x, y); a[i] = fma(a[i], y, x); a[i] = fma(a[i], x, y); a[i] = fma(a[i], y, x);
                                    if (MODE == 0) { a[i] = fma(a[i], x, y); a[i] = fma(a[i], y, x); a[i] = fma(a[i], x, y); a[i] = fma(a[i], y, x); } }
      if (MODE == 1 || MODE == 2) { dmma(c0[i], c1[i], x, y); if (MODE == 1) dmma(c0[i], c1[i], y, x); }
    }
  }
  double r = 0;
  for (int i = 0; i < 8; i++) r += a[i] + c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name, double fma_per_thread_iter, double mma_per_warp_iter) {
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 1.0000001);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, 1.0000001);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double threads = 148.0 * 8 * 256;
  double flops = threads * ITERS * fma_per_thread_iter * 2 + (threads / 32) * ITERS * mma_per_warp_iter * 2 * 256;
  printf("%-12s %8.3f ms  %7.2f TFLOP/s  (err=%s)\n", name, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  run<0>("dfma", 64, 0);
  run<1>("dmma", 0, 16);
  run<2>("dfma+dmma", 32, 8);
  return 0;
}
