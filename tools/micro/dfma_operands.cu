// Microbenchmark: DFMA throughput with register vs constant-bank (kernel parameter) operands on B200.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
struct P { double m[32]; };
template <int MODE> __global__ void __launch_bounds__(256) k(double *out, const __grid_constant__ P p, const double *gm) {
  double a[16];
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-9 + i;
  double r[32];
  if (MODE == 1) for (int i = 0; i < 32; i++) r[i] = gm[i];
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        if (MODE == 0) { a[i] = fma(a[i], p.m[j], a[i + 1]); a[i + 1] = fma(a[i + 1], p.m[j], a[i + 2]); a[i + 2] = fma(a[i + 2], p.m[j], a[i + 3]); a[i + 3] = fma(a[i + 3], p.m[j], a[i]); }
        else           { a[i] = fma(a[i], r[j], a[i + 1]);   a[i + 1] = fma(a[i + 1], r[j], a[i + 2]);   a[i + 2] = fma(a[i + 2], r[j], a[i + 3]);   a[i + 3] = fma(a[i + 3], r[j], a[i]); }
      }
    }
  }
  double s = 0; for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int blocks_per_sm) {
  double *out, *gm; cudaMalloc(&out, 148 * 8 * 256 * 8); cudaMalloc(&gm, 256); cudaMemset(gm, 0, 256);
  P p; for (int i = 0; i < 32; i++) p.m[i] = 1.0 + 1e-9 * i;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * blocks_per_sm, 256>>>(out, p, gm); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * blocks_per_sm, 256>>>(out, p, gm); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 148.0 * blocks_per_sm * 256 * ITERS * 32 * 16 * 2;
  printf("%-28s blocks/SM=%d %8.3f ms %7.2f TFLOP/s\n", name, blocks_per_sm, ms, fl / ms / 1e9);
}
int main() {
  run<0>("dfma const-bank operand", 2); run<0>("dfma const-bank operand", 4); run<0>("dfma const-bank operand", 1);
  run<1>("dfma register operand", 2); run<1>("dfma register operand", 4); run<1>("dfma register operand", 1);
  return 0;
}
