// Copy-engine peer bandwidth between GPU 0 and GPU 1: 1-D vs 2-D copies, one vs several streams, one vs both directions.
//   nvcc -O3 -o peer_copy peer_copy.cu && ./peer_copy
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t bytes = 8ull << 30;
  char *a0, *b0, *a1, *b1;
  CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, bytes)); CK(cudaMalloc(&b0, bytes)); cudaDeviceEnablePeerAccess(1, 0);
  CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, bytes)); CK(cudaMalloc(&b1, bytes)); cudaDeviceEnablePeerAccess(0, 0);
  cudaGetLastError();
  for (int both = 0; both < 2; both++)
    for (int mode = 0; mode < 3; mode++)        // 0: one 1-D copy, 1: 2-D (16 MiB rows, pitch 32 MiB), 2: 512 1-D copies of 16 MiB
      for (int ns : {1, 4}) {
        std::vector<cudaStream_t> s0(ns), s1(ns);
        CK(cudaSetDevice(0)); for (auto &s : s0) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        CK(cudaSetDevice(1)); for (auto &s : s1) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        cudaEvent_t e0, e1;
        CK(cudaSetDevice(0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
          CK(cudaSetDevice(0));
          CK(cudaEventRecord(e0, s0[0]));
          auto issue = [&](int dev, char *dst, char *src, std::vector<cudaStream_t> &st) {
            cudaSetDevice(dev);
            const size_t run = 16ull << 20, rows = bytes / 2 / run;  // the 2-D source spans the whole buffer with holes
            if (mode == 0) for (int j = 0; j < ns; j++) cudaMemcpyAsync(dst + j * (bytes / 2 / ns), src + j * (bytes / 2 / ns), bytes / 2 / ns, cudaMemcpyDeviceToDevice, st[j]);
            if (mode == 1) for (int j = 0; j < ns; j++) cudaMemcpy2DAsync(dst + j * (rows / ns) * run, run, src + j * (rows / ns) * 2 * run, 2 * run, run, rows / ns, cudaMemcpyDeviceToDevice, st[j]);
            if (mode == 2) for (size_t r = 0; r < rows; r++) cudaMemcpyAsync(dst + r * run, src + r * 2 * run, run, cudaMemcpyDeviceToDevice, st[r % ns]);
          };
          issue(0, b1, a0, s0);              // GPU 0 pushes 4 GiB into GPU 1
          if (both) issue(1, b0, a1, s1);    // and GPU 1 pushes 4 GiB into GPU 0
          for (auto &s : s0) CK(cudaStreamSynchronize(s));
          CK(cudaSetDevice(1)); for (auto &s : s1) CK(cudaStreamSynchronize(s));
          CK(cudaSetDevice(0));
          CK(cudaEventRecord(e1, s0[0])); CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        printf("%s  %-28s streams=%d  %7.2f ms  %6.1f GB/s per direction\n", both ? "both directions" : "one direction  ",
               mode == 0 ? "1-D, one copy per stream" : mode == 1 ? "2-D, 16 MiB rows" : "1-D, 16 MiB pieces", ns, best, (bytes / 2) / best / 1e6);
      }
  return 0;
}
