// Microbenchmark + layout check: tile rounds (two 2-qubit gates on round bits (0,1) and (2,3) of 16-amplitude blocks
// held in a swizzled shared-memory tile) on the FP64 tensor path: mma.sync.m8n8k4.f64, one amplitude per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_round dmma_round.cu && ./dmma_round
// Per 32-amplitude chunk and round: 1 LDS.128, 2 DMMA (gate A), 4 64-bit shuffles + 2 selects (accumulator layout of
// gate A -> B-operand layout of gate B), 2 DMMA (gate B), 2 STS.64.  Prints TFLOP/s for several CTA shapes and the
// max error against a host evaluation of the same rounds.
#include <cmath>
#include <cstdint>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef std::complex<double> cd;
constexpr int kTB = 12;

__host__ __device__ constexpr int swz_vec(int u) {
  return u < 3 ? (1 << u) : u == 3 ? 1 : u == 4 ? 2 : u == 5 ? 4 : u == 6 ? 3 : u == 7 ? 6 : u == 8 ? 5 : u == 9 ? 7 : u == 10 ? 1 : 2;
}
__host__ __device__ inline uint32_t phys_slot(uint32_t j) {
  uint32_t s = 0;
  for (int u = 3; u < kTB; u++)
    if ((j >> u) & 1u) s ^= (uint32_t)swz_vec(u);
  return j ^ s;
}

struct Round {
  unsigned short eoff[16];  // phys(sum_i bit_i(e) << pos[i])
  unsigned short gbit[8];   // phys(1 << gpos[i])
  double2 mats[2][16];      // row-major 4x4, matrix bit 0 <-> lower round bit
};
struct Params {
  Round r[4];
  int nrounds, reps;
};

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// one round over the tile by `nwarps` warps of this group; warp w takes chunks w, w + nwarps, ...
template <int U, int MODE>
__device__ __forceinline__ void round_dmma(double2 *tile, const Round &R, int w, int nwarps, int lane) {
  // A fragments: lane (r = lane / 4, c = lane % 4) holds Mreal[r][c] and Mreal[r][4 + c]
  const int r = lane >> 2, c = lane & 3;
  const double2 ma = R.mats[0][(r & 3) * 4 + c], mb = R.mats[1][(r & 3) * 4 + c];
  const double a0 = r < 4 ? ma.x : ma.y, a1 = r < 4 ? -ma.y : ma.x;
  const double b0 = r < 4 ? mb.x : mb.y, b1 = r < 4 ? -mb.y : mb.x;
  // gate-A operand layout: lane = b0 + 2 b1 + 4 b2 + 8 b3 + 16 x  (element e = lane & 15, x = group bit 0)
  const uint32_t lane_off = R.eoff[lane & 15] ^ ((lane >> 4) ? R.gbit[0] : 0);
  // transition: destination lane l' = b2 + 2 b3 + 4 b0 + 8 b1 + 16 x reads from s = 4 (b0 + 2 b1) + b3 + 2 x, slot b2
  const int d_b2 = lane & 1, d_b3 = (lane >> 1) & 1, d_b0 = (lane >> 2) & 1, d_b1 = (lane >> 3) & 1, d_x = lane >> 4;
  const int src = 4 * (d_b0 + 2 * d_b1) + d_b3 + 2 * d_x;
  // gate-B accumulator layout: lane (r, c) holds rows (b2 + 2 b3 [+4: im]) of columns n' = 2c + j = b0 + 2 b1 + 4 x
  const int o_b2 = r & 1, o_b3 = (r >> 1) & 1, o_b1 = c & 1, o_x = c >> 1;
  const uint32_t e_hi = (uint32_t)(o_b1 << 1 | o_b2 << 2 | o_b3 << 3);
  const uint32_t st0 = R.eoff[e_hi] ^ (o_x ? R.gbit[0] : 0), st1 = R.eoff[e_hi | 1] ^ (o_x ? R.gbit[0] : 0);
  const int im_off = r >= 4 ? 1 : 0;
  double *t = reinterpret_cast<double *>(tile);
  for (int C0 = w * U; C0 < 128; C0 += nwarps * U) {
    uint32_t base[U];
    double2 v[U];
    double d0[U], d1[U], e0[U], e1[U], xr[U], xi[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int C = C0 + u;
      base[u] = 0;
#pragma unroll
      for (int i = 0; i < 7; i++)
        if ((C >> i) & 1) base[u] ^= R.gbit[i + 1];
      if (MODE & 2) v[u] = make_double2(1.0 + C, 2.0 + lane); else v[u] = tile[base[u] ^ lane_off];
    }
#pragma unroll
    for (int u = 0; u < U; u++) { d0[u] = 0; d1[u] = 0; dmma(d0[u], d1[u], a0, v[u].x); }
#pragma unroll
    for (int u = 0; u < U; u++) dmma(d0[u], d1[u], a1, v[u].y);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (MODE & 1) { xr[u] = d0[u]; xi[u] = d1[u]; continue; }
      const double re0 = __shfl_sync(0xffffffffu, d0[u], src), re1 = __shfl_sync(0xffffffffu, d1[u], src);
      const double im0 = __shfl_sync(0xffffffffu, d0[u], src + 16), im1 = __shfl_sync(0xffffffffu, d1[u], src + 16);
      xr[u] = d_b2 ? re1 : re0;
      xi[u] = d_b2 ? im1 : im0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) { e0[u] = 0; e1[u] = 0; dmma(e0[u], e1[u], b0, xr[u]); }
#pragma unroll
    for (int u = 0; u < U; u++) dmma(e0[u], e1[u], b1, xi[u]);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (MODE & 2) { if (e0[u] + e1[u] == 12345.678) t[lane] = e0[u]; continue; }
      t[2 * (base[u] ^ st0) + im_off] = e0[u];
      t[2 * (base[u] ^ st1) + im_off] = e1[u];
    }
  }
}

template <int U, int MODE = 0>
__global__ void bench_kernel(double2 *io, const __grid_constant__ Params p, int nwarps) {
  extern __shared__ __align__(16) double2 tile[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < 4096; j += blockDim.x) tile[phys_slot(j)] = io[(size_t)blockIdx.x * 4096 + j];
  __syncthreads();
  for (int rep = 0; rep < p.reps; rep++)
    for (int r = 0; r < p.nrounds; r++) {
      round_dmma<U, MODE>(tile, p.r[r], w, nwarps, lane);
      __syncthreads();
    }
  for (int j = threadIdx.x; j < 4096; j += blockDim.x) io[(size_t)blockIdx.x * 4096 + j] = tile[phys_slot(j)];
}

static void host_round(std::vector<cd> &t, const int pos[4], const cd ma[16], const cd mb[16]) {
  for (int gate = 0; gate < 2; gate++) {
    const int p0 = pos[2 * gate], p1 = pos[2 * gate + 1];
    const cd *m = gate ? mb : ma;
    for (int j = 0; j < 4096; j++) {
      if ((j >> p0 & 1) || (j >> p1 & 1)) continue;
      const int idx[4] = {j, j | 1 << p0, j | 1 << p1, j | 1 << p0 | 1 << p1};
      cd x[4], y[4];
      for (int a = 0; a < 4; a++) x[a] = t[idx[a]];
      for (int a = 0; a < 4; a++) { y[a] = 0; for (int b = 0; b < 4; b++) y[a] += m[a * 4 + b] * x[b]; }
      for (int a = 0; a < 4; a++) t[idx[a]] = y[a];
    }
  }
}

int main() {
  srand(1);
  Params p;
  p.nrounds = 4;
  int pos[4][4] = {{3, 5, 8, 10}, {0, 1, 6, 11}, {2, 4, 7, 9}, {1, 3, 9, 11}};
  std::vector<std::vector<cd>> M(8, std::vector<cd>(16));
  for (int r = 0; r < 4; r++) {
    int gpos[8], ng = 0;
    for (int u = 0; u < 12; u++) {
      bool used = false;
      for (int i = 0; i < 4; i++) used = used || pos[r][i] == u;
      if (!used) gpos[ng++] = u;
    }
    for (int e = 0; e < 16; e++) {
      uint32_t j = 0;
      for (int i = 0; i < 4; i++)
        if ((e >> i) & 1) j |= 1u << pos[r][i];
      p.r[r].eoff[e] = (unsigned short)phys_slot(j);
    }
    for (int i = 0; i < 8; i++) p.r[r].gbit[i] = (unsigned short)phys_slot(1u << gpos[i]);
    for (int g = 0; g < 2; g++)
      for (int i = 0; i < 16; i++) {
        M[2 * r + g][i] = cd((rand() % 2001 - 1000) / 2000.0, (rand() % 2001 - 1000) / 2000.0) * 0.5;
        p.r[r].mats[g][i] = make_double2(M[2 * r + g][i].real(), M[2 * r + g][i].imag());
      }
  }
  const int nblk = 148;
  std::vector<cd> h((size_t)nblk * 4096);
  for (auto &v : h) v = cd((rand() % 2001 - 1000) / 1000.0, (rand() % 2001 - 1000) / 1000.0);
  double2 *d;
  cudaMalloc(&d, h.size() * 16);
  cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(bench_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(bench_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  std::vector<cd> out(h.size());
  for (int U : {1, 2, 4, 8}) {
    cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice);
    p.reps = 1;
    if (U == 1) bench_kernel<1><<<nblk, 512, 65536>>>(d, p, 16);
    if (U == 2) bench_kernel<2><<<nblk, 512, 65536>>>(d, p, 16);
    if (U == 4) bench_kernel<4><<<nblk, 512, 65536>>>(d, p, 16);
    if (U == 8) bench_kernel<8><<<nblk, 512, 65536>>>(d, p, 16);
    cudaMemcpy(out.data(), d, h.size() * 16, cudaMemcpyDeviceToHost);
    std::vector<cd> ref(h.begin(), h.begin() + 4096);
    for (int r = 0; r < 4; r++) host_round(ref, pos[r], M[2 * r].data(), M[2 * r + 1].data());
    double err = 0;
    for (int j = 0; j < 4096; j++) err = fmax(err, std::abs(out[j] - ref[j]));
    printf("U=%d max |err| vs host rounds: %.3e  (%s)\n", U, err, cudaGetErrorString(cudaGetLastError()));
  }
  p.reps = 400;
  cudaFuncSetAttribute(bench_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(bench_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(bench_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int mode = 0; mode < 4; mode++)
    for (int nw : {8, 16, 32}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int it = 0; it < 2; it++) {
        cudaEventRecord(e0);
        if (mode == 0) bench_kernel<2, 0><<<148, nw * 32, 65536>>>(d, p, nw);
        if (mode == 1) bench_kernel<2, 1><<<148, nw * 32, 65536>>>(d, p, nw);
        if (mode == 2) bench_kernel<2, 2><<<148, nw * 32, 65536>>>(d, p, nw);
        if (mode == 3) bench_kernel<2, 3><<<148, nw * 32, 65536>>>(d, p, nw);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 148.0 * p.reps * p.nrounds * 4096.0 * 2 * 16 * 2;
      printf("mode=%d (1: no shuffles, 2: no smem) warps/CTA=%2d  %8.3f ms  %6.2f TFLOP/s (%s)\n", mode, nw, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
