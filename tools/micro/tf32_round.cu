// TF32 study for the single-precision tile rounds: a round = two 4x4 complex gates on bit pairs
// (0,1) and (2,3) of 16-amplitude blocks.  Variant T: the gate as a real 8x8 matrix on the tensor cores,
// mma.sync.m16n8k8 tf32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi) for FP32-level accuracy; amplitudes are the A
// operand (16 groups x 8 reals), the matrix is B, and the accumulator layout equals the A layout (lane (g, k) owns
// amplitude k of groups g and g + 8), so consecutive gates need one lane PERMUTATION (swap lane bits (0,1) <-> (2,3):
// 4 SHFL.32) and no selects.  Variant F: the same round with FFMA on 16-amplitude register blocks (what
// tile_pipe2_kernel<3> does today).  Registers only (no shared memory): this measures the arithmetic side.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_round tf32_round.cu && ./tf32_round
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef std::complex<double> cd;

struct Mats {
  float2 m[2][16];  // row-major 4x4 complex, two gates
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// B fragment of gate M (row-major 4x4 complex) for lane (n = lane >> 2: output real, k = lane & 3): W[k][n], W[k+4][n]
// with W[j][2c] = Re M[c][j], W[4+j][2c] = -Im M[c][j], W[j][2c+1] = Im M[c][j], W[4+j][2c+1] = Re M[c][j]
__device__ __forceinline__ void bfrag(const float2 *M, int lane, float &w0, float &w1) {
  const int n = lane >> 2, k = lane & 3, c = n >> 1;
  const float2 e = M[c * 4 + k];
  if (n & 1) { w0 = e.y; w1 = e.x; } else { w0 = e.x; w1 = -e.y; }
}
template <bool SPLIT3>
__device__ __forceinline__ void gate_mma(float2 &x0, float2 &x1, float w0, float w1) {
  // x0 / x1: amplitude k of group g / g + 8.  A = (re0, re1, im0, im1) as (a0, a1, a2, a3)
  const float av[4] = {x0.x, x1.x, x0.y, x1.y};
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    ah[i] = to_tf32(av[i]);
    al[i] = to_tf32(av[i] - __uint_as_float(ah[i]));
  }
  bh[0] = to_tf32(w0); bh[1] = to_tf32(w1);
  bl[0] = to_tf32(w0 - __uint_as_float(bh[0])); bl[1] = to_tf32(w1 - __uint_as_float(bh[1]));
  float c[4] = {0, 0, 0, 0};
  if (SPLIT3) { mma_tf32(c, al, bh); mma_tf32(c, ah, bl); }
  mma_tf32(c, ah, bh);
  x0 = make_float2(c[0], c[1]);
  x1 = make_float2(c[2], c[3]);
}
// MODE 0: FFMA register blocks; 1: TF32 single pass; 2: 3xTF32
template <int MODE>
__global__ void round_kernel(float2 *io, const __grid_constant__ Mats p, int reps) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (MODE == 0) {
    // thread = one 16-amplitude block; 32 blocks per warp = 512 amplitudes
    float2 a[16];
    float2 *base = io + warp * 512 + lane * 16;
    for (int e = 0; e < 16; e++) a[e] = base[e];
    for (int r = 0; r < reps; r++)
#pragma unroll
      for (int gte = 0; gte < 2; gte++) {
        const float2 *M = p.m[gte];
#pragma unroll
        for (int o = 0; o < 4; o++) {
          int idx[4];
#pragma unroll
          for (int t = 0; t < 4; t++) idx[t] = gte == 0 ? (o * 4 + t) : (o + 4 * t);
          float2 x[4], y[4];
#pragma unroll
          for (int t = 0; t < 4; t++) x[t] = a[idx[t]];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            y[i] = make_float2(0, 0);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 m = M[i * 4 + j];
              y[i].x = fmaf(m.x, x[j].x, y[i].x); y[i].x = fmaf(-m.y, x[j].y, y[i].x);
              y[i].y = fmaf(m.x, x[j].y, y[i].y); y[i].y = fmaf(m.y, x[j].x, y[i].y);
            }
          }
#pragma unroll
          for (int t = 0; t < 4; t++) a[idx[t]] = y[t];
        }
      }
    for (int e = 0; e < 16; e++) base[e] = a[e];
  } else {
    // warp = 8 chunks of 64 amplitudes; chunk layout: amplitude index j (6 bits) = b0 b1 | b2 b3 | x | y; lane = j & 31
    // for gate A (k = b0 b1 in lane bits 0..1), register = y
    float wA0, wA1, wB0, wB1;
    bfrag(p.m[0], lane, wA0, wA1);
    bfrag(p.m[1], lane, wB0, wB1);
    const int perm = ((lane & 3) << 2) | ((lane >> 2) & 3) | (lane & 16);  // swap lane bits (0,1) <-> (2,3)
    float2 x0[8], x1[8];
    float2 *base = io + warp * 512;
    for (int c = 0; c < 8; c++) { x0[c] = base[c * 64 + lane]; x1[c] = base[c * 64 + 32 + lane]; }
    for (int r = 0; r < reps; r++) {
#pragma unroll
      for (int c = 0; c < 8; c++) gate_mma<MODE == 2>(x0[c], x1[c], wA0, wA1);
#pragma unroll
      for (int c = 0; c < 8; c++) {
        x0[c].x = __shfl_sync(0xffffffffu, x0[c].x, perm); x0[c].y = __shfl_sync(0xffffffffu, x0[c].y, perm);
        x1[c].x = __shfl_sync(0xffffffffu, x1[c].x, perm); x1[c].y = __shfl_sync(0xffffffffu, x1[c].y, perm);
      }
#pragma unroll
      for (int c = 0; c < 8; c++) gate_mma<MODE == 2>(x0[c], x1[c], wB0, wB1);
#pragma unroll
      for (int c = 0; c < 8; c++) {  // back to the gate-A layout (the permutation is an involution)
        x0[c].x = __shfl_sync(0xffffffffu, x0[c].x, perm); x0[c].y = __shfl_sync(0xffffffffu, x0[c].y, perm);
        x1[c].x = __shfl_sync(0xffffffffu, x1[c].x, perm); x1[c].y = __shfl_sync(0xffffffffu, x1[c].y, perm);
      }
    }
    for (int c = 0; c < 8; c++) { base[c * 64 + lane] = x0[c]; base[c * 64 + 32 + lane] = x1[c]; }
  }
}

// host reference of one round on 64-amplitude chunks (MMA layout) or 16-amplitude blocks (FFMA layout)
static void host_round(std::vector<cd> &v, const cd *MA, const cd *MB, bool mma_layout) {
  const size_t n = v.size();
  if (mma_layout) {
    for (size_t base = 0; base < n; base += 64)
      for (int gte = 0; gte < 2; gte++) {
        const cd *M = gte ? MB : MA;
        const int p0 = gte ? 2 : 0;
        for (int j = 0; j < 64; j++) {
          if ((j >> p0) & 3) continue;
          cd x[4], y[4];
          for (int t = 0; t < 4; t++) x[t] = v[base + j + (t << p0)];
          for (int i = 0; i < 4; i++) { y[i] = 0; for (int t = 0; t < 4; t++) y[i] += M[i * 4 + t] * x[t]; }
          for (int t = 0; t < 4; t++) v[base + j + (t << p0)] = y[t];
        }
      }
  } else {
    for (size_t base = 0; base < n; base += 16)
      for (int gte = 0; gte < 2; gte++) {
        const cd *M = gte ? MB : MA;
        for (int o = 0; o < 4; o++) {
          int idx[4];
          for (int t = 0; t < 4; t++) idx[t] = gte == 0 ? (o * 4 + t) : (o + 4 * t);
          cd x[4], y[4];
          for (int t = 0; t < 4; t++) x[t] = v[base + idx[t]];
          for (int i = 0; i < 4; i++) { y[i] = 0; for (int t = 0; t < 4; t++) y[i] += M[i * 4 + t] * x[t]; }
          for (int t = 0; t < 4; t++) v[base + idx[t]] = y[t];
        }
      }
  }
}

int main() {
  srand(3);
  // two random 4x4 unitaries (Gram-Schmidt)
  cd M[2][16];
  for (int g = 0; g < 2; g++) {
    cd a[4][4];
    for (auto &row : a) for (auto &z : row) z = cd(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    for (int i = 0; i < 4; i++) {
      for (int k = 0; k < i; k++) {
        cd d = 0;
        for (int j = 0; j < 4; j++) d += std::conj(a[k][j]) * a[i][j];
        for (int j = 0; j < 4; j++) a[i][j] -= d * a[k][j];
      }
      double nn = 0;
      for (int j = 0; j < 4; j++) nn += std::norm(a[i][j]);
      for (int j = 0; j < 4; j++) a[i][j] /= std::sqrt(nn);
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[g][i * 4 + j] = a[i][j];
  }
  Mats p;
  for (int g = 0; g < 2; g++) for (int i = 0; i < 16; i++) p.m[g][i] = make_float2((float)M[g][i].real(), (float)M[g][i].imag());
  const int blocks = 148 * 8, threads = 256;
  const size_t namp = (size_t)blocks * threads / 32 * 512;
  std::vector<cd> h(namp);
  double nrm = 0;
  for (auto &z : h) { z = cd(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5); nrm += std::norm(z); }
  for (auto &z : h) z /= std::sqrt(nrm / namp) ;  // amplitudes of unit RMS
  std::vector<float2> hf(namp), out(namp);
  for (size_t i = 0; i < namp; i++) hf[i] = make_float2((float)h[i].real(), (float)h[i].imag());
  float2 *d;
  cudaMalloc(&d, namp * 8);
  const int acc_reps = 20;  // 40 gates deep: accuracy after a realistic number of gates per amplitude
  const char *names[3] = {"FFMA register blocks", "TF32 (single pass)", "3xTF32"};
  for (int mode = 0; mode < 3; mode++) {
    cudaMemcpy(d, hf.data(), namp * 8, cudaMemcpyHostToDevice);
    if (mode == 0) round_kernel<0><<<blocks, threads>>>(d, p, acc_reps);
    if (mode == 1) round_kernel<1><<<blocks, threads>>>(d, p, acc_reps);
    if (mode == 2) round_kernel<2><<<blocks, threads>>>(d, p, acc_reps);
    cudaMemcpy(out.data(), d, namp * 8, cudaMemcpyDeviceToHost);
    std::vector<cd> ref(h.begin(), h.begin() + 4096);
    for (int r = 0; r < acc_reps; r++) host_round(ref, M[0], M[1], mode != 0);
    double err = 0, ov_re = 0, ov_im = 0, n1 = 0, n2 = 0;
    for (int i = 0; i < 4096; i++) {
      const cd g(out[i].x, out[i].y);
      err = fmax(err, std::abs(g - ref[i]));
      const cd o = std::conj(ref[i]) * g;
      ov_re += o.real(); ov_im += o.imag(); n1 += std::norm(ref[i]); n2 += std::norm(g);
    }
    const double fid_gap = fabs(1.0 - std::sqrt(ov_re * ov_re + ov_im * ov_im) / std::sqrt(n1 * n2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 2000;
    float ms = 0;
    for (int it = 0; it < 2; it++) {
      cudaEventRecord(e0);
      if (mode == 0) round_kernel<0><<<blocks, threads>>>(d, p, reps);
      if (mode == 1) round_kernel<1><<<blocks, threads>>>(d, p, reps);
      if (mode == 2) round_kernel<2><<<blocks, threads>>>(d, p, reps);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double gate_amps = (double)namp * reps * 2;  // amplitude-gate applications
    printf("%-22s after %d gates: max |err| %.2e (unit-RMS amplitudes), 1 - |<ref|psi>| = %.2e;  %8.2f ms  %6.1f G amp-gates/s  = %.1f FP32 TFLOP/s equivalent (%s)\n",
           names[mode], 2 * acc_reps, err, fid_gap, ms, gate_amps / ms / 1e6, gate_amps * 32 / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  }
  printf("HBM view: a pass at 6.55 TB/s moves 409 G float2 amplitudes/s (read + write): g gates per pass need 409 * g G amp-gates/s\n");
  return 0;
}
