// b200sv reductions (sm_100a): norm, Kraus norm, marginal probabilities, Pauli
// expectation values, inner product and the non-destructive sampler.
//
// All reductions are single-pass over the state with FP64 accumulation:
// per-thread partial -> warp shuffle tree -> one double per block, then one
// tiny finishing kernel that sums the block partials in a FIXED order, so the
// result is deterministic run to run (the reference needs log_1024(N) relaunches
// plus a sync per outcome, chunk_container.hpp:481-649,1068-1070).
#include "common.cuh"

namespace b200sv {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) wsum[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x + 31) / 32 ? wsum[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// finish: out[s*nout + o] = sum_b partial[(s*nblocks + b)*nout + o]
__global__ void finish_kernel(const double *__restrict__ partial, int nblocks, int nout, double *__restrict__ out) {
  const int s = blockIdx.x;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    double acc = 0;
    for (int b = 0; b < nblocks; b++) acc += partial[((size_t)s * nblocks + b) * nout + o];
    out[(size_t)s * nout + o] = acc;
  }
}

static int blocks_per_state(const State &s, uint64_t items_per_state) {
  uint64_t want = (items_per_state + kRedThreads * 4 - 1) / (kRedThreads * 4);
  uint64_t cap = std::max<uint64_t>(1, (uint64_t)s.num_sms * 8 / (uint64_t)s.nstates);
  return (int)std::max<uint64_t>(1, std::min(want, cap));
}

// runs `kernel-launch lambda` that fills partial[nstates][nblocks][nout], finishes and copies to host
template <typename F>
static void run_reduction(State &s, int nblocks, int nout, double *out_host, F launch) {
  const size_t npart = (size_t)s.nstates * nblocks * nout, nres = (size_t)s.nstates * nout;
  double *scratch = (double *)s.ensure_scratch((npart + nres) * sizeof(double) + 256);
  double *partial = scratch, *res = scratch + npart;
  launch(partial);
  B200_CUDA(cudaGetLastError());
  finish_kernel<<<(unsigned)s.nstates, 128, 0, s.stream>>>(partial, nblocks, nout, res);
  B200_CUDA(cudaGetLastError());
  double *pin = (double *)s.ensure_pinned(nres * sizeof(double));
  B200_CUDA(cudaMemcpyAsync(pin, res, nres * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(out_host, pin, nres * sizeof(double));
}

// ------------------------------------------------------------------ norm / inner product
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
norm_kernel(const cx<T> *__restrict__ psi, int nq, double *__restrict__ partial) {
  const uint64_t n = 1ull << nq;
  const cx<T> *st = psi + ((uint64_t)blockIdx.y << nq);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const cx<T> v0 = st[i], v1 = st[i + stride], v2 = st[i + 2 * stride], v3 = st[i + 3 * stride];
    a0 += (double)(v0.x * v0.x + v0.y * v0.y);
    a1 += (double)(v1.x * v1.x + v1.y * v1.y);
    a2 += (double)(v2.x * v2.x + v2.y * v2.y);
    a3 += (double)(v3.x * v3.x + v3.y * v3.y);
  }
  for (; i < n; i += stride) {
    const cx<T> v = st[i];
    a0 += (double)(v.x * v.x + v.y * v.y);
  }
  const double tot = block_sum((a0 + a1) + (a2 + a3));
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}
void reduce_norm(State &s, double *out) {
  NvtxRange nvtx("b200sv norm");
  const int nb = blocks_per_state(s, s.amps_per_state());
  run_reduction(s, nb, 1, out, [&](double *partial) {
    dim3 grid(nb, (unsigned)s.nstates);
    if (s.precision == B200SV_F64) norm_kernel<double><<<grid, kRedThreads, 0, s.stream>>>((const double2 *)s.data, s.nq, partial);
    else norm_kernel<float><<<grid, kRedThreads, 0, s.stream>>>((const float2 *)s.data, s.nq, partial);
  });
}

// z = data * conj(checkpoint)  (qubitvector.hpp:1030-1041)
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
inner_kernel(const cx<T> *__restrict__ psi, const cx<T> *__restrict__ chk, int nq, double *__restrict__ partial) {
  const uint64_t n = 1ull << nq;
  const uint64_t sb = (uint64_t)blockIdx.y << nq;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double re = 0, im = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const cx<T> d = psi[sb + i], c = chk[sb + i];
    re += (double)(d.x * c.x + d.y * c.y);
    im += (double)(d.y * c.x - d.x * c.y);
  }
  re = block_sum(re);
  im = block_sum(im);
  if (threadIdx.x == 0) {
    const size_t o = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    partial[o] = re;
    partial[o + 1] = im;
  }
}
void reduce_inner_product(State &s, const void *other, double *re, double *im) {
  const int nb = blocks_per_state(s, s.amps_per_state());
  std::vector<double> out(2 * s.nstates);
  run_reduction(s, nb, 2, out.data(), [&](double *partial) {
    dim3 grid(nb, (unsigned)s.nstates);
    if (s.precision == B200SV_F64)
      inner_kernel<double><<<grid, kRedThreads, 0, s.stream>>>((const double2 *)s.data, (const double2 *)other, s.nq, partial);
    else
      inner_kernel<float><<<grid, kRedThreads, 0, s.stream>>>((const float2 *)s.data, (const float2 *)other, s.nq, partial);
  });
  for (int64_t i = 0; i < s.nstates; i++) { re[i] = out[2 * i]; im[i] = out[2 * i + 1]; }
}

// ------------------------------------------------------------------ || M psi ||^2  (Kraus probability)
struct NormMatParams {
  uint64_t off_bits[kMaxRegQubits];
  uint64_t groups_per_state;
  InsertList ins;
  int k;
};
template <typename T, int K>
__global__ void __launch_bounds__(kRedThreads)
norm_matrix_kernel(const cx<T> *__restrict__ psi, const cx<T> *__restrict__ mat /* row major */, int nq,
                   const __grid_constant__ NormMatParams p, double *__restrict__ partial) {
  constexpr int DIM = 1 << K;
  __shared__ cx<T> m[DIM * DIM];
  for (int i = threadIdx.x; i < DIM * DIM; i += blockDim.x) m[i] = mat[i];
  __syncthreads();
  const cx<T> *st = psi + ((uint64_t)blockIdx.y << nq);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.groups_per_state; g += stride) {
    const uint64_t base = insert_zeros(g, p.ins);
    cx<T> in[DIM];
#pragma unroll
    for (int e = 0; e < DIM; e++) {
      uint64_t o = 0;
#pragma unroll
      for (int b = 0; b < K; b++)
        if ((e >> b) & 1) o |= p.off_bits[b];
      in[e] = st[base + o];
    }
#pragma unroll 4
    for (int i = 0; i < DIM; i++) {
      cx<T> v = mk<T>(0, 0);
#pragma unroll
      for (int j = 0; j < DIM; j++) cfma(v, m[i * DIM + j], in[j]);
      acc += (double)(v.x * v.x + v.y * v.y);
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
void reduce_norm_matrix(State &s, const int *qubits, int k, const double *mat, double *out) {
  NvtxRange nvtx("b200sv norm(qubits, mat)");
  if (k > kMaxRegQubits) throw Error("norm(qubits, mat): more than 5 qubits is not supported");
  const int dim = 1 << k;
  const size_t bytes = (size_t)dim * dim * s.amp_bytes();
  // row-major matrix in device scratch *after* the reduction area (offset 1 MiB region is ours)
  void *hm = s.ensure_pinned(bytes);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) {
      const double re = mat[2 * (i + dim * j)], im = mat[2 * (i + dim * j) + 1];
      if (s.precision == B200SV_F64) ((double2 *)hm)[i * dim + j] = mk<double>(re, im);
      else ((float2 *)hm)[i * dim + j] = mk<float>((float)re, (float)im);
    }
  NormMatParams p;
  p.k = k;
  std::vector<int> all(qubits, qubits + k);
  for (int b = 0; b < k; b++) p.off_bits[b] = 1ull << qubits[b];
  std::sort(all.begin(), all.end());
  p.ins.n = k;
  for (int i = 0; i < k; i++) p.ins.pos[i] = (uint8_t)all[i];
  p.groups_per_state = s.amps_per_state() >> k;
  const int nb = blocks_per_state(s, p.groups_per_state);
  const size_t red_bytes = ((size_t)s.nstates * (nb + 1)) * sizeof(double) + 256;
  char *scr = (char *)s.ensure_scratch(red_bytes + bytes + 256);
  void *dm = scr + ((red_bytes + 255) & ~(size_t)255);
  B200_CUDA(cudaMemcpyAsync(dm, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  run_reduction(s, nb, 1, out, [&](double *partial) {
    dim3 grid(nb, (unsigned)s.nstates);
#define CASE(K)                                                                                                    \
  case K:                                                                                                          \
    if (s.precision == B200SV_F64)                                                                                 \
      norm_matrix_kernel<double, K><<<grid, kRedThreads, 0, s.stream>>>((const double2 *)s.data, (const double2 *)dm, s.nq, p, partial); \
    else                                                                                                           \
      norm_matrix_kernel<float, K><<<grid, kRedThreads, 0, s.stream>>>((const float2 *)s.data, (const float2 *)dm, s.nq, p, partial);   \
    break;
    switch (k) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) }
#undef CASE
  });
}

// ------------------------------------------------------------------ Pauli expectation value
// MODE 0: Z-only (one amplitude per item)   MODE 1: X/Y present (pairs)   MODE 2: pair chunk
template <typename T, int MODE>
__global__ void __launch_bounds__(kRedThreads)
expval_kernel(const cx<T> *__restrict__ psi, const cx<T> *__restrict__ pair, int nq, uint64_t items, uint64_t x_mask,
              uint64_t z_mask, int x_max, cx<T> phase, uint64_t zc, uint64_t zcp, double *__restrict__ partial) {
  const cx<T> *st = psi + ((uint64_t)blockIdx.y << nq);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
    if (MODE == 0) {
      const cx<T> d = st[i];
      const cx<T> t = cmul(phase, d);
      double v = (double)(t.x * d.x + t.y * d.y);  // Re(phase * d * conj(d))
      if (__popcll(i & z_mask) & 1) v = -v;
      acc += v;
    } else {
      uint64_t i0, i1;
      cx<T> d0, d1;
      if (MODE == 1) {
        i0 = insert_zero(i, x_max);
        i1 = i0 ^ x_mask;
        d0 = st[i0];
        d1 = st[i1];
      } else {
        i0 = i;
        i1 = i ^ x_mask;
        d0 = st[i0];
        d1 = pair[i1];
      }
      cx<T> t = cmul(phase, d1);
      double v0 = (double)(t.x * d0.x + t.y * d0.y);  // Re(phase * d1 * conj(d0))
      t = cmul(phase, d0);
      double v1 = (double)(t.x * d1.x + t.y * d1.y);  // Re(phase * d0 * conj(d1))
      if ((__popcll(i0 & z_mask) + zc) & 1) v0 = -v0;
      if ((__popcll(i1 & z_mask) + zcp) & 1) v1 = -v1;
      acc += v0 + v1;
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
void reduce_expval_pauli(State &s, uint64_t x_mask, uint64_t z_mask, int x_max, double pre, double pim,
                         const void *pair, uint64_t zc, uint64_t zcp, double *out) {
  const int mode = pair ? 2 : (x_mask ? 1 : 0);
  const uint64_t items = mode == 1 ? s.amps_per_state() >> 1 : s.amps_per_state();
  const int nb = blocks_per_state(s, items);
  run_reduction(s, nb, 1, out, [&](double *partial) {
    dim3 grid(nb, (unsigned)s.nstates);
#define LAUNCH(T, M)                                                                                             \
  expval_kernel<T, M><<<grid, kRedThreads, 0, s.stream>>>((const cx<T> *)s.data, (const cx<T> *)pair, s.nq, items, \
                                                          x_mask, z_mask, x_max, mk<T>((T)pre, (T)pim), zc, zcp, partial)
    if (s.precision == B200SV_F64) {
      if (mode == 0) LAUNCH(double, 0); else if (mode == 1) LAUNCH(double, 1); else LAUNCH(double, 2);
    } else {
      if (mode == 0) LAUNCH(float, 0); else if (mode == 1) LAUNCH(float, 1); else LAUNCH(float, 2);
    }
#undef LAUNCH
  });
}

// ------------------------------------------------------------------ marginal probabilities, all 2^k outcomes in ONE pass
// Thread tid of a block always reads amplitudes whose low LB index bits equal
// tid, so the measured bits below LB are fixed per thread; the measured bits
// at or above LB are fixed per block (blockIdx selects them), and the block
// loops over (a slice of) the remaining free high bits accumulating ONE
// register per thread.  No atomics, fixed summation order -> deterministic.
struct ProbParams {
  int nq, LB, k;
  int n_mhigh, n_mlow;
  uint8_t mhigh_pos[kMaxInsert];   // sorted measured positions >= LB, relative to LB
  uint8_t mhigh_bit[kMaxInsert];   // outcome bit index j of that qubit
  uint8_t mlow_pos[8];             // sorted measured positions < LB
  uint8_t mlow_bit[8];
  uint64_t free_count;             // 2^(free high bits)
  uint32_t slices;                 // S: blocks sharing one H
};
template <typename T>
__global__ void __launch_bounds__(256)
prob_kernel(const cx<T> *__restrict__ psi, const __grid_constant__ ProbParams p, double *__restrict__ partial /*[state][slice][2^k]*/) {
  __shared__ double vals[256];
  const uint32_t slice = blockIdx.x % p.slices;
  const uint64_t H = blockIdx.x / p.slices;
  const cx<T> *st = psi + ((uint64_t)blockIdx.y << p.nq);
  // fixed high part of the address and of the outcome
  uint64_t Haddr = 0, Hm = 0;
  for (int i = 0; i < p.n_mhigh; i++)
    if ((H >> i) & 1) {
      Haddr |= 1ull << (p.mhigh_pos[i] + p.LB);
      Hm |= 1ull << p.mhigh_bit[i];
    }
  InsertList ins;
  ins.n = p.n_mhigh;
  for (int i = 0; i < p.n_mhigh; i++) ins.pos[i] = p.mhigh_pos[i];
  double a0 = 0, a1 = 0;
  const uint64_t per = p.free_count / p.slices;
  const uint64_t u_begin = per * slice, u_end = u_begin + per;
  uint64_t u = u_begin;
  for (; u + 1 < u_end; u += 2) {
    const uint64_t i0 = (insert_zeros(u, ins) << p.LB) | Haddr | threadIdx.x;
    const uint64_t i1 = (insert_zeros(u + 1, ins) << p.LB) | Haddr | threadIdx.x;
    const cx<T> v0 = st[i0], v1 = st[i1];
    a0 += (double)(v0.x * v0.x + v0.y * v0.y);
    a1 += (double)(v1.x * v1.x + v1.y * v1.y);
  }
  for (; u < u_end; u++) {
    const cx<T> v0 = st[(insert_zeros(u, ins) << p.LB) | Haddr | threadIdx.x];
    a0 += (double)(v0.x * v0.x + v0.y * v0.y);
  }
  vals[threadIdx.x] = a0 + a1;
  __syncthreads();
  // combine threads that share the same measured-low pattern, in increasing tid order
  const int nlow = 1 << p.n_mlow;
  if ((int)threadIdx.x < nlow) {
    const int c = threadIdx.x;
    uint32_t fixed = 0;   // tid bits forced by pattern c
    uint32_t lowmask = 0;
    uint64_t m = Hm;
    for (int i = 0; i < p.n_mlow; i++) {
      lowmask |= 1u << p.mlow_pos[i];
      if ((c >> i) & 1) {
        fixed |= 1u << p.mlow_pos[i];
        m |= 1ull << p.mlow_bit[i];
      }
    }
    double acc = 0;
    for (uint32_t t = 0; t < blockDim.x; t++)
      if ((t & lowmask) == fixed) acc += vals[t];
    partial[(((size_t)blockIdx.y * p.slices + slice) << p.k) + m] = acc;
  }
}
void reduce_probabilities(State &s, const int *qubits, int k, double *out) {
  NvtxRange nvtx("b200sv probabilities");
  if (k > 26) throw Error("probabilities(qubits): more than 26 measured qubits per call is not supported");
  ProbParams p;
  p.nq = s.nq; p.k = k;
  p.LB = std::min(8, s.nq);
  std::vector<std::pair<int, int>> lo, hi;  // (position, outcome bit)
  for (int j = 0; j < k; j++) (qubits[j] < p.LB ? lo : hi).push_back({qubits[j], j});
  std::sort(lo.begin(), lo.end());
  std::sort(hi.begin(), hi.end());
  p.n_mlow = (int)lo.size(); p.n_mhigh = (int)hi.size();
  for (size_t i = 0; i < lo.size(); i++) { p.mlow_pos[i] = (uint8_t)lo[i].first; p.mlow_bit[i] = (uint8_t)lo[i].second; }
  for (size_t i = 0; i < hi.size(); i++) { p.mhigh_pos[i] = (uint8_t)(hi[i].first - p.LB); p.mhigh_bit[i] = (uint8_t)hi[i].second; }
  const int free_bits = s.nq - p.LB - p.n_mhigh;
  p.free_count = 1ull << free_bits;
  // slices: enough blocks to fill the machine, at most free_count
  uint64_t hblocks = 1ull << p.n_mhigh;
  uint64_t want = ((uint64_t)s.num_sms * 8 + hblocks * s.nstates - 1) / (hblocks * s.nstates);
  uint64_t S = 1;
  while (S < want && S < p.free_count) S <<= 1;
  p.slices = (uint32_t)S;
  const int threads = 1 << p.LB;
  const int nout = 1 << k;
  const size_t npart = (size_t)s.nstates * S * nout, nres = (size_t)s.nstates * nout;
  double *scratch = (double *)s.ensure_scratch((npart + nres) * sizeof(double) + 256);
  double *partial = scratch, *res = scratch + npart;
  dim3 grid((unsigned)(hblocks * S), (unsigned)s.nstates);
  if (s.precision == B200SV_F64) prob_kernel<double><<<grid, threads, 0, s.stream>>>((const double2 *)s.data, p, partial);
  else prob_kernel<float><<<grid, threads, 0, s.stream>>>((const float2 *)s.data, p, partial);
  B200_CUDA(cudaGetLastError());
  finish_kernel<<<(unsigned)s.nstates, 256, 0, s.stream>>>(partial, (int)S, nout, res);
  B200_CUDA(cudaGetLastError());
  double *pin = (double *)s.ensure_pinned(nres * sizeof(double));
  B200_CUDA(cudaMemcpyAsync(pin, res, nres * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(out, pin, nres * sizeof(double));
}

// ------------------------------------------------------------------ density-matrix reductions
// The state is vec(rho) of a 2^m x 2^m matrix (index = row + col * 2^m, densitymatrix.hpp:292-343).  Pauli
// expectation values and marginal probabilities only touch one 2^m-entry line rho[i ^ x, i]
// (densitymatrix.hpp:470-520, :590-593; GPU twins density_expval_pauli_func / probability functors,
// densitymatrix_thrust.hpp:1011-1188): one strided pass on the device, fixed-order finish, no line on the host.
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
dm_expval_kernel(const cx<T> *__restrict__ rho, int m, uint64_t x_mask, uint64_t z_mask, double pre, double pim,
                 double *__restrict__ partial) {
  const uint64_t n = 1ull << m, stride = (uint64_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const cx<T> v = rho[(i ^ x_mask) + (i << m)];
    double t = x_mask ? pre * (double)v.x - pim * (double)v.y : (double)v.x;  // Re(phase * rho[i ^ x, i])
    if (__popcll(i & z_mask) & 1) t = -t;
    acc += t;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
void reduce_dm_expval(State &s, int m, uint64_t x_mask, uint64_t z_mask, double pre, double pim, double *out) {
  NvtxRange nvtx("b200sv density-matrix expval_pauli");
  const int nb = (int)std::max<uint64_t>(1, std::min<uint64_t>(((1ull << m) + kRedThreads - 1) / kRedThreads, (uint64_t)s.num_sms * 4));
  run_reduction(s, nb, 1, out, [&](double *partial) {
    if (s.precision == B200SV_F64) dm_expval_kernel<double><<<nb, kRedThreads, 0, s.stream>>>((const double2 *)s.data, m, x_mask, z_mask, pre, pim, partial);
    else dm_expval_kernel<float><<<nb, kRedThreads, 0, s.stream>>>((const float2 *)s.data, m, x_mask, z_mask, pre, pim, partial);
  });
}
// marginal probabilities of k measured qubits from the diagonal: every block owns a contiguous slice of the
// diagonal and 2^k bins in shared memory; the bins of all blocks are summed in a fixed order
struct DmProbParams {
  uint8_t q[16];
  int k, m;
};
template <typename T>
__global__ void __launch_bounds__(kRedThreads)
dm_prob_kernel(const cx<T> *__restrict__ rho, const __grid_constant__ DmProbParams p, double *__restrict__ partial) {
  extern __shared__ double bins[];
  const int nb = 1 << p.k;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) bins[b] = 0.0;
  __syncthreads();
  const uint64_t n = 1ull << p.m, per = (n + gridDim.x - 1) / gridDim.x;
  const uint64_t i0 = per * blockIdx.x, i1 = i0 + per < n ? i0 + per : n;
  for (uint64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    uint32_t o = 0;
    for (int j = 0; j < p.k; j++) o |= (uint32_t)((i >> p.q[j]) & 1ull) << j;
    atomicAdd(&bins[o], (double)rho[i + (i << p.m)].x);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += blockDim.x) partial[(size_t)blockIdx.x * nb + b] = bins[b];
}
void reduce_dm_probabilities(State &s, int m, const int *qubits, int k, double *out) {
  NvtxRange nvtx("b200sv density-matrix probabilities");
  if (k > 12) throw Error("density-matrix probabilities: more than 12 measured qubits at once");
  DmProbParams p;
  p.k = k; p.m = m;
  for (int j = 0; j < k; j++) p.q[j] = (uint8_t)qubits[j];
  const int nb = (int)std::max<uint64_t>(1, std::min<uint64_t>(((1ull << m) + 4 * kRedThreads - 1) / (4 * kRedThreads), (uint64_t)s.num_sms * 2));
  run_reduction(s, nb, 1 << k, out, [&](double *partial) {
    const size_t smem = sizeof(double) << k;
    if (s.precision == B200SV_F64) dm_prob_kernel<double><<<nb, kRedThreads, smem, s.stream>>>((const double2 *)s.data, p, partial);
    else dm_prob_kernel<float><<<nb, kRedThreads, smem, s.stream>>>((const float2 *)s.data, p, partial);
  });
}

// ------------------------------------------------------------------ sampler (non-destructive, no 2^n temporary)
// level 1: sums of contiguous blocks of 2^B amplitudes (one pass over the state)
// level 2: exclusive scan of the 2^(n-B) block sums (one CTA, fixed order)
// level 3: one warp per shot: binary search for the block, then a sequential
//          warp-scan inside that block only.  Matches qubitvector.hpp:2149-2228:
//          sample = first index with rnd < cumulative probability, clamped to END-1.
constexpr int kSampleB = 12;

template <typename T>
__global__ void __launch_bounds__(256)
block_sums_kernel(const cx<T> *__restrict__ psi, int B, uint64_t nblocks, double *__restrict__ bsum) {
  // blockIdx.y = state of a batched container (its amplitudes and its nblocks + 1 sums are contiguous)
  psi += ((uint64_t)blockIdx.y * nblocks) << B;
  bsum += (uint64_t)blockIdx.y * (nblocks + 1);
  for (uint64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const cx<T> *blk = psi + (b << B);
    double a = 0;
    for (uint32_t i = threadIdx.x; i < (1u << B); i += blockDim.x) {
      const cx<T> v = blk[i];
      a += (double)(v.x * v.x + v.y * v.y);
    }
    a = block_sum(a);
    if (threadIdx.x == 0) bsum[b] = a;
    __syncthreads();
  }
}
// in-place exclusive scan by one CTA of 1024 threads; writes total to bsum[n]
__global__ void __launch_bounds__(1024) scan_kernel(double *__restrict__ bsum, uint64_t n) {
  __shared__ double tsum[1024];
  bsum += (uint64_t)blockIdx.x * (n + 1);  // one CTA per state
  const uint64_t per = (n + 1023) / 1024;
  const uint64_t b0 = per * threadIdx.x, b1 = b0 + per < n ? b0 + per : n;
  double a = 0;
  for (uint64_t i = b0; i < b1; i++) a += bsum[i];
  tsum[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0;
    for (int t = 0; t < 1024; t++) { const double v = tsum[t]; tsum[t] = run; run += v; }
    bsum[n] = run;
  }
  __syncthreads();
  double run = tsum[threadIdx.x];
  for (uint64_t i = b0; i < b1; i++) { const double v = bsum[i]; bsum[i] = run; run += v; }
}
template <typename T>
__global__ void __launch_bounds__(256)
sample_kernel(const cx<T> *__restrict__ psi, int B, uint64_t nblocks, const double *__restrict__ excl,
              const double *__restrict__ rnds, int64_t shots, uint64_t *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t shot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (shot >= shots) return;
  psi += ((uint64_t)blockIdx.y * nblocks) << B;  // blockIdx.y = state; rnds / out are [state][shot]
  excl += (uint64_t)blockIdx.y * (nblocks + 1);
  rnds += (int64_t)blockIdx.y * shots;
  out += (int64_t)blockIdx.y * shots;
  const double rnd = rnds[shot];
  // last block whose exclusive prefix is <= rnd  (== first block with rnd < inclusive prefix)
  uint64_t lo = 0, hi = nblocks;  // invariant: excl[lo] <= rnd (excl[0] = 0), answer in [lo, hi)
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (excl[mid] <= rnd) lo = mid; else hi = mid;
  }
  const uint64_t END = nblocks << B;
  uint64_t sample = END - 1;
  double run = excl[lo];
  bool found = false;
  // scan forward from block lo (normally terminates inside it; rounding may spill into the next)
  for (uint64_t base = lo << B; base < END && !found; base += 32) {
    double pr = 0;
    if (base + lane < END) {
      const cx<T> v = psi[base + lane];
      pr = (double)(v.x * v.x + v.y * v.y);
    }
    // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, pr, o);
      if (lane >= o) pr += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, rnd < run + pr);
    if (hit) {
      sample = base + (__ffs(hit) - 1);
      found = true;
    }
    run += __shfl_sync(0xffffffffu, pr, 31);
  }
  if (sample > END - 1) sample = END - 1;
  if (lane == 0) out[shot] = sample;
}

void sample_measure(State &s, const double *rnds, int64_t shots, uint64_t *out) {
  NvtxRange nvtx("b200sv sample_measure");
  if (shots <= 0) return;
  const int B = std::min(kSampleB, s.nq);
  const uint64_t nblocks = s.amps_per_state() >> B;
  const size_t S = (size_t)s.nstates, tot = S * (size_t)shots;
  // scratch layout: [bsum (nblocks+1) per state] [rnds] [out]; all states ride on one launch of each kernel and the
  // draws / samples cross PCIe in one transfer each (batched measure of 10k shots: 3 launches, not 30k)
  const size_t off_r = (S * (nblocks + 1) * sizeof(double) + 255) & ~(size_t)255;
  const size_t off_o = off_r + ((tot * sizeof(double) + 255) & ~(size_t)255);
  const size_t total = off_o + tot * sizeof(uint64_t);
  char *scr = (char *)s.ensure_scratch(total);
  char *pin = (char *)s.ensure_pinned(tot * 8);
  double *bsum = (double *)scr;
  double *d_r = (double *)(scr + off_r);
  uint64_t *d_o = (uint64_t *)(scr + off_o);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(pin, rnds, tot * 8);
  B200_CUDA(cudaMemcpyAsync(d_r, pin, tot * 8, cudaMemcpyHostToDevice, s.stream));
  const int per_state = (int)std::max<uint64_t>(1, (uint64_t)s.num_sms * 8 / S);
  dim3 g1((unsigned)std::min<uint64_t>(nblocks, (uint64_t)per_state), (unsigned)S);
  dim3 g3((unsigned)((shots * 32 + 255) / 256), (unsigned)S);
  if (s.precision == B200SV_F64) {
    const double2 *psi = (const double2 *)s.data;
    block_sums_kernel<double><<<g1, 256, 0, s.stream>>>(psi, B, nblocks, bsum);
    scan_kernel<<<(unsigned)S, 1024, 0, s.stream>>>(bsum, nblocks);
    sample_kernel<double><<<g3, 256, 0, s.stream>>>(psi, B, nblocks, bsum, d_r, shots, d_o);
  } else {
    const float2 *psi = (const float2 *)s.data;
    block_sums_kernel<float><<<g1, 256, 0, s.stream>>>(psi, B, nblocks, bsum);
    scan_kernel<<<(unsigned)S, 1024, 0, s.stream>>>(bsum, nblocks);
    sample_kernel<float><<<g3, 256, 0, s.stream>>>(psi, B, nblocks, bsum, d_r, shots, d_o);
  }
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(pin, d_o, tot * 8, cudaMemcpyDeviceToHost, s.stream));
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(out, pin, tot * 8);
}

}  // namespace b200sv
