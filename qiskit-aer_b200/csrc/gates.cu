// b200sv gate-application kernels (sm_100a).
//
// Every kernel is a streaming pass over the amplitudes it must touch:
//   * one thread owns one "group" (the 2^k amplitudes coupled by a k-qubit
//     gate); the 2^k loads are each a 128-bit access (one complex<double>)
//     and are issued back to back before any math, so a warp keeps
//     32 * 2^k independent 16-byte requests in flight;
//   * controls are folded into index generation: a gate with c controls
//     launches 2^(n-c-k) groups, never 2^(n-k) masked threads (the reference
//     launches and masks, thrust_kernels.hpp:1190);
//   * gate matrices travel as __grid_constant__ kernel parameters (constant
//     bank, no H2D copy, no per-gate cudaMemcpy as in
//     device_chunk_container.hpp:566-594) and feed DFMA directly.
// Index algebra: base = insert_zeros(g, sorted(targets U controls)) | ctrl_mask,
// element e at base + off[e]  (indexes.hpp:212-250).
#include <complex>

#include "common.cuh"

namespace b200sv {

static inline int grid_for(const State &s, uint64_t work_items, int threads, int per_sm) {
  uint64_t blocks = (work_items + threads - 1) / threads;
  uint64_t cap = (uint64_t)s.num_sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------ dense, registers
template <typename T, int K> struct DenseParams {
  cx<T> m[1 << (2 * K)];     // row major: m[i*DIM + j] = M[i][j]
  uint64_t off[1 << K];      // amplitude offset of matrix index e
  uint64_t ctrl_mask;
  uint64_t ngroups;
  InsertList ins;
};

template <typename T, int K>
__global__ void __launch_bounds__(K >= 5 ? 128 : 256, (K == 3 || K == 4) ? 2 : 1)
dense_kernel(cx<T> *__restrict__ psi, const __grid_constant__ DenseParams<T, K> p) {
  constexpr int DIM = 1 << K;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
    const uint64_t base = insert_zeros(g, p.ins) | p.ctrl_mask;
    cx<T> in[DIM];
#pragma unroll
    for (int e = 0; e < DIM; e++) in[e] = psi[base + p.off[e]];
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      cx<T> acc = mk<T>(0, 0);
#pragma unroll
      for (int j = 0; j < DIM; j++) cfma(acc, p.m[i * DIM + j], in[j]);
      psi[base + p.off[i]] = acc;
    }
  }
}

// Variant for blocks that contain global qubit 0 (the host makes it matrix bit 0): elements 2j and 2j + 1 of a group
// are neighbours in memory, so every access is one 256-bit load / store = a whole 32-byte sector per lane.  With
// 128-bit accesses a group whose targets all sit in the low bits costs every sector twice (two instructions, each
// using half of it): 0.50-0.58 of the HBM peak for K = 3, 4 before (profiles/r01_sweep_n30_v1.txt).
template <int K>
__global__ void __launch_bounds__(256, (K == 3 || K == 4) ? 2 : 1)
dense_pair0_kernel(double2 *__restrict__ psi, const __grid_constant__ DenseParams<double, K> p) {
  constexpr int DIM = 1 << K;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
    const uint64_t base = insert_zeros(g, p.ins) | p.ctrl_mask;
    double2 in[DIM];
#pragma unroll
    for (int e = 0; e < DIM; e += 2)
      asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                   : "=d"(in[e].x), "=d"(in[e].y), "=d"(in[e + 1].x), "=d"(in[e + 1].y)
                   : "l"(psi + base + p.off[e]));
#pragma unroll
    for (int i = 0; i < DIM; i += 2) {
      double2 a0 = mk<double>(0, 0), a1 = mk<double>(0, 0);
#pragma unroll
      for (int j = 0; j < DIM; j++) {
        cfma(a0, p.m[i * DIM + j], in[j]);
        cfma(a1, p.m[(i + 1) * DIM + j], in[j]);
      }
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(psi + base + p.off[i]), "d"(a0.x), "d"(a0.y), "d"(a1.x),
                   "d"(a1.y)
                   : "memory");
    }
  }
}

// ------------------------------------------------------------------ dense k = 5 on the FP64 tensor path (DMMA)
// A 5-qubit block is 128 DFMA per amplitude: FP64 bound, and in the register kernel above also issue bound (every
// complex matrix entry feeds one complex FMA per thread: one LDCU per two DFMA).  Here the block is a real 64x64
// matrix [[Re M, -Im M], [Im M, Re M]] applied to columns of 64 reals (re 0..31, im 0..31 of a group) with
// mma.sync.m8n8k4.f64: 256 FMA per instruction, the matrix fragment comes from shared memory with ONE LDS.64 per MMA.
// A warp takes 8 groups per step.  Input mapping: lane l holds the amplitudes 4i + (l & 3), i = 0..7, of group l >> 2
// -- re/im of one amplitude are exactly the B-fragment values of k-steps i and i + 8, so amplitudes are loaded as
// whole double2.  Output mapping (the m8n8 accumulator layout): lane l ends up with re (row blocks 0..3) and im (row
// blocks 4..7) of the amplitudes (l >> 2) + 8j of groups 2(l & 3) + {0,1}: again whole double2, stored straight from
// registers.  No shared-memory transposition of amplitudes at all.  (North star: "only large-k fused blocks ... on
// FP64 DMMA"; tools/micro/fp64_pipes.cu: DMMA and DFMA share one 37 TFLOP/s datapath, DMMA needs 8x fewer issue slots.)
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// PAIR: global qubit 0 is not a target, so the two groups a lane holds results for (2c, 2c + 1) are neighbours in
// memory: one 256-bit store per amplitude instead of two 16-byte pieces in separate instructions (full 32-byte sectors).
template <bool PAIR>
__global__ void __launch_bounds__(256, 3)
dense5_dmma_kernel(double2 *__restrict__ psi, const __grid_constant__ DenseParams<double, 5> p) {
  __shared__ double afrag[128 * 32];  // fragment f = rb * 16 + ks, lane l: R[8 rb + (l >> 2)][4 ks + (l & 3)]
  for (int e = threadIdx.x; e < 128 * 32; e += blockDim.x) {
    const int f = e >> 5, l = e & 31, o = 8 * (f >> 4) + (l >> 2), i = 4 * (f & 15) + (l & 3);
    const double2 m = p.m[(o & 31) * 32 + (i & 31)];
    afrag[e] = (o < 32) == (i < 32) ? m.x : (o < 32 ? -m.y : m.y);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint64_t in_off[8], out_off[4];
#pragma unroll
  for (int i = 0; i < 8; i++) in_off[i] = p.off[4 * i + (lane & 3)];
#pragma unroll
  for (int j = 0; j < 4; j++) out_off[j] = p.off[(lane >> 2) + 8 * j];
  const uint64_t nbatch = p.ngroups >> 3;
  const uint64_t wstride = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t gb = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gb < nbatch; gb += wstride) {
    const uint64_t base_in = insert_zeros(gb * 8 + (lane >> 2), p.ins);
    double xr[8], xi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const double2 v = psi[base_in + in_off[i]];
      xr[i] = v.x;
      xi[i] = v.y;
    }
    double acc[8][2];
#pragma unroll
    for (int rb = 0; rb < 8; rb++) acc[rb][0] = acc[rb][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 16; ks++) {
      const double b = ks < 8 ? xr[ks] : xi[ks - 8];
#pragma unroll
      for (int rb = 0; rb < 8; rb++) dmma_m8n8k4(acc[rb][0], acc[rb][1], afrag[(rb * 16 + ks) * 32 + lane], b);
    }
    // all lanes of the warp have consumed their inputs (mma.sync is warp-collective): safe to overwrite in place
    if (PAIR) {
      const uint64_t base_out = insert_zeros(gb * 8 + 2 * (lane & 3), p.ins);
#pragma unroll
      for (int j = 0; j < 4; j++)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(psi + base_out + out_off[j]), "d"(acc[j][0]),
                     "d"(acc[j + 4][0]), "d"(acc[j][1]), "d"(acc[j + 4][1])
                     : "memory");
    } else {
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const uint64_t base_out = insert_zeros(gb * 8 + 2 * (lane & 3) + g, p.ins);
#pragma unroll
        for (int j = 0; j < 4; j++) psi[base_out + out_off[j]] = make_double2(acc[j][g], acc[j + 4][g]);
      }
    }
  }
}

template <typename T, int K>
static void launch_dense_t(State &s, const int *targets, const int *controls, int nc, const double *mat) {
  constexpr int DIM = 1 << K;
  static thread_local DenseParams<T, K> p;  // large: keep off the stack; thread_local because Aer calls from OpenMP threads
  // matrix bit order: as given, except that global qubit 0 (if it is a target) becomes matrix bit 0 for the
  // 256-bit-access variant (double precision, K = 2..4): bit b0 <-> bit 0 of the matrix indices
  int tq[K];
  for (int b = 0; b < K; b++) tq[b] = targets[b];
  int b0 = -1;
  if constexpr (std::is_same<T, double>::value && K >= 2 && K <= 4) {
    for (int b = 0; b < K; b++)
      if (targets[b] == 0) b0 = b;
    if (((uintptr_t)s.data & 31) != 0) b0 = -1;
    if (b0 > 0) std::swap(tq[0], tq[b0]);
  }
  auto perm = [&](int i) {  // index in the caller's bit order of index i in ours
    if (b0 <= 0) return i;
    const int lo = i & 1, hi = (i >> b0) & 1;
    return (i & ~(1 | (1 << b0))) | (hi) | (lo << b0);
  };
  for (int i = 0; i < DIM; i++)
    for (int j = 0; j < DIM; j++) {
      const int si = perm(i), sj = perm(j);
      p.m[i * DIM + j] = mk<T>((T)mat[2 * (si + DIM * sj)], (T)mat[2 * (si + DIM * sj) + 1]);
    }
  for (int e = 0; e < DIM; e++) {
    uint64_t o = 0;
    for (int b = 0; b < K; b++)
      if ((e >> b) & 1) o |= 1ull << tq[b];
    p.off[e] = o;
  }
  std::vector<int> all(targets, targets + K);
  p.ctrl_mask = 0;
  for (int c = 0; c < nc; c++) {
    all.push_back(controls[c]);
    p.ctrl_mask |= 1ull << controls[c];
  }
  std::sort(all.begin(), all.end());
  p.ins.n = (int)all.size();
  for (size_t i = 0; i < all.size(); i++) p.ins.pos[i] = (uint8_t)all[i];
  p.ngroups = s.total_amps() >> (K + nc);
  if constexpr (K == 5 && std::is_same<T, double>::value) {
    static const int env_dmma = [] { const char *e = getenv("B200SV_DENSE5_DMMA"); return e ? atoi(e) : 1; }();
    // the kernel walks batches of 8 groups: batched containers whose group count is not a multiple of 8 (e.g. 10 states
    // of 5 qubits) take the register kernel; the 256-bit stores of the PAIR variant need a 32-byte aligned state
    if (env_dmma && nc == 0 && p.ngroups >= 8 && (p.ngroups & 7) == 0) {
      const int grid5 = (int)std::min<uint64_t>((p.ngroups / 8 + 7) / 8, (uint64_t)s.num_sms * 3);
      if (p.ins.pos[0] > 0 && ((uintptr_t)s.data & 31) == 0) dense5_dmma_kernel<true><<<grid5, 256, 0, s.stream>>>((double2 *)s.data, p);
      else dense5_dmma_kernel<false><<<grid5, 256, 0, s.stream>>>((double2 *)s.data, p);
      B200_CUDA(cudaGetLastError());
      return;
    }
  }
  const int threads = K >= 5 ? 128 : 256;
  const int grid = grid_for(s, p.ngroups, threads, K >= 5 ? 12 : 16);
  if constexpr (std::is_same<T, double>::value && K >= 2 && K <= 4) {
    static const int env_pair0 = [] { const char *e = getenv("B200SV_DENSE_PAIR0"); return e ? atoi(e) : 1; }();
    if (b0 >= 0 && env_pair0) {
      dense_pair0_kernel<K><<<grid, threads, 0, s.stream>>>((double2 *)s.data, p);
      B200_CUDA(cudaGetLastError());
      return;
    }
  }
  dense_kernel<T, K><<<grid, threads, 0, s.stream>>>((cx<T> *)s.data, p);
  B200_CUDA(cudaGetLastError());
}

void launch_dense(State &s, const int *targets, int k, const int *controls, int nc, const double *mat) {
  const bool f64 = s.precision == B200SV_F64;
#define CASE(K)                                                                  \
  case K:                                                                        \
    if (f64) launch_dense_t<double, K>(s, targets, controls, nc, mat);           \
    else launch_dense_t<float, K>(s, targets, controls, nc, mat);                \
    break;
  switch (k) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5)
  default:
    throw Error("launch_dense: k out of range");
  }
#undef CASE
}

// ------------------------------------------------------------------ dense, generic (k = 6..10)
// One CTA per group: stage the 2^k inputs in shared memory, every thread then
// produces outputs i = tid, tid+NT, ... reading the column-major matrix from
// global memory (coalesced across i).  Correctness fallback for blocks larger
// than fusion ever emits (reference K3/K4, thrust_kernels.hpp:900-1100).
struct GenericParams {
  uint64_t off_bits[kMaxDenseQubits];
  uint64_t ngroups;
  InsertList ins;
  int k;
};
template <typename T>
__global__ void dense_generic_kernel(cx<T> *__restrict__ psi, const cx<T> *__restrict__ mat,
                                     const __grid_constant__ GenericParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *in = reinterpret_cast<cx<T> *>(smem_raw);
  const int DIM = 1 << p.k;
  for (uint64_t g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
    const uint64_t base = insert_zeros(g, p.ins);
    for (int e = threadIdx.x; e < DIM; e += blockDim.x) {
      uint64_t o = 0;
      for (int b = 0; b < p.k; b++)
        if ((e >> b) & 1) o |= p.off_bits[b];
      in[e] = psi[base + o];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DIM; i += blockDim.x) {
      cx<T> acc = mk<T>(0, 0);
      for (int j = 0; j < DIM; j++) cfma(acc, mat[i + (size_t)DIM * j], in[j]);
      uint64_t o = 0;
      for (int b = 0; b < p.k; b++)
        if ((i >> b) & 1) o |= p.off_bits[b];
      psi[base + o] = acc;
    }
    __syncthreads();
  }
}

void launch_dense_generic(State &s, const int *targets, int k, const double *mat) {
  if (k > kMaxDenseQubits) throw Error("apply_matrix: more than 10 qubits is not supported");
  const size_t dim = 1ull << k, nelem = dim * dim;
  const size_t abytes = s.amp_bytes();
  // stage matrix: pinned host -> device scratch (stream ordered)
  void *hm = s.ensure_pinned(nelem * abytes);
  void *dm = s.ensure_scratch(nelem * abytes);
  B200_CUDA(cudaStreamSynchronize(s.stream));  // pinned buffer reuse
  if (s.precision == B200SV_F64) {
    memcpy(hm, mat, nelem * 16);
  } else {
    float *f = (float *)hm;
    for (size_t i = 0; i < 2 * nelem; i++) f[i] = (float)mat[i];
  }
  B200_CUDA(cudaMemcpyAsync(dm, hm, nelem * abytes, cudaMemcpyHostToDevice, s.stream));
  GenericParams p;
  p.k = k;
  std::vector<int> all(targets, targets + k);
  for (int b = 0; b < k; b++) p.off_bits[b] = 1ull << targets[b];
  std::sort(all.begin(), all.end());
  p.ins.n = k;
  for (int i = 0; i < k; i++) p.ins.pos[i] = (uint8_t)all[i];
  p.ngroups = s.total_amps() >> k;
  const int threads = (int)std::min<size_t>(dim, 256);
  const int grid = (int)std::min<uint64_t>(p.ngroups, (uint64_t)s.num_sms * 8);
  if (s.precision == B200SV_F64)
    dense_generic_kernel<double><<<grid, threads, dim * 16, s.stream>>>((double2 *)s.data, (const double2 *)dm, p);
  else
    dense_generic_kernel<float><<<grid, threads, dim * 8, s.stream>>>((float2 *)s.data, (const float2 *)dm, p);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ diagonal (streaming)
template <typename T> struct DiagParams {
  cx<T> d[1 << 5];   // tables up to 5 qubits ride in the parameter bank
  uint64_t total;
  int k;
  uint8_t q[kMaxDiagQubits];
};
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256)
diag_kernel(cx<T> *__restrict__ psi, const cx<T> *__restrict__ big_table, const __grid_constant__ DiagParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *tab = reinterpret_cast<cx<T> *>(smem_raw);
  const int dim = 1 << p.k;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) tab[i] = big_table ? big_table[i] : p.d[i];
  __syncthreads();
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t i0 = tid; i0 < p.total; i0 += nthreads * UNROLL) {
    cx<T> v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint64_t idx = i0 + u * nthreads;
      if (idx < p.total) v[u] = psi[idx];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint64_t idx = i0 + u * nthreads;
      if (idx < p.total) {
        uint32_t iv = 0;
        for (int j = 0; j < p.k; j++) iv |= (uint32_t)((idx >> p.q[j]) & 1ull) << j;
        psi[idx] = cmul(v[u], tab[iv]);
      }
    }
  }
}

// large diagonals (k > 10, e.g. the 2^m projector of a per-shot measurement of m qubits,
// statevector_state.hpp:960-1014): table stays in global memory (L2 resident), same streaming pass
struct BigDiagParams {
  uint64_t total;
  int k;
  uint8_t q[32];
};
template <typename T>
__global__ void __launch_bounds__(256)
diag_big_kernel(cx<T> *__restrict__ psi, const cx<T> *__restrict__ table, const __grid_constant__ BigDiagParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
    uint32_t iv = 0;
    for (int j = 0; j < p.k; j++) iv |= (uint32_t)((idx >> p.q[j]) & 1ull) << j;
    psi[idx] = cmul(psi[idx], table[iv]);
  }
}
static void launch_diagonal_big(State &s, const int *qubits, int k, const double *diag) {
  if (k > 28) throw Error("apply_diagonal_matrix: more than 28 qubits is not supported");
  const size_t dim = 1ull << k, bytes = dim * s.amp_bytes();
  void *hm = s.ensure_pinned(bytes);
  void *dm = s.ensure_scratch(bytes);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  if (s.precision == B200SV_F64) memcpy(hm, diag, bytes);
  else for (size_t i = 0; i < 2 * dim; i++) ((float *)hm)[i] = (float)diag[i];
  B200_CUDA(cudaMemcpyAsync(dm, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  BigDiagParams p;
  p.k = k; p.total = s.total_amps();
  for (int j = 0; j < k; j++) p.q[j] = (uint8_t)qubits[j];
  const int grid = grid_for(s, p.total, 256, 16);
  if (s.precision == B200SV_F64) diag_big_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, (const double2 *)dm, p);
  else diag_big_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, (const float2 *)dm, p);
  B200_CUDA(cudaGetLastError());
}

void launch_diagonal(State &s, const int *qubits, int k, const double *diag) {
  if (k > kMaxDiagQubits) { launch_diagonal_big(s, qubits, k, diag); return; }
  const int dim = 1 << k;
  const bool f64 = s.precision == B200SV_F64;
  void *big = nullptr;
  if (k > 5) {
    const size_t bytes = (size_t)dim * s.amp_bytes();
    void *hm = s.ensure_pinned(bytes);
    big = s.ensure_scratch(bytes);
    B200_CUDA(cudaStreamSynchronize(s.stream));
    if (f64) memcpy(hm, diag, bytes);
    else for (int i = 0; i < 2 * dim; i++) ((float *)hm)[i] = (float)diag[i];
    B200_CUDA(cudaMemcpyAsync(big, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  }
  const uint64_t total = s.total_amps();
  const int threads = 256;
  const int grid = grid_for(s, (total + 3) / 4, threads, 8);
  if (f64) {
    DiagParams<double> p;
    p.k = k; p.total = total;
    for (int j = 0; j < k; j++) p.q[j] = (uint8_t)qubits[j];
    if (k <= 5) for (int i = 0; i < dim; i++) p.d[i] = mk<double>(diag[2 * i], diag[2 * i + 1]);
    diag_kernel<double, 4><<<grid, threads, dim * 16, s.stream>>>((double2 *)s.data, (const double2 *)big, p);
  } else {
    DiagParams<float> p;
    p.k = k; p.total = total;
    for (int j = 0; j < k; j++) p.q[j] = (uint8_t)qubits[j];
    if (k <= 5) for (int i = 0; i < dim; i++) p.d[i] = mk<float>((float)diag[2 * i], (float)diag[2 * i + 1]);
    diag_kernel<float, 4><<<grid, threads, dim * 8, s.stream>>>((float2 *)s.data, (const float2 *)big, p);
  }
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ sub-cube kernels (controls in the index)
struct CubeParams {
  uint64_t ngroups;
  uint64_t mask0;  // bits OR'ed into the first element
  uint64_t mask1;  // bits OR'ed into the second element (pair kernels)
  InsertList ins;
};

template <typename T>
__global__ void __launch_bounds__(256) mcphase_kernel(cx<T> *__restrict__ psi, const __grid_constant__ CubeParams p,
                                                      cx<T> phase) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
    const uint64_t i = insert_zeros(g, p.ins) | p.mask0;
    psi[i] = cmul(psi[i], phase);
  }
}

// MODE 0: swap (mcx / mcswap)   MODE 1: mcy (d0 = -i*d1, d1 = i*d0) -- pure moves + sign flips: bit exact
template <typename T, int MODE>
__global__ void __launch_bounds__(256) pair_perm_kernel(cx<T> *__restrict__ psi, const __grid_constant__ CubeParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;  // pairs in flight per thread: 8 independent 16-byte loads before the first store
  for (uint64_t g0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < p.ngroups; g0 += stride * U) {
    cx<T> a[U], b[U];
    uint64_t i0[U], i1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t g = g0 + u * stride;
      const uint64_t base = insert_zeros(g < p.ngroups ? g : 0, p.ins);
      i0[u] = base | p.mask0;
      i1[u] = base | p.mask1;
      if (g < p.ngroups) { a[u] = psi[i0[u]]; b[u] = psi[i1[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (g0 + u * stride >= p.ngroups) continue;
      if (MODE == 0) {
        psi[i0[u]] = b[u];
        psi[i1[u]] = a[u];
      } else {
        psi[i0[u]] = mk<T>(b[u].y, -b[u].x);
        psi[i1[u]] = mk<T>(-a[u].y, a[u].x);
      }
    }
  }
}

static CubeParams cube_params(const State &s, std::vector<int> all) {
  CubeParams p;
  std::sort(all.begin(), all.end());
  p.ins.n = (int)all.size();
  for (size_t i = 0; i < all.size(); i++) p.ins.pos[i] = (uint8_t)all[i];
  p.ngroups = s.total_amps() >> all.size();
  p.mask0 = p.mask1 = 0;
  return p;
}

void launch_mcphase(State &s, const int *qubits, int k, double re, double im) {
  CubeParams p = cube_params(s, std::vector<int>(qubits, qubits + k));
  for (int j = 0; j < k; j++) p.mask0 |= 1ull << qubits[j];
  const int grid = grid_for(s, p.ngroups, 256, 16);
  if (s.precision == B200SV_F64)
    mcphase_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, p, mk<double>(re, im));
  else
    mcphase_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, p, mk<float>((float)re, (float)im));
  B200_CUDA(cudaGetLastError());
}

template <int MODE> static void launch_pair(State &s, const CubeParams &p) {
  const int grid = grid_for(s, p.ngroups, 256, 16);
  if (s.precision == B200SV_F64)
    pair_perm_kernel<double, MODE><<<grid, 256, 0, s.stream>>>((double2 *)s.data, p);
  else
    pair_perm_kernel<float, MODE><<<grid, 256, 0, s.stream>>>((float2 *)s.data, p);
  B200_CUDA(cudaGetLastError());
}

void launch_mcx(State &s, const int *controls, int nc, int target) {
  std::vector<int> all(controls, controls + nc);
  all.push_back(target);
  CubeParams p = cube_params(s, all);
  for (int c = 0; c < nc; c++) p.mask0 |= 1ull << controls[c];
  p.mask1 = p.mask0 | (1ull << target);
  launch_pair<0>(s, p);
}
void launch_mcy(State &s, const int *controls, int nc, int target) {
  std::vector<int> all(controls, controls + nc);
  all.push_back(target);
  CubeParams p = cube_params(s, all);
  for (int c = 0; c < nc; c++) p.mask0 |= 1ull << controls[c];
  p.mask1 = p.mask0 | (1ull << target);
  launch_pair<1>(s, p);
}
void launch_mcswap(State &s, const int *controls, int nc, int t0, int t1) {
  std::vector<int> all(controls, controls + nc);
  all.push_back(t0);
  all.push_back(t1);
  CubeParams p = cube_params(s, all);
  uint64_t cm = 0;
  for (int c = 0; c < nc; c++) cm |= 1ull << controls[c];
  p.mask0 = cm | (1ull << t0);
  p.mask1 = cm | (1ull << t1);
  launch_pair<0>(s, p);
}

// ------------------------------------------------------------------ general permutation (sequential swaps per group)
struct PermParams {
  uint64_t off_bits[kMaxDenseQubits];
  uint64_t ngroups;
  InsertList ins;
  int k, npairs;
  uint16_t pairs[2 * 1024];
};
template <typename T>
__global__ void __launch_bounds__(256) perm_kernel(cx<T> *psi, const __grid_constant__ PermParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
    const uint64_t base = insert_zeros(g, p.ins);
    for (int t = 0; t < p.npairs; t++) {
      uint64_t oa = 0, ob = 0;
      const int ea = p.pairs[2 * t], eb = p.pairs[2 * t + 1];
      for (int b = 0; b < p.k; b++) {
        if ((ea >> b) & 1) oa |= p.off_bits[b];
        if ((eb >> b) & 1) ob |= p.off_bits[b];
      }
      volatile cx<T> *vp = psi;  // sequential swaps may alias within a thread's group
      const T ax = vp[base + oa].x, ay = vp[base + oa].y;
      const T bx = vp[base + ob].x, by = vp[base + ob].y;
      vp[base + oa].x = bx; vp[base + oa].y = by;
      vp[base + ob].x = ax; vp[base + ob].y = ay;
    }
  }
}
void launch_permutation(State &s, const int *qubits, int k, const uint64_t *pairs, int npairs) {
  if (k > kMaxDenseQubits) throw Error("apply_permutation_matrix: more than 10 qubits is not supported");
  if (npairs > 1024) throw Error("apply_permutation_matrix: more than 1024 pairs is not supported");
  static thread_local PermParams p;
  p.k = k; p.npairs = npairs;
  std::vector<int> all(qubits, qubits + k);
  for (int b = 0; b < k; b++) p.off_bits[b] = 1ull << qubits[b];
  std::sort(all.begin(), all.end());
  p.ins.n = k;
  for (int i = 0; i < k; i++) p.ins.pos[i] = (uint8_t)all[i];
  for (int t = 0; t < 2 * npairs; t++) {
    if (pairs[t] >= (1ull << k)) throw Error("apply_permutation_matrix: pair index out of range");
    p.pairs[t] = (uint16_t)pairs[t];
  }
  p.ngroups = s.total_amps() >> k;
  const int grid = grid_for(s, p.ngroups, 256, 16);
  if (s.precision == B200SV_F64) perm_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, p);
  else perm_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, p);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ Pauli string
// pairs (i0, i0 ^ x_mask), i0 = insert_zero(i, x_max); Z-only strings use pos 0 and no swap.
template <typename T>
__device__ __forceinline__ void pauli_pair(cx<T> *__restrict__ psi, uint64_t i, uint64_t x_mask, uint64_t z_mask,
                                           int pos, cx<T> phase) {
  const uint64_t i0 = insert_zero(i, pos);
  const uint64_t i1 = i0 ^ (x_mask ? x_mask : 1ull);
  cx<T> a = psi[i0], b = psi[i1];
  if (x_mask) { const cx<T> t = a; a = b; b = t; }
  if (__popcll(i0 & z_mask) & 1) a = mk<T>(-a.x, -a.y);
  if (__popcll(i1 & z_mask) & 1) b = mk<T>(-b.x, -b.y);
  psi[i0] = cmul(a, phase);
  psi[i1] = cmul(b, phase);
}
template <typename T>
__global__ void __launch_bounds__(256) pauli_kernel(cx<T> *__restrict__ psi, uint64_t npairs, uint64_t x_mask,
                                                    uint64_t z_mask, int pos, cx<T> phase) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += stride)
    pauli_pair<T>(psi, i, x_mask, z_mask, pos, phase);
}
void launch_pauli(State &s, uint64_t x_mask, uint64_t z_mask, int x_max, double pre, double pim) {
  const uint64_t npairs = s.total_amps() >> 1;
  const int pos = x_mask ? x_max : 0;
  const int grid = grid_for(s, npairs, 256, 16);
  if (s.precision == B200SV_F64)
    pauli_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, npairs, x_mask, z_mask, pos,
                                                     mk<double>(pre, pim));
  else
    pauli_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, npairs, x_mask, z_mask, pos,
                                                    mk<float>((float)pre, (float)pim));
  B200_CUDA(cudaGetLastError());
}

// per-state Pauli (batched noisy shots): masks4[s] = {x_mask, z_mask, num_y, apply}
template <typename T>
__global__ void __launch_bounds__(256) batched_pauli_kernel(cx<T> *__restrict__ psi, int nq, uint64_t pairs_per_state,
                                                            const uint64_t *__restrict__ masks4) {
  const uint64_t s = blockIdx.y;
  const uint64_t x_mask = masks4[4 * s], z_mask = masks4[4 * s + 1], ny = masks4[4 * s + 2];
  if (!masks4[4 * s + 3] || (x_mask | z_mask) == 0) return;
  cx<T> phase = mk<T>(1, 0);                     // (-i)^num_y, add_y_phase (qubitvector.hpp:2275)
  if ((ny & 3) == 1) phase = mk<T>(0, -1);
  if ((ny & 3) == 2) phase = mk<T>(-1, 0);
  if ((ny & 3) == 3) phase = mk<T>(0, 1);
  const int pos = x_mask ? 63 - __clzll((long long)x_mask) : 0;
  cx<T> *st = psi + (s << nq);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs_per_state; i += stride)
    pauli_pair<T>(st, i, x_mask, z_mask, pos, phase);
}
void launch_batched_pauli(State &s, const uint64_t *masks4_host) {
  const size_t bytes = (size_t)s.nstates * 4 * sizeof(uint64_t);
  void *hm = s.ensure_pinned(bytes);
  void *dm = s.ensure_scratch(bytes);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(hm, masks4_host, bytes);
  B200_CUDA(cudaMemcpyAsync(dm, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  const uint64_t pps = s.amps_per_state() >> 1;
  int gx = (int)std::min<uint64_t>((pps + 255) / 256, std::max<uint64_t>(1, (uint64_t)s.num_sms * 16 / s.nstates));
  if (gx < 1) gx = 1;
  if (s.nstates > 65535) throw Error("apply_batched_pauli_ops: more than 65535 states per container");
  dim3 grid(gx, (unsigned)s.nstates);
  if (s.precision == B200SV_F64)
    batched_pauli_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, s.nq, pps, (const uint64_t *)dm);
  else
    batched_pauli_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, s.nq, pps, (const uint64_t *)dm);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ per-state collapse (batched measure / reset)
struct CollapseParams {
  int nq, k;
  uint8_t q[32];
};
template <typename T>
__global__ void __launch_bounds__(256)
collapse_kernel(cx<T> *__restrict__ psi, const __grid_constant__ CollapseParams p, const uint64_t *__restrict__ outcome,
                const double *__restrict__ scale, const uint8_t *__restrict__ active) {
  const uint64_t s = blockIdx.y;
  if (!active[s]) return;
  const uint64_t want = outcome[s];
  const T sc = (T)scale[s];
  cx<T> *st = psi + (s << p.nq);
  const uint64_t n = 1ull << p.nq, stride = (uint64_t)gridDim.x * blockDim.x;
  // the outcome deposited at the measured positions: one mask compare per amplitude; amplitudes that do not survive
  // are only written (a full-register measurement is a write-only pass)
  uint64_t qmask = 0, wbits = 0;
  for (int j = 0; j < p.k; j++) {
    qmask |= 1ull << p.q[j];
    wbits |= ((want >> j) & 1ull) << p.q[j];
  }
#pragma unroll 4
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if ((i & qmask) != wbits) { st[i] = mk<T>(0, 0); continue; }
    const cx<T> v = st[i];
    if (sc > (T)0) { st[i] = mk<T>(v.x * sc, v.y * sc); continue; }
    // scale <= 0: every qubit was measured, the single survivor is normalised by its own modulus
    const T r = (T)1 / sqrt(v.x * v.x + v.y * v.y);
    st[i] = mk<T>(v.x * r, v.y * r);
  }
}
void launch_collapse(State &s, const int *qubits, int k, const uint64_t *outcomes, const double *scales,
                     const uint8_t *active) {
  if (k > 32) throw Error("collapse: too many qubits");
  if (s.nstates > 65535) throw Error("collapse: more than 65535 states per container");
  const size_t S = (size_t)s.nstates;
  const size_t b_out = 0, b_sc = S * 8, b_act = 2 * S * 8, total = 2 * S * 8 + ((S + 15) & ~(size_t)15);
  char *hm = (char *)s.ensure_pinned(total);
  char *dm = (char *)s.ensure_scratch(total);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(hm + b_out, outcomes, S * 8);
  memcpy(hm + b_sc, scales, S * 8);
  memcpy(hm + b_act, active, S);
  B200_CUDA(cudaMemcpyAsync(dm, hm, total, cudaMemcpyHostToDevice, s.stream));
  CollapseParams p;
  p.nq = s.nq; p.k = k;
  for (int j = 0; j < k; j++) p.q[j] = (uint8_t)qubits[j];
  int gx = (int)std::min<uint64_t>((s.amps_per_state() + 255) / 256, std::max<uint64_t>(1, (uint64_t)s.num_sms * 16 / S));
  dim3 grid(std::max(gx, 1), (unsigned)S);
  if (s.precision == B200SV_F64)
    collapse_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, p, (const uint64_t *)(dm + b_out),
                                                        (const double *)(dm + b_sc), (const uint8_t *)(dm + b_act));
  else
    collapse_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, p, (const uint64_t *)(dm + b_out),
                                                       (const double *)(dm + b_sc), (const uint8_t *)(dm + b_act));
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ per-state matrices in one launch
// State s of a batched container applies matrix index[s] of a table (or nothing when index[s] < 0), multiplied by
// scale[s].  One launch for all shots replaces the reference's conditional-kernel loops: batched Kraus
// (MatrixMultNxN_conditional + apply_batched_kraus, qubitvector_thrust.hpp:2996-3177: every shot applies the Kraus
// operator its random draw selected, renormalised by 1/sqrt(p)) and per-parameter matrices of a bound circuit
// (apply_batched_matrix, :1578-1611).  K <= 3; the table (<= a few KiB) is read through L1/L2.
struct BatchedMatParams {
  uint64_t off[8];
  InsertList ins;
  uint64_t groups_per_state;
  int nq;
};
template <typename T, int K>
__global__ void __launch_bounds__(256)
batched_matrix_kernel(cx<T> *__restrict__ psi, const double2 *__restrict__ table, const int *__restrict__ index,
                      const double *__restrict__ scale, const __grid_constant__ BatchedMatParams p) {
  constexpr int DIM = 1 << K;
  const uint64_t st = blockIdx.y;
  const int mi = index[st];
  if (mi < 0) return;
  const double sc = scale[st];
  const double2 *m = table + (size_t)mi * DIM * DIM;  // column major like apply_matrix: m[i + DIM * j]
  cx<T> *base_ptr = psi + (st << p.nq);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.groups_per_state; g += stride) {
    const uint64_t base = insert_zeros(g, p.ins);
    cx<T> in[DIM];
#pragma unroll
    for (int e = 0; e < DIM; e++) in[e] = base_ptr[base + p.off[e]];
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      cx<T> acc = mk<T>(0, 0);
#pragma unroll
      for (int j = 0; j < DIM; j++) {
        const double2 e = m[i + DIM * j];
        cfma(acc, mk<T>((T)(e.x * sc), (T)(e.y * sc)), in[j]);
      }
      base_ptr[base + p.off[i]] = acc;
    }
  }
}
void launch_batched_matrix(State &s, const int *qubits, int k, const double *mats, int nmats, const int *index,
                           const double *scale) {
  if (k < 1 || k > 3) throw Error("batched matrix: 1..3 qubits");
  if (s.nstates > 65535) throw Error("batched matrix: more than 65535 states per container");
  const size_t S = (size_t)s.nstates, msz = ((size_t)1 << (2 * k)) * 16 * (size_t)nmats;
  const size_t b_idx = (msz + 15) & ~(size_t)15, b_sc = b_idx + ((S * 4 + 15) & ~(size_t)15), total = b_sc + S * 8;
  char *hm = (char *)s.ensure_pinned(total);
  char *dm = (char *)s.ensure_scratch(total);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(hm, mats, msz);
  memcpy(hm + b_idx, index, S * 4);
  memcpy(hm + b_sc, scale, S * 8);
  B200_CUDA(cudaMemcpyAsync(dm, hm, total, cudaMemcpyHostToDevice, s.stream));
  BatchedMatParams p;
  p.nq = s.nq;
  std::vector<int> sorted(qubits, qubits + k);
  for (int e = 0; e < (1 << k); e++) {
    uint64_t o = 0;
    for (int b = 0; b < k; b++)
      if ((e >> b) & 1) o |= 1ull << qubits[b];
    p.off[e] = o;
  }
  std::sort(sorted.begin(), sorted.end());
  p.ins.n = k;
  for (int i = 0; i < k; i++) p.ins.pos[i] = (uint8_t)sorted[i];
  p.groups_per_state = s.amps_per_state() >> k;
  int gx = (int)std::min<uint64_t>((p.groups_per_state + 255) / 256, std::max<uint64_t>(1, (uint64_t)s.num_sms * 16 / S));
  dim3 grid(std::max(gx, 1), (unsigned)S);
  const double2 *tb = (const double2 *)dm;
  const int *ix = (const int *)(dm + b_idx);
  const double *sc = (const double *)(dm + b_sc);
#define BM(K)                                                                                                     \
  case K:                                                                                                         \
    if (s.precision == B200SV_F64) batched_matrix_kernel<double, K><<<grid, 256, 0, s.stream>>>((double2 *)s.data, tb, ix, sc, p); \
    else batched_matrix_kernel<float, K><<<grid, 256, 0, s.stream>>>((float2 *)s.data, tb, ix, sc, p);             \
    break;
  switch (k) { BM(1) BM(2) BM(3) }
#undef BM
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ density-matrix line gather
template <typename T>
__global__ void __launch_bounds__(256) gather_line_kernel(const cx<T> *__restrict__ psi, cx<T> *__restrict__ out,
                                                          int row_bits, uint64_t xor_mask) {
  const uint64_t n = 1ull << row_bits, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = psi[(i ^ xor_mask) + (i << row_bits)];
}
void launch_gather_line(State &s, int row_bits, uint64_t xor_mask, void *host_out) {
  const size_t n = 1ull << row_bits, bytes = n * s.amp_bytes();
  void *dm = s.ensure_scratch(bytes);
  void *hm = s.ensure_pinned(bytes);
  const int grid = grid_for(s, n, 256, 16);
  if (s.precision == B200SV_F64) gather_line_kernel<double><<<grid, 256, 0, s.stream>>>((const double2 *)s.data, (double2 *)dm, row_bits, xor_mask);
  else gather_line_kernel<float><<<grid, 256, 0, s.stream>>>((const float2 *)s.data, (float2 *)dm, row_bits, xor_mask);
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(hm, dm, bytes, cudaMemcpyDeviceToHost, s.stream));
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(host_out, hm, bytes);
}

// ------------------------------------------------------------------ init
template <typename T> __global__ void set_ket0_kernel(cx<T> *psi, int nq, int64_t nstates) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nstates) psi[(uint64_t)s << nq] = mk<T>(1, 0);
}
void launch_init(State &s, bool ket0) {
  B200_CUDA(cudaMemsetAsync(s.data, 0, s.total_amps() * s.amp_bytes(), s.stream));
  if (!ket0) return;
  const int grid = (int)((s.nstates + 255) / 256);
  if (s.precision == B200SV_F64) set_ket0_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, s.nq, s.nstates);
  else set_ket0_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, s.nq, s.nstates);
  B200_CUDA(cudaGetLastError());
}

// initialize_component (qubitvector.hpp:879-900): for each group, cache = d[inds[0]];
// d[inds[i]] = cache * state[i].
template <typename T>
__global__ void __launch_bounds__(256)
init_component_kernel(cx<T> *__restrict__ psi, const cx<T> *__restrict__ comp, const __grid_constant__ GenericParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const int DIM = 1 << p.k;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
    const uint64_t base = insert_zeros(g, p.ins);
    const cx<T> cache = psi[base];
    for (int e = 0; e < DIM; e++) {
      uint64_t o = 0;
      for (int b = 0; b < p.k; b++)
        if ((e >> b) & 1) o |= p.off_bits[b];
      psi[base + o] = cmul(cache, comp[e]);
    }
  }
}
void launch_init_component(State &s, const int *qubits, int k, const double *state) {
  if (k > kMaxDenseQubits) throw Error("initialize_component: more than 10 qubits is not supported");
  const size_t dim = 1ull << k, bytes = dim * s.amp_bytes();
  void *hm = s.ensure_pinned(bytes);
  void *dm = s.ensure_scratch(bytes);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  if (s.precision == B200SV_F64) memcpy(hm, state, bytes);
  else for (size_t i = 0; i < 2 * dim; i++) ((float *)hm)[i] = (float)state[i];
  B200_CUDA(cudaMemcpyAsync(dm, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  GenericParams p;
  p.k = k;
  std::vector<int> all(qubits, qubits + k);
  for (int b = 0; b < k; b++) p.off_bits[b] = 1ull << qubits[b];
  std::sort(all.begin(), all.end());
  p.ins.n = k;
  for (int i = 0; i < k; i++) p.ins.pos[i] = (uint8_t)all[i];
  p.ngroups = s.total_amps() >> k;
  const int grid = grid_for(s, p.ngroups, 256, 16);
  if (s.precision == B200SV_F64)
    init_component_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, (const double2 *)dm, p);
  else
    init_component_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, (const float2 *)dm, p);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ global-qubit exchange helpers
// pack: buf[j] = psi[insert_zero(begin+j, q) | bit<<q]   unpack: the inverse.
template <typename T, bool UNPACK>
__global__ void __launch_bounds__(256) pack_half_kernel(cx<T> *__restrict__ psi, cx<T> *__restrict__ buf, int q,
                                                        uint64_t bitmask, uint64_t begin, uint64_t count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
    const uint64_t i = insert_zero(begin + j, q) | bitmask;
    if (UNPACK) psi[i] = buf[j];
    else buf[j] = psi[i];
  }
}
void launch_pack_half(State &s, int q, int bit, uint64_t begin, uint64_t count, void *buf, bool unpack) {
  const int grid = grid_for(s, count, 256, 16);
  const uint64_t bm = bit ? (1ull << q) : 0;
  if (s.precision == B200SV_F64) {
    if (unpack) pack_half_kernel<double, true><<<grid, 256, 0, s.stream>>>((double2 *)s.data, (double2 *)buf, q, bm, begin, count);
    else pack_half_kernel<double, false><<<grid, 256, 0, s.stream>>>((double2 *)s.data, (double2 *)buf, q, bm, begin, count);
  } else {
    if (unpack) pack_half_kernel<float, true><<<grid, 256, 0, s.stream>>>((float2 *)s.data, (float2 *)buf, q, bm, begin, count);
    else pack_half_kernel<float, false><<<grid, 256, 0, s.stream>>>((float2 *)s.data, (float2 *)buf, q, bm, begin, count);
  }
  B200_CUDA(cudaGetLastError());
}

// In-place swap with a peer-mapped partner chunk over NVLink (CSwapChunk_func,
// thrust_kernels.hpp:1884): lower chunk's (q=1) half <-> upper chunk's (q=0) half.
template <typename T>
__global__ void __launch_bounds__(256) chunk_swap_peer_kernel(cx<T> *__restrict__ mine, cx<T> *__restrict__ peer, int q,
                                                              uint64_t mine_mask, uint64_t peer_mask, uint64_t begin,
                                                              uint64_t count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;  // remote loads in flight per thread
  for (uint64_t j0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < count; j0 += stride * U) {
    cx<T> a[U], b[U];
    uint64_t iz[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      iz[u] = insert_zero(begin + (j < count ? j : 0), q);
      if (j < count) { b[u] = peer[iz[u] | peer_mask]; a[u] = mine[iz[u] | mine_mask]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      if (j < count) { mine[iz[u] | mine_mask] = b[u]; peer[iz[u] | peer_mask] = a[u]; }
    }
  }
}
void launch_chunk_swap_peer(State &s, int q, void *peer, int upper, int half) {
  const uint64_t npairs = s.amps_per_state() >> 1;
  const uint64_t count = npairs >> 1, begin = half ? count : 0;
  const uint64_t bit = 1ull << q;
  const uint64_t mine_mask = upper ? 0 : bit, peer_mask = upper ? bit : 0;
  const int grid = grid_for(s, count, 256, 16);
  if (s.precision == B200SV_F64)
    chunk_swap_peer_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, (double2 *)peer, q, mine_mask, peer_mask, begin, count);
  else
    chunk_swap_peer_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, (float2 *)peer, q, mine_mask, peer_mask, begin, count);
  B200_CUDA(cudaGetLastError());
}

// Contiguous range swap between two chunks (apply_chunk_swap(chunk, dest_offset, src_offset, size),
// qubitvector.hpp:1824-1840: the sub-block shuffle of apply_multi_chunk_swap; also the whole-chunk exchange of a swap
// between two global qubits).  `peer` may live on another GPU (peer access over NVLink) or alias this chunk's device.
// 32-byte accesses, 4 in flight per thread in each direction.
__global__ void __launch_bounds__(256) swap_range_kernel(ulonglong4 *__restrict__ a, ulonglong4 *__restrict__ b,
                                                         uint64_t count32) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (uint64_t j0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < count32; j0 += stride * U) {
    ulonglong4 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      if (j < count32) { y[u] = b[j]; x[u] = a[j]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      if (j < count32) { a[j] = y[u]; b[j] = x[u]; }
    }
  }
}
__global__ void __launch_bounds__(256) swap_range16_kernel(uint4 *__restrict__ a, uint4 *__restrict__ b, uint64_t count16) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count16; j += stride) {
    const uint4 x = a[j], y = b[j];
    a[j] = y;
    b[j] = x;
  }
}
void launch_swap_range_peer(State &s, uint64_t dest_offset, void *peer, uint64_t src_offset, uint64_t count) {
  char *a = (char *)s.data + dest_offset * s.amp_bytes();
  char *b = (char *)peer + src_offset * s.amp_bytes();
  const uint64_t bytes = count * s.amp_bytes();
  if (bytes == 0) return;
  if ((((uintptr_t)a | (uintptr_t)b | bytes) & 31) == 0) {
    const uint64_t n32 = bytes >> 5;
    swap_range_kernel<<<grid_for(s, (n32 + 3) / 4, 256, 8), 256, 0, s.stream>>>((ulonglong4 *)a, (ulonglong4 *)b, n32);
  } else if ((((uintptr_t)a | (uintptr_t)b | bytes) & 15) == 0) {
    const uint64_t n16 = bytes >> 4;
    swap_range16_kernel<<<grid_for(s, n16, 256, 8), 256, 0, s.stream>>>((uint4 *)a, (uint4 *)b, n16);
  } else {
    throw Error("swap range: ranges must be 16-byte aligned");
  }
  B200_CUDA(cudaGetLastError());
}

// k global qubits <-> k local qubits in one in-place pass over peer mappings
struct MultiSwapParams {
  void *peer[16];       // by global-bit value
  uint32_t lvals[16];   // the l classes this rank initiates
  int nl_classes;
  uint32_t my_g;
  int k;
  uint8_t lq[4];
  InsertList ins;       // sorted local swap positions
  uint64_t count;       // 2^(nq-k) indices per class
};
template <typename T>
__global__ void __launch_bounds__(256) multi_swap_kernel(cx<T> *__restrict__ mine, const __grid_constant__ MultiSwapParams p) {
  // classes (= partner shards) are interleaved over consecutive CTAs, so that all partners are served at the same
  // rate from the first wave on (one class after the other would aim every shard at the same partner first)
  const uint32_t ncls = (uint32_t)p.nl_classes;
  const uint32_t l = p.lvals[blockIdx.x % ncls];
  uint64_t lmask = 0, gmask = 0;
  for (int b = 0; b < p.k; b++) {
    if ((l >> b) & 1) lmask |= 1ull << p.lq[b];
    if ((p.my_g >> b) & 1) gmask |= 1ull << p.lq[b];
  }
  cx<T> *peer = (cx<T> *)p.peer[l];
  const uint64_t stride = (uint64_t)(gridDim.x / ncls) * blockDim.x;
  constexpr int U = 4;  // remote loads in flight per thread (NVLink latency ~2 us)
  for (uint64_t j0 = (uint64_t)(blockIdx.x / ncls) * blockDim.x + threadIdx.x; j0 < p.count; j0 += stride * U) {
    cx<T> a[U], b[U];
    uint64_t base[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      base[u] = insert_zeros(j < p.count ? j : 0, p.ins);
      if (j < p.count) { b[u] = peer[base[u] | gmask]; a[u] = mine[base[u] | lmask]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      if (j < p.count) { mine[base[u] | lmask] = b[u]; peer[base[u] | gmask] = a[u]; }
    }
  }
}
void launch_multi_swap_peer(State &s, int k, const int *local_q, uint32_t my_g, void *const *peers) {
  if (k < 1 || k > 4) throw Error("multi swap: 1..4 qubits");
  MultiSwapParams p;
  p.k = k; p.my_g = my_g;
  std::vector<int> sorted(local_q, local_q + k);
  for (int b = 0; b < k; b++) p.lq[b] = (uint8_t)local_q[b];
  std::sort(sorted.begin(), sorted.end());
  p.ins.n = k;
  for (int b = 0; b < k; b++) p.ins.pos[b] = (uint8_t)sorted[b];
  p.count = s.amps_per_state() >> k;
  const uint32_t dim = 1u << k, half = dim >> 1;
  p.nl_classes = 0;
  for (uint32_t l = 0; l < dim; l++) {
    if (l == my_g) continue;
    const uint32_t d = (l - my_g) & (dim - 1);
    if (d < half || (d == half && my_g < l)) p.lvals[p.nl_classes++] = l;
    p.peer[l] = peers[l];
  }
  if (p.nl_classes == 0) return;
  int gx = (int)std::min<uint64_t>((p.count + 255) / 256, std::max<uint64_t>(1, (uint64_t)s.num_sms * 16 / p.nl_classes));
  const unsigned grid = (unsigned)std::max(gx, 1) * (unsigned)p.nl_classes;
  if (s.precision == B200SV_F64) multi_swap_kernel<double><<<grid, 256, 0, s.stream>>>((double2 *)s.data, p);
  else multi_swap_kernel<float><<<grid, 256, 0, s.stream>>>((float2 *)s.data, p);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ wide diagonal layers
// ANY number of commuting diagonal 1-/2-qubit gates (cp, cz, rz, p, rzz ... over all qubits of the register) in ONE
// streaming pass, without a 2^k table: with every gate written as d(ba, bb) = c * A_a^ba * A_b^bb * G_ab^(ba bb) the
// whole layer is   phase(i) = C * prod_{a set} A_a * prod_{a<b both set} G_ab,
// a quadratic form in the index bits.  A CTA walks chunks of 2^12 consecutive amplitudes: per chunk the high bits are
// fixed, so S(hi) and the twelve low-bit multipliers M_u(hi) = A_u * prod_{b high, set} G_ub are computed once (by 44
// threads), each thread then builds the factor of its 8 thread-id bits (<= 36 predicated complex multiplies) and the 16
// phases of its 16 amplitudes by doubling: ~6 complex multiplies per amplitude, far below the ~22 a pass can afford at
// HBM speed.  Replaces chains of DiagonalMult* launches / 2^k-entry tables (thrust_kernels.hpp:1318-1444) for layers
// wider than a table can be (QFT: all controlled phases between two groups of Hadamards).
struct DiagLayerParams {
  const double2 *A;   // [nq]
  const double2 *G;   // [nq][nq], G[a * nq + b] for a != b (symmetric), 1 elsewhere
  double2 C;
  int nq;
};
__device__ __forceinline__ double2 cmulz(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename T, int LB>
__global__ void __launch_bounds__(288, LB == 12 ? 1 : 2) diag_layer_kernel(cx<T> *__restrict__ psi, const __grid_constant__ DiagLayerParams p,
                                                         uint64_t nchunks, int lo_bits) {
  extern __shared__ double2 sTab[];            // A[nq], then G[nq][nq]: staged once per CTA
  __shared__ double2 sMbuf[2][12], sSbuf[2];   // per-chunk factors, double buffered
  constexpr int EB = LB - 8, NE = 1 << EB;     // element bits / amplitudes per thread (chunk = 2^LB amplitudes)
  const int tid = threadIdx.x, nq = p.nq;
  double2 *sA = sTab, *sG = sTab + nq;
  for (int e = tid; e < nq + nq * nq; e += blockDim.x) sTab[e] = p.A[e];  // A and G are contiguous in device memory
  __syncthreads();
  // warp 8 (threads 256..287) is the producer: it computes the factors of the NEXT chunk while warps 0..7 apply the
  // current one; one barrier per chunk hands the buffers over
  auto produce = [&](uint64_t hi, int buf) {
    const int lane = tid - 256;
    double2 part = make_double2(1.0, 0.0);  // S(hi) = C * prod_{a high, set} (A_a * prod_{b > a, set} G_ab): one a per lane
    for (int a = lo_bits + lane; a < nq; a += 32) {
      if (!((hi >> (a - lo_bits)) & 1)) continue;
      part = cmulz(part, sA[a]);
      for (int b = a + 1; b < nq; b++)
        if ((hi >> (b - lo_bits)) & 1) part = cmulz(part, sG[a * nq + b]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double2 other = make_double2(__shfl_xor_sync(0xffffffffu, part.x, o), __shfl_xor_sync(0xffffffffu, part.y, o));
      part = cmulz(part, other);
    }
    if (lane == 0) sSbuf[buf] = cmulz(p.C, part);
    if (lane < 12) {  // M_u(hi) = A_u * prod_{b high, set} G_ub
      double2 m = make_double2(1.0, 0.0);
      if (lane < lo_bits) {
        m = sA[lane];
        for (int b = lo_bits; b < nq; b++)
          if ((hi >> (b - lo_bits)) & 1) m = cmulz(m, sG[lane * nq + b]);
      }
      sMbuf[buf][lane] = m;
    }
  };
  if (tid >= 256 && blockIdx.x < nchunks) produce(blockIdx.x, 0);
  __syncthreads();
  int it = 0;
  for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x, it++) {
    const double2 *sM = sMbuf[it & 1];
    const double2 sS = sSbuf[it & 1];
    if (tid >= 256 && c + gridDim.x < nchunks) produce(c + gridDim.x, (it + 1) & 1);
    // ---- thread id = chunk-local bits 0..7 (consecutive threads touch consecutive amplitudes: coalesced), the 16
    // amplitudes of a thread differ in chunk-local bits 8..11.  Thread factor over its id bits, then the 16 phases by doubling.
    const bool active = tid < 256 && (lo_bits >= 8 || tid < (1 << lo_bits));
    const int ebits = lo_bits >= 8 ? lo_bits - 8 : 0;   // element bits that exist in this chunk (EB for full chunks)
    if (active) {
      // loads first: their latency runs under the phase arithmetic below
      cx<T> *base = psi + (c << lo_bits) + (uint64_t)tid;
      cx<T> v[NE];
#pragma unroll
      for (int e = 0; e < NE; e++)
        if (e < (1 << ebits)) v[e] = base[(uint64_t)e << 8];
      double2 f = sS;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (!((tid >> u) & 1)) continue;
        f = cmulz(f, sM[u]);
#pragma unroll
        for (int v = u + 1; v < 8; v++)
          if ((tid >> v) & 1) f = cmulz(f, sG[u * nq + v]);
      }
      double2 w[EB];  // multiplier of element bit e (position 8 + e) given this thread's id bits
#pragma unroll
      for (int e = 0; e < EB; e++) {
        double2 m = make_double2(1.0, 0.0);
        if (e < ebits) {
          m = sM[8 + e];
#pragma unroll
          for (int u = 0; u < 8; u++)
            if ((tid >> u) & 1) m = cmulz(m, sG[(8 + e) * nq + u]);
        }
        w[e] = m;
      }
      double2 ph[NE];
      ph[0] = f;
#pragma unroll
      for (int e = 0; e < EB; e++)
#pragma unroll
        for (int j = 0; j < (1 << e); j++) {
          double2 m = cmulz(ph[j], w[e]);
          if (e < ebits) {
#pragma unroll
            for (int x = 0; x < e; x++)
              if ((j >> x) & 1) m = cmulz(m, sG[(8 + x) * nq + 8 + e]);
          }
          ph[j | (1 << e)] = m;
        }
#pragma unroll
      for (int e = 0; e < NE; e++)
        if (e < (1 << ebits))
          base[(uint64_t)e << 8] = mk<T>((T)(ph[e].x * (double)v[e].x - ph[e].y * (double)v[e].y),
                                         (T)(ph[e].x * (double)v[e].y + ph[e].y * (double)v[e].x));
    }
    __syncthreads();
  }
}
void launch_diag_layer(State &s, int ngates, const int *nq, const uint64_t *qubits, const double *diags) {
  typedef std::complex<double> cd;
  if (s.nq < 4) throw Error("diagonal layer: needs at least 4 qubits");
  const int n = s.nq;
  std::vector<cd> A((size_t)n, cd(1.0)), G((size_t)n * n, cd(1.0));
  cd Cc(1.0);
  for (int g = 0; g < ngates; g++) {
    const cd *d = reinterpret_cast<const cd *>(diags + 8 * (size_t)g);
    const int a = (int)qubits[2 * g];
    if (nq[g] == 1) {
      if (d[0] == cd(0.0)) throw Error("diagonal layer: zero entry (use apply_diagonal_matrix)");
      Cc *= d[0];
      A[a] *= d[1] / d[0];
    } else {
      const int b = (int)qubits[2 * g + 1];
      if (d[0] == cd(0.0) || d[1] == cd(0.0) || d[2] == cd(0.0)) throw Error("diagonal layer: zero entry (use apply_diagonal_matrix)");
      Cc *= d[0];
      A[a] *= d[1] / d[0];     // index = ba + 2 bb
      A[b] *= d[2] / d[0];
      const cd gab = d[3] * d[0] / (d[1] * d[2]);
      G[(size_t)a * n + b] *= gab;
      G[(size_t)b * n + a] *= gab;
    }
  }
  const size_t bytes = ((size_t)n + (size_t)n * n) * 16;
  char *hm = (char *)s.ensure_pinned(bytes);
  char *dm = (char *)s.ensure_scratch(bytes);
  B200_CUDA(cudaStreamSynchronize(s.stream));
  memcpy(hm, A.data(), (size_t)n * 16);
  memcpy(hm + (size_t)n * 16, G.data(), (size_t)n * n * 16);
  B200_CUDA(cudaMemcpyAsync(dm, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  DiagLayerParams p;
  p.A = (const double2 *)dm;
  p.G = (const double2 *)(dm + (size_t)n * 16);
  p.C = make_double2(Cc.real(), Cc.imag());
  p.nq = n;
  // chunk = 2^11 amplitudes (8 per thread, two CTAs per SM: one CTA's loads run under the other's arithmetic) for double,
  // B200SV_DIAG_LAYER_CHUNK_BITS = 12 selects the 16-per-thread form
  static const int env_lb = [] { const char *e = getenv("B200SV_DIAG_LAYER_CHUNK_BITS"); return e ? atoi(e) : 11; }();
  const int LB = env_lb == 12 ? 12 : 11;
  const int lo_bits = std::min(LB, n);
  const size_t smem = ((size_t)n + (size_t)n * n) * 16;  // <= 26 KiB at 40 qubits
  auto launch = [&](void *data, int grid, uint64_t nchunks) {
    if (s.precision == B200SV_F64) {
      if (LB == 12) diag_layer_kernel<double, 12><<<grid, 288, smem, s.stream>>>((double2 *)data, p, nchunks, lo_bits);
      else diag_layer_kernel<double, 11><<<grid, 288, smem, s.stream>>>((double2 *)data, p, nchunks, lo_bits);
    } else {
      if (LB == 12) diag_layer_kernel<float, 12><<<grid, 288, smem, s.stream>>>((float2 *)data, p, nchunks, lo_bits);
      else diag_layer_kernel<float, 11><<<grid, 288, smem, s.stream>>>((float2 *)data, p, nchunks, lo_bits);
    }
  };
  // batched containers: every state is its own run of chunks with the same layer (hi wraps per state)
  const uint64_t chunks_per_state = 1ull << (n - lo_bits);
  if (s.nstates != 1 && chunks_per_state != 1) {
    for (int64_t st = 0; st < s.nstates; st++) {
      State v = s;
      v.nstates = 1;
      v.data = (char *)s.data + ((uint64_t)st << n) * s.amp_bytes();
      const int grid = (int)std::min<uint64_t>(chunks_per_state, (uint64_t)s.num_sms * 8);
      launch(v.data, grid, chunks_per_state);
    }
    B200_CUDA(cudaGetLastError());
    return;
  }
  const uint64_t nchunks = s.nstates == 1 ? chunks_per_state : (uint64_t)s.nstates;
  const int grid = (int)std::min<uint64_t>(nchunks, (uint64_t)s.num_sms * 8);
  launch(s.data, grid, nchunks);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ FP64 issue-peak probe (measurement aid)
// DFMA chains whose multiplier comes straight from the constant bank (the operand form the tile rounds use): what the
// FP64 pipe of THIS device sustains, measured next to the bench numbers it is the denominator of
// (tools/micro/dfma_operands.cu is the standalone version; profiles/r01_fp64_pipe_microbench.md).
struct PeakParams { double m[32]; };
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, const __grid_constant__ PeakParams p, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        a[i] = fma(a[i], p.m[j], a[i + 1]);
        a[i + 1] = fma(a[i + 1], p.m[j], a[i + 2]);
        a[i + 2] = fma(a[i + 2], p.m[j], a[i + 3]);
        a[i + 3] = fma(a[i + 3], p.m[j], a[i]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// runs the probe for ~duration_ms; burst = best single launch, sustained = mean over the second half (power-capped clock)
void measure_fp64_peak(int device, double duration_ms, double *burst_tflops, double *sustained_tflops) {
  B200_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, iters = 1024;
  double *out = nullptr;
  B200_CUDA(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double)));
  PeakParams p;
  for (int i = 0; i < 32; i++) p.m[i] = 1.0 + 1e-9 * i;
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const double flops = (double)blocks * 256 * iters * 32 * 16 * 2;
  std::vector<double> tf;
  double spent = 0;
  dfma_peak_kernel<<<blocks, 256>>>(out, p, iters);  // warm-up
  B200_CUDA(cudaDeviceSynchronize());
  while (spent < duration_ms || tf.size() < 4) {
    B200_CUDA(cudaEventRecord(e0));
    dfma_peak_kernel<<<blocks, 256>>>(out, p, iters);
    B200_CUDA(cudaEventRecord(e1));
    B200_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    tf.push_back(flops / ms / 1e9);
    spent += ms;
  }
  double best = 0, sum = 0;
  for (double v : tf) best = std::max(best, v);
  for (size_t i = tf.size() / 2; i < tf.size(); i++) sum += tf[i];
  *burst_tflops = best;
  *sustained_tflops = sum / (double)(tf.size() - tf.size() / 2);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
}

}  // namespace b200sv
