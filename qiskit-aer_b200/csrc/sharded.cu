// b200sv sharded executor: one register of n qubits split over 2^g shards by its top g qubits (host code + the small
// device helpers of the exchange).
//
// Role in the reference: ParallelStateExecutor (src/simulators/parallel_state_executor.hpp: chunk placement :318-376,
// apply_ops_chunks :772, apply_chunk_swap :1134-1336, apply_multi_chunk_swap :1339-1552), the chunk manager's device
// placement / peer access (statevector/chunk/chunk_manager.hpp:129-135,166-396) and Statevector::Executor's cross-chunk
// reductions (statevector/statevector_executor.hpp: expval_pauli :551, sample_measure :1149).  Redesigned for
// B200 + NVLink:
//
//   * one slice per GPU, no buffer chunk, a logical -> physical qubit map instead of swap-backs, epoch scheduling
//     (planner.cu) -- whole epochs of gates ride the tile engine between two exchanges;
//   * a shard is (device, slice, compute stream, copy stream, staging area, flag words).  The shards of one handle live
//     in this process (one process drives several GPUs: the layout Aer's Controller uses) and / or in other processes
//     (one process per GPU, `torchrun`): those are attached through CUDA IPC handles.  The executor code is the same;
//     only the cross-shard ordering primitive differs -- CUDA events between shards of one process, flag words written
//     over NVLink (st.release.sys) and awaited by a one-warp kernel (ld.acquire.sys) between processes.  No host
//     rendezvous, no collective library on the data path;
//   * the global-qubit exchange (k local <-> k global qubits, all-to-all among 2^k shards) is STAGED and PIPELINED:
//     the slice is cut into 2^s slabs along s index bits that neither the exchange nor the tile passes next to it
//     touch; per slab, the copy engines PUSH the outgoing sub-blocks into the receivers' staging areas
//     (cudaMemcpy2DAsync over NVLink: posted writes, no SM involved) while the SMs run the last pass before / the
//     first pass after the exchange on the neighbouring slabs; an unstage kernel moves a received slab into place.
//     When no slab bits are free or the staging area is too small the exchange falls back to the in-place peer swap
//     kernel (gates.cu: multi_swap_kernel) between two rendezvous.
#include <array>
#include <complex>
#include <memory>

#include "common.cuh"

namespace b200sv {

enum { OP_MATRIX = 0, OP_DIAGONAL = 1, OP_MCX = 2, OP_MCY = 3, OP_MCPHASE = 4, OP_MCSWAP = 5, OP_MCU = 6 };
enum { F_READY = 0, F_DONE = 1, F_PUSHED = 2, F_UNSTAGED = 3, F_KINDS = 4 };
constexpr int kMaxWorld = 64;
constexpr int kEvRing = 64;
constexpr size_t kFlagBytes = 4096;  // flags [F_KINDS][world] + error word, at the start of the aux allocation

// ------------------------------------------------------------------------------------------ device helpers
struct SignalParams {
  uint64_t *target[kMaxWorld];
  int n;
  uint64_t value;
};
// everything issued before this kernel on its stream (tile passes, copies, unstage) is complete; publish `value`
__global__ void shard_signal_kernel(const __grid_constant__ SignalParams p) {
  if ((int)threadIdx.x < p.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.target[threadIdx.x]), "l"(p.value) : "memory");
  }
}
struct WaitParams {
  const uint64_t *flag[kMaxWorld];
  int n;
  uint64_t value;
  uint32_t *err;
  unsigned long long timeout_ns;
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// blocks its stream until every awaited flag has reached `value` (a partner that never arrives trips the error word
// instead of hanging the GPU)
__global__ void shard_wait_kernel(const __grid_constant__ WaitParams p) {
  if ((int)threadIdx.x < p.n) {
    const unsigned long long t0 = global_ns();
    for (;;) {
      uint64_t v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p.flag[threadIdx.x]) : "memory");
      if (v >= p.value) break;
      __nanosleep(256);
      if (global_ns() - t0 > p.timeout_ns) { atomicExch(p.err, 1u); break; }
    }
  }
}

// staging slot u (compact, 2^sub_bits units) -> the amplitudes whose fixed bits read fixed[u]; one launch moves every
// sender's slot of a slab into place.  UNIT = one amplitude (uint4: complex<double>, uint2: complex<float>).
struct UnstageParams {
  uint64_t fixed[16];   // per slot: OR mask of the fixed index bits (exchange positions = sender id, slab bits)
  InsertList ins;       // sorted fixed positions
  uint64_t sub;         // units per slot
  int nslots;
};
template <typename UNIT>
__global__ void __launch_bounds__(256) unstage_kernel(UNIT *__restrict__ psi, const UNIT *__restrict__ staging,
                                                      const __grid_constant__ UnstageParams p) {
  const UNIT *src = staging + (uint64_t)blockIdx.y * p.sub;
  const uint64_t fixed = p.fixed[blockIdx.y];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (uint64_t j0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < p.sub; j0 += stride * U) {
    UNIT v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      if (j < p.sub) v[u] = src[j];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t j = j0 + u * stride;
      if (j < p.sub) psi[insert_zeros(j, p.ins) | fixed] = v[u];
    }
  }
}

// ------------------------------------------------------------------------------------------ host structures
struct ShOp {
  int kind;
  std::vector<int> q;        // logical qubits
  std::vector<double> data;  // matrix / diagonal / phase, as the C ABI takes them
};

struct Shard {
  int rank = -1;
  bool local = false;
  // valid in this process for every attached shard
  char *data = nullptr;       // slice
  char *staging = nullptr;    // staging area (aux allocation + kFlagBytes)
  uint64_t *flags = nullptr;  // [F_KINDS][world], then the error word
  // local shards only
  State *st = nullptr;
  char *aux = nullptr;
  cudaStream_t xs = nullptr;  // copy stream (ordering of the pushes of one slab)
  cudaStream_t us = nullptr;  // unstage stream (copy-engine unstage)
  cudaStream_t px[8] = {};    // push streams: the copies of one slab fan out over several copy engines
  cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
  unsigned push_rr = 0;
  cudaEvent_t ev[F_KINDS][kEvRing] = {};
  cudaEvent_t ev_pass[kEvRing] = {};  // compute stream -> copy stream: "pass on slab i done"
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  bool ipc_data = false, ipc_aux = false;  // mapped through CUDA IPC (close on destroy)
  // profiling (first local shard only): CUDA-event pairs around passes, exchange regions, pushes, unstages
  struct Rec { int kind; cudaEvent_t a, b; };
  std::vector<Rec> prof;
  std::vector<cudaEvent_t> pool;
  size_t pool_used = 0;
};
enum { PR_PASS = 0, PR_REGION = 1, PR_PUSH = 2, PR_UNSTAGE = 3, PR_SLAB_PASS = 4, PR_INPLACE = 5, PR_KINDS = 6 };

struct Sharded {
  int n = 0, nl = 0, gbits = 0, world = 1, precision = B200SV_F64;
  std::vector<Shard> sh;        // by rank
  std::vector<int> local;       // ranks hosted by this process
  std::vector<int> phys;        // logical qubit -> physical position (>= nl: selects the shard)
  size_t staging_bytes = 0;
  uint64_t seq[F_KINDS] = {0, 0, 0, 0};
  int min_run_bits = 20;
  int want_slab_bits = 3;
  bool allow_staged = true;
  bool unstage_dma = false;   // move received slabs into place with the copy engines instead of a kernel
  int push_streams = 4;
  int max_ride = 2;           // passes per side of an exchange that may run slab-wise under it
  bool profile = false;
  bool emulate = false;       // executor self-test: shards are host arrays, work is interpreted in issue order
  // statistics of the last run
  int64_t stat_passes = 0, stat_exchanges = 0, stat_staged = 0, stat_inplace = 0, stat_launches = 0, stat_copies = 0;
  int64_t stat_overlapped_passes = 0;
  double stat_bytes_exchanged = 0;  // per shard, one direction
  size_t amp_bytes() const { return precision == B200SV_F64 ? 16 : 8; }
  bool multi_process() const { return (int)local.size() < world; }
};

static void sel(const Shard &s) {
  if (s.st->selftest_host) return;  // emulated shard: no device
  B200_CUDA(cudaSetDevice(s.st->device));
}

// profiling scope: records an event pair on `stream` around a piece of work of the first local shard
struct ProfScope {
  Shard *m = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  int kind = 0;
  static cudaEvent_t take(Shard &m) {
    if (m.pool_used == m.pool.size()) {
      cudaEvent_t e;
      B200_CUDA(cudaEventCreate(&e));
      m.pool.push_back(e);
    }
    return m.pool[m.pool_used++];
  }
  ProfScope(Sharded &S, Shard &sh, cudaStream_t st, int k) {
    if (!S.profile || sh.rank != S.local[0]) return;
    m = &sh; stream = st; kind = k;
    sel(sh);
    a = take(sh); b = take(sh);
    B200_CUDA(cudaEventRecord(a, stream));
  }
  void end() {
    if (!m) return;
    cudaSetDevice(m->st->device);
    cudaEventRecord(b, stream);
    m->prof.push_back({kind, a, b});
    m = nullptr;
  }
  ~ProfScope() { end(); }
};

// ------------------------------------------------------------------------------------------ cross-shard ordering
// signal: "everything issued so far on `stream` of local shard `me` is done" becomes visible under (kind, seq) to the
// shards in `to`.  wait: `stream` of local shard `me` does not proceed before every shard in `from` has signalled
// (kind, >= seq).  Shards of this process are ordered with events (the host issues their work in dependency order),
// shards of other processes with flag words.
static void signal(Sharded &S, int me, cudaStream_t stream, int kind, uint64_t seq, const std::vector<int> &to) {
  if (S.emulate) return;  // the emulation executes in issue order, which is a valid serialisation of the partial order
  Shard &m = S.sh[me];
  sel(m);
  B200_CUDA(cudaEventRecord(m.ev[kind][seq % kEvRing], stream));
  SignalParams p;
  p.n = 0;
  p.value = seq;
  for (int r : to) {
    if (r == me || S.sh[r].local) continue;
    if (!S.sh[r].flags) throw Error("sharded: shard " + std::to_string(r) + " is not attached");
    p.target[p.n++] = S.sh[r].flags + (size_t)kind * S.world + me;
  }
  if (p.n) {
    shard_signal_kernel<<<1, 64, 0, stream>>>(p);
    B200_CUDA(cudaGetLastError());
    S.stat_launches++;
  }
}
static void wait(Sharded &S, int me, cudaStream_t stream, int kind, uint64_t seq, const std::vector<int> &from) {
  if (seq == 0 || S.emulate) return;
  Shard &m = S.sh[me];
  sel(m);
  WaitParams p;
  p.n = 0;
  p.value = seq;
  p.err = reinterpret_cast<uint32_t *>(m.flags + (size_t)F_KINDS * S.world);
  static const double env_timeout = [] { const char *e = getenv("B200SV_SHARD_TIMEOUT_S"); return e ? atof(e) : 120.0; }();
  p.timeout_ns = (unsigned long long)(env_timeout * 1e9);
  for (int r : from) {
    if (S.sh[r].local) {
      B200_CUDA(cudaStreamWaitEvent(stream, S.sh[r].ev[kind][seq % kEvRing], 0));
    } else {
      p.flag[p.n++] = m.flags + (size_t)kind * S.world + r;
    }
  }
  if (p.n) {
    shard_wait_kernel<<<1, 64, 0, stream>>>(p);
    B200_CUDA(cudaGetLastError());
    S.stat_launches++;
  }
}

// ------------------------------------------------------------------------------------------ exchange geometry
struct Exchange {
  int k = 0;
  int lpos[4], gbit[4];
};
// the 2^k shards that trade sub-blocks with `rank`: by value v of the k global bits
static int peer_of(const Exchange &x, int rank, uint32_t v) {
  int r = rank;
  for (int i = 0; i < x.k; i++) r = (r & ~(1 << x.gbit[i])) | (int)((v >> i) & 1u) << x.gbit[i];
  return r;
}
static uint32_t gval_of(const Exchange &x, int rank) {
  uint32_t g = 0;
  for (int i = 0; i < x.k; i++) g |= (uint32_t)((rank >> x.gbit[i]) & 1) << i;
  return g;
}
static std::vector<int> group_of(const Exchange &x, int rank, bool with_self) {
  std::vector<int> g;
  for (uint32_t v = 0; v < (1u << x.k); v++) {
    const int r = peer_of(x, rank, v);
    if (with_self || r != rank) g.push_back(r);
  }
  return g;
}

static std::vector<int> all_others(const Sharded &S, int rank) {
  std::vector<int> g;
  for (int r = 0; r < S.world; r++)
    if (r != rank) g.push_back(r);
  return g;
}

// DMA push of one (slab, receiver) sub-block: the amplitudes whose exchange positions read `lval` and whose slab bits
// read `slab`, in index order, into a contiguous destination.  Runs of 2^F0 amplitudes; runs that repeat with a fixed
// stride go out as one 2-D copy.
static void push_subblock(Sharded &S, Shard &m, const std::vector<int> &fixed_pos /*sorted*/, uint64_t fixed_mask,
                          char *compact, bool scatter = false, cudaStream_t only = nullptr) {
  const size_t ab = S.amp_bytes();
  const int nf = (int)fixed_pos.size();
  const int free_bits = S.nl - nf;
  const int f0 = fixed_pos[0];
  if (f0 == 0) throw Error("sharded: exchange position 0 cannot be pushed in runs");
  const uint64_t run = 1ull << f0, run_bytes = run * ab;
  const uint64_t nruns = 1ull << (free_bits - f0);
  InsertList ins;
  ins.n = nf;
  for (int i = 0; i < nf; i++) ins.pos[i] = (uint8_t)fixed_pos[i];
  // runs r .. r + g - 1 share a stride of 2 runs when the next fixed position is not adjacent
  uint64_t group = 1;
  if (nf >= 2) group = 1ull << (fixed_pos[1] - f0 - 1);
  else group = nruns;
  group = std::min(group, nruns);
  const bool use_2d = group > 1 && 2 * run_bytes < (1ull << 31);
  // a 2-D copy that spans a whole group is split so that every push stream gets a share
  uint64_t rows = use_2d ? group : 1;
  const int nstreams = only ? 1 : std::max(1, S.push_streams);
  if (use_2d && nruns / group < (uint64_t)nstreams) rows = std::max<uint64_t>(1, group * (nruns / group) / nstreams);
  for (uint64_t r = 0; r < nruns;) {
    const uint64_t in_group = use_2d ? std::min(rows, group - (r % group)) : 1;
    const uint64_t idx = insert_zeros(r << f0, ins) | fixed_mask;
    char *strided = m.data + idx * ab;
    char *packed = compact + (r << f0) * ab;
    if (S.emulate) {
      for (uint64_t g = 0; g < in_group; g++) {
        char *sp = strided + g * 2 * run_bytes, *pp = packed + g * run_bytes;
        if (scatter) memcpy(sp, pp, run_bytes);
        else memcpy(pp, sp, run_bytes);
      }
      S.stat_copies++;
      r += in_group;
      continue;
    }
    cudaStream_t st = only ? only : m.px[m.push_rr++ % nstreams];
    if (use_2d && in_group > 1) {
      if (scatter) B200_CUDA(cudaMemcpy2DAsync(strided, 2 * run_bytes, packed, run_bytes, run_bytes, in_group, cudaMemcpyDeviceToDevice, st));
      else B200_CUDA(cudaMemcpy2DAsync(packed, run_bytes, strided, 2 * run_bytes, run_bytes, in_group, cudaMemcpyDeviceToDevice, st));
    } else {
      if (scatter) B200_CUDA(cudaMemcpyAsync(strided, packed, run_bytes, cudaMemcpyDeviceToDevice, st));
      else B200_CUDA(cudaMemcpyAsync(packed, strided, run_bytes, cudaMemcpyDeviceToDevice, st));
    }
    S.stat_copies++;
    r += in_group;
  }
}

// ------------------------------------------------------------------------------------------ program
struct Step {
  int type = 0;                     // 0: tile passes (per-shard plans), 1: direct op, 2: exchange
  std::vector<TilePlan *> plans;    // type 0: by rank (every rank is planned, also the ones hosted elsewhere)
  int op = -1;                      // type 1
  std::vector<int> pq;              // type 1: physical qubits
  Exchange x;                       // type 2
};
struct Program {
  std::vector<Step> steps;
  ~Program() {
    for (auto &s : steps)
      for (auto *p : s.plans) tile_plan_free(p);
  }
};

// gates waiting for the next tile segment, per rank
struct RankQueue {
  std::vector<int> nq;
  std::vector<uint64_t> qubits;
  std::vector<double> mats;
  void push(int k, const int *q, const std::complex<double> *m) {
    nq.push_back(k);
    qubits.push_back((uint64_t)q[0]);
    qubits.push_back(k == 2 ? (uint64_t)q[1] : 0);
    const size_t off = mats.size();
    mats.resize(off + 32, 0.0);
    memcpy(&mats[off], m, ((size_t)1 << (2 * k)) * sizeof(std::complex<double>));
  }
};
typedef std::complex<double> cd;

// Rewrite op (physical qubits pq; positions >= nl are resolved from `rank`) as a dense 1-/2-qubit gate for the tile
// queue.  Returns 0: rides the queue (or is the identity on this shard), 1: needs its own kernel.
static int enqueue_op(const Sharded &S, const ShOp &op, const std::vector<int> &pq, int rank, RankQueue &Q) {
  const int nl = S.nl;
  auto bit_of = [&](int p) { return (rank >> (p - nl)) & 1; };
  const int k = (int)pq.size();
  const cd *M = reinterpret_cast<const cd *>(op.data.data());
  if (op.kind == OP_MATRIX) {
    if (k > 2) return 1;
    Q.push(k, pq.data(), M);
    return 0;
  }
  if (op.kind == OP_DIAGONAL) {
    // restrict to the local qubits given this shard's global bits (chunk_utils.hpp:82-118)
    std::vector<int> lq, lbit;
    uint64_t fixed = 0;
    for (int j = 0; j < k; j++) {
      if (pq[j] < nl) { lq.push_back(pq[j]); lbit.push_back(j); }
      else if (bit_of(pq[j])) fixed |= 1ull << j;
    }
    if (lq.size() > 2) return 1;
    auto entry = [&](uint64_t i) {
      uint64_t src = fixed;
      for (size_t b = 0; b < lbit.size(); b++)
        if ((i >> b) & 1) src |= 1ull << lbit[b];
      return M[src];
    };
    if (lq.empty()) {  // a scalar on this shard: {d, d} on position 0
      const cd d = entry(0);
      const cd m2[4] = {d, 0, 0, d};
      const int q0 = 0;
      Q.push(1, &q0, m2);
      return 0;
    }
    const int kk = (int)lq.size(), dim = 1 << kk;
    cd full[16] = {};
    for (int i = 0; i < dim; i++) full[i + dim * i] = entry((uint64_t)i);
    Q.push(kk, lq.data(), full);
    return 0;
  }
  // controlled families: global controls are resolved from the shard index
  const int ntgt = op.kind == OP_MCSWAP ? 2 : (op.kind == OP_MCPHASE ? 0 : 1);
  std::vector<int> ctrl, tgt;
  for (int j = 0; j < k; j++) (j < k - ntgt ? ctrl : tgt).push_back(pq[j]);
  bool mcu_diag = false;
  if (op.kind == OP_MCU) mcu_diag = M[1] == 0.0 && M[2] == 0.0;
  std::vector<int> lctrl;
  for (int c : ctrl) {
    if (c < nl) lctrl.push_back(c);
    else if (!bit_of(c)) return 0;  // identity on this shard
  }
  if (op.kind == OP_MCPHASE) {  // every listed local qubit is a control of the phase
    if (lctrl.size() > 2) return 1;
    const cd ph(op.data[0], op.data[1]);
    if (lctrl.empty()) { const cd m2[4] = {ph, 0, 0, ph}; const int q0 = 0; Q.push(1, &q0, m2); return 0; }
    const int kk = (int)lctrl.size(), dim = 1 << kk;
    cd full[16] = {};
    for (int i = 0; i < dim; i++) full[i + dim * i] = 1.0;
    full[dim * dim - 1] = ph;
    Q.push(kk, lctrl.data(), full);
    return 0;
  }
  if (op.kind == OP_MCU && mcu_diag && tgt[0] >= nl) {  // diagonal on a global target: a phase under the local controls
    const cd ph = bit_of(tgt[0]) ? M[3] : M[0];
    if (lctrl.size() > 2) return 1;
    if (lctrl.empty()) { const cd m2[4] = {ph, 0, 0, ph}; const int q0 = 0; Q.push(1, &q0, m2); return 0; }
    const int kk = (int)lctrl.size(), dim = 1 << kk;
    cd full[16] = {};
    for (int i = 0; i < dim; i++) full[i + dim * i] = 1.0;
    full[dim * dim - 1] = ph;
    Q.push(kk, lctrl.data(), full);
    return 0;
  }
  for (int t : tgt)
    if (t >= nl) throw Error("sharded: a non-diagonal target is on a global qubit (planner bug)");
  if ((int)lctrl.size() + ntgt > 2) return 1;
  cd u[4];
  if (op.kind == OP_MCX) { u[0] = 0; u[1] = 1; u[2] = 1; u[3] = 0; }
  else if (op.kind == OP_MCY) { u[0] = 0; u[1] = cd(0, 1); u[2] = cd(0, -1); u[3] = 0; }
  else if (op.kind == OP_MCU) { for (int i = 0; i < 4; i++) u[i] = M[i]; }
  if (op.kind == OP_MCSWAP) {
    static const cd SWAP[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    if (!lctrl.empty()) return 1;
    Q.push(2, tgt.data(), SWAP);
    return 0;
  }
  if (lctrl.empty()) { Q.push(1, tgt.data(), u); return 0; }
  // one local control: 4x4 with U in the control = 1 block; matrix bit 0 <-> control, column major
  cd CU[16] = {};
  CU[0] = 1; CU[2 + 4 * 2] = 1;
  for (int tr = 0; tr < 2; tr++)
    for (int tc = 0; tc < 2; tc++) CU[(1 + 2 * tr) + 4 * (1 + 2 * tc)] = u[tr + 2 * tc];
  const int q2[2] = {lctrl[0], tgt[0]};
  Q.push(2, q2, CU);
  return 0;
}

static void run_direct(Sharded &S, Shard &m, const ShOp &op, const std::vector<int> &pq) {
  if (S.emulate) throw Error("sharded self-test: ops that need their own kernel are not emulated");
  std::vector<uint64_t> q(pq.begin(), pq.end());
  b200sv_handle h = (b200sv_handle)m.st;
  int rc = 0;
  switch (op.kind) {
  case OP_MATRIX: rc = b200sv_apply_matrix(h, q.data(), (int)q.size(), op.data.data()); break;
  case OP_DIAGONAL: rc = b200sv_apply_diagonal(h, q.data(), (int)q.size(), op.data.data()); break;
  case OP_MCX: rc = b200sv_apply_mcx(h, q.data(), (int)q.size()); break;
  case OP_MCY: rc = b200sv_apply_mcy(h, q.data(), (int)q.size()); break;
  case OP_MCPHASE: rc = b200sv_apply_mcphase(h, q.data(), (int)q.size(), op.data[0], op.data[1]); break;
  case OP_MCSWAP: rc = b200sv_apply_mcswap(h, q.data(), (int)q.size()); break;
  case OP_MCU: rc = b200sv_apply_mcu(h, q.data(), (int)q.size(), op.data.data()); break;
  default: throw Error("sharded: unknown op kind");
  }
  if (rc) throw Error(std::string("sharded: ") + b200sv_last_error());
  S.stat_launches++;
}

// Compile: epoch plan (planner.cu) -> steps.  Queueable ops between two exchanges / direct ops become one tile step
// with a plan per rank.
static std::unique_ptr<Program> compile(Sharded &S, const std::vector<ShOp> &ops) {
  NvtxRange nvtx("b200sv sharded compile (epoch plan + tile plans)");
  const int nops = (int)ops.size();
  std::vector<int> off(1, 0), qs;
  std::vector<uint8_t> need;
  for (const ShOp &op : ops) {
    const int k = (int)op.q.size();
    for (int j = 0; j < k; j++) {
      qs.push_back(op.q[j]);
      bool nl_needed = false;
      switch (op.kind) {
      case OP_MATRIX: nl_needed = true; break;
      case OP_DIAGONAL: case OP_MCPHASE: nl_needed = false; break;
      case OP_MCSWAP: nl_needed = j >= k - 2; break;
      case OP_MCU: {
        const cd *M = reinterpret_cast<const cd *>(op.data.data());
        nl_needed = j == k - 1 && !(M[1] == 0.0 && M[2] == 0.0);
        break;
      }
      default: nl_needed = j == k - 1; break;  // mcx / mcy: target
      }
      need.push_back(nl_needed ? 1 : 0);
    }
    off.push_back((int)qs.size());
  }
  if (qs.empty()) { qs.push_back(0); need.push_back(0); }
  std::vector<int64_t> plan;
  plan_epochs(S.n, S.nl, S.gbits, nops, off.data(), qs.data(), need.data(), std::min(S.min_run_bits, std::max(S.nl - 1, 0)),
              true, S.phys.data(), plan);
  std::unique_ptr<Program> prog(new Program());
  std::vector<RankQueue> Q(S.world);
  State proto;  // planning needs the slice geometry only
  proto.nq = S.nl;
  proto.nstates = 1;
  proto.precision = S.precision;
  auto flush = [&] {
    bool any = false;
    for (auto &q : Q) any = any || !q.nq.empty();
    if (!any) return;
    Step st;
    st.type = 0;
    st.plans.resize(S.world, nullptr);
    for (int r = 0; r < S.world; r++) {
      st.plans[r] = tile_plan_build(proto, (int)Q[r].nq.size(), Q[r].nq.data(), Q[r].qubits.data(), Q[r].mats.data());
      Q[r] = RankQueue();
    }
    prog->steps.push_back(std::move(st));
  };
  for (size_t i = 0; i < plan.size();) {
    if (plan[i] == 0) {
      const ShOp &op = ops[(size_t)plan[i + 1]];
      const int k = (int)plan[i + 2];
      std::vector<int> pq(k);
      for (int j = 0; j < k; j++) pq[j] = (int)plan[i + 3 + j];
      bool direct = false;
      for (int r = 0; r < S.world && !direct; r++) {
        RankQueue tmp;
        if (enqueue_op(S, op, pq, r, tmp)) direct = true;
      }
      if (direct) {
        flush();
        Step st;
        st.type = 1;
        st.op = (int)plan[i + 1];
        st.pq = pq;
        prog->steps.push_back(std::move(st));
      } else {
        for (int r = 0; r < S.world; r++) enqueue_op(S, op, pq, r, Q[r]);
      }
      i += 3 + k;
    } else {
      flush();
      Step st;
      st.type = 2;
      if (plan[i] == 1) {
        st.x.k = 1;
        st.x.lpos[0] = (int)plan[i + 1];
        st.x.gbit[0] = (int)plan[i + 2];
        i += 3;
      } else {
        const int k = (int)plan[i + 1];
        if (k > 4) throw Error("sharded: more than 4 simultaneous global swaps");
        st.x.k = k;
        for (int j = 0; j < k; j++) { st.x.lpos[j] = (int)plan[i + 2 + j]; st.x.gbit[j] = (int)plan[i + 2 + k + j]; }
        i += 2 + 2 * k;
      }
      prog->steps.push_back(std::move(st));
    }
  }
  flush();
  return prog;
}

// ------------------------------------------------------------------------------------------ execution
// `count` passes at the end (last = true: they run before an exchange) or at the start of a tile step, taken along by
// the exchange next to them and run slab by slab
struct PassRef {
  const Step *step = nullptr;  // tile step the passes belong to (null: none)
  bool last = false;
  int count = 0;
  int index(int rank, int j) const { return last ? tile_plan_passes(step->plans[rank]) - count + j : j; }
  bool valid() const { return step != nullptr && count > 0; }
};

static void launch_pass(Sharded &S, Shard &m, const Step &st, int pass, const SlabSpec *slab) {
  sel(m);
  tile_plan_launch(*m.st, st.plans[m.rank], pass, slab);
  S.stat_launches++;
}

// host version of the in-place all-to-all (self-test): amplitude with exchange positions l on the shard whose global
// bits read g trades places with the amplitude whose positions read g on the shard whose bits read l
template <typename C> static void emulate_inplace(Sharded &S, const Exchange &x) {
  std::vector<int> pos(x.lpos, x.lpos + x.k);
  std::sort(pos.begin(), pos.end());
  InsertList ins;
  ins.n = x.k;
  for (int i = 0; i < x.k; i++) ins.pos[i] = (uint8_t)pos[i];
  auto lmask = [&](uint32_t v) {
    uint64_t m = 0;
    for (int i = 0; i < x.k; i++)
      if ((v >> i) & 1) m |= 1ull << x.lpos[i];
    return m;
  };
  const uint64_t count = 1ull << (S.nl - x.k);
  for (int me = 0; me < S.world; me++) {
    const uint32_t g = gval_of(x, me);
    for (uint32_t l = 0; l < (1u << x.k); l++) {
      const int peer = peer_of(x, me, l);
      if (peer <= me) continue;  // each unordered pair of sub-blocks once
      C *a = reinterpret_cast<C *>(S.sh[me].data), *b = reinterpret_cast<C *>(S.sh[peer].data);
      for (uint64_t j = 0; j < count; j++) {
        const uint64_t base = insert_zeros(j, ins);
        std::swap(a[base | lmask(l)], b[base | lmask(g)]);
      }
    }
  }
}

static void exchange_inplace(Sharded &S, const Exchange &x) {
  NvtxRange nvtx("b200sv exchange (in place)");
  if (S.emulate) {
    if (S.precision == B200SV_F64) emulate_inplace<std::complex<double>>(S, x);
    else emulate_inplace<std::complex<float>>(S, x);
    S.stat_inplace++;
    return;
  }
  const uint64_t s_ready = ++S.seq[F_READY], s_done = ++S.seq[F_DONE];
  for (int me : S.local) signal(S, me, S.sh[me].st->stream, F_READY, s_ready, group_of(x, me, false));
  for (int me : S.local) {
    Shard &m = S.sh[me];
    wait(S, me, m.st->stream, F_READY, s_ready, group_of(x, me, false));
    void *peers[16] = {};
    const uint32_t my_g = gval_of(x, me);
    for (uint32_t v = 0; v < (1u << x.k); v++) {
      const int r = peer_of(x, me, v);
      if (!S.sh[r].data) throw Error("sharded: shard " + std::to_string(r) + " is not attached");
      peers[v] = S.sh[r].data;
    }
    sel(m);
    launch_multi_swap_peer(*m.st, x.k, x.lpos, my_g, peers);
    S.stat_launches++;
    signal(S, me, m.st->stream, F_DONE, s_done, group_of(x, me, false));
  }
  for (int me : S.local) wait(S, me, S.sh[me].st->stream, F_DONE, s_done, group_of(x, me, false));
  S.stat_inplace++;
}

// slab bits for a staged exchange: positions >= min_run_bits outside the exchange and outside the neighbouring passes
static std::vector<int> pick_slab_bits(const Sharded &S, const Exchange &x, const PassRef &before, const PassRef &after,
                                       int want) {
  uint64_t used = 0;
  for (int i = 0; i < x.k; i++) used |= 1ull << x.lpos[i];
  for (int r = 0; r < S.world; r++) {
    for (int j = 0; before.valid() && j < before.count; j++) used |= tile_plan_pass_mask(before.step->plans[r], before.index(r, j));
    for (int j = 0; after.valid() && j < after.count; j++) used |= tile_plan_pass_mask(after.step->plans[r], after.index(r, j));
  }
  std::vector<int> bits;
  const int lo = std::min(S.min_run_bits, std::max(S.nl - 1, 1));
  for (int p = S.nl - 1; p >= lo && (int)bits.size() < want; p--)
    if (!((used >> p) & 1)) bits.push_back(p);
  std::sort(bits.begin(), bits.end());
  return bits;
}

static void exchange_staged(Sharded &S, const Exchange &x, const PassRef &before, const PassRef &after,
                            const std::vector<int> &slab_bits, int nbuf) {
  NvtxRange nvtx("b200sv exchange (staged, slab pipeline)");
  const int s = (int)slab_bits.size(), nslab = 1 << s, k = x.k;
  const size_t ab = S.amp_bytes();
  const uint64_t sub = 1ull << (S.nl - s - k);            // amplitudes per (slab, sender) sub-block
  const size_t slot_bytes = sub * ab, buf_bytes = slot_bytes * ((1u << k) - 1);
  std::vector<int> fixed_pos(x.lpos, x.lpos + k);
  fixed_pos.insert(fixed_pos.end(), slab_bits.begin(), slab_bits.end());
  std::sort(fixed_pos.begin(), fixed_pos.end());
  auto fixed_mask = [&](uint32_t lval, int slab) {
    uint64_t m = 0;
    for (int i = 0; i < k; i++)
      if ((lval >> i) & 1) m |= 1ull << x.lpos[i];
    for (int b = 0; b < s; b++)
      if ((slab >> b) & 1) m |= 1ull << slab_bits[b];
    return m;
  };
  SlabSpec spec;
  spec.nbits = s;
  for (int b = 0; b < s; b++) spec.pos[b] = slab_bits[b];
  const uint64_t c0_pushed = S.seq[F_PUSHED], c0_unstaged = S.seq[F_UNSTAGED];
  // Software pipeline over the slabs, `nbuf` slabs ahead on the sending side (the staging area holds nbuf slabs):
  //   compute stream:  before(0) .. before(nbuf-1) | before(i+nbuf), unstage(i), after(i)   for i = 0, 1, ...
  //   copy streams:    push(s) after before(s) and after the receivers have unstaged slab s - nbuf
  // so the compute stream always has work that does not depend on the link (the "before" passes of a later slab)
  // while slab i is still in flight.
  auto issue_before = [&](int i) {  // the passes before the exchange on slab i; completion handed to the copy stream
    for (int me : S.local) {
      Shard &m = S.sh[me];
      for (int j = 0; before.valid() && j < before.count; j++) {
        spec.value = (uint32_t)i;
        ProfScope ps(S, m, m.st->stream, PR_SLAB_PASS);
        launch_pass(S, m, *before.step, before.index(me, j), s ? &spec : nullptr);
        if (i == 0) S.stat_overlapped_passes++;
      }
      sel(m);
      if (!S.emulate) B200_CUDA(cudaEventRecord(m.ev_pass[i % kEvRing], m.st->stream));
    }
  };
  auto issue_push = [&](int i) {    // pushes of slab i (copy engines), once the receivers' buffer is free again
    const int b = i % nbuf;
    for (int me : S.local) {
      Shard &m = S.sh[me];
      sel(m);
      if (!S.emulate) B200_CUDA(cudaStreamWaitEvent(m.xs, m.ev_pass[i % kEvRing], 0));
      const uint64_t need_unstaged = c0_unstaged + (uint64_t)std::max(0, i - nbuf + 1);
      wait(S, me, m.xs, F_UNSTAGED, need_unstaged, group_of(x, me, false));
      const uint32_t my_g = gval_of(x, me);
      ProfScope ps_push(S, m, m.xs, PR_PUSH);
      if (!S.emulate) {
        B200_CUDA(cudaEventRecord(m.ev_fork, m.xs));
        for (int j = 0; j < S.push_streams; j++) B200_CUDA(cudaStreamWaitEvent(m.px[j], m.ev_fork, 0));
      }
      // receivers in XOR order: at step d every shard g sends to g ^ d, a perfect matching
      for (uint32_t dstep = 1; dstep < (1u << k); dstep++) {
        const uint32_t v = my_g ^ dstep;
        const Shard &peer = S.sh[peer_of(x, me, v)];
        if (!peer.staging) throw Error("sharded: shard " + std::to_string(peer.rank) + " is not attached");
        const uint32_t slot = my_g < v ? my_g : my_g - 1;  // the receiver (id v) skips its own id
        push_subblock(S, m, fixed_pos, fixed_mask(v, i), peer.staging + (size_t)b * buf_bytes + (size_t)slot * slot_bytes);
      }
      for (int j = 0; j < S.push_streams && !S.emulate; j++) {
        B200_CUDA(cudaEventRecord(m.ev_join[j], m.px[j]));
        B200_CUDA(cudaStreamWaitEvent(m.xs, m.ev_join[j], 0));
      }
      ps_push.end();
      signal(S, me, m.xs, F_PUSHED, c0_pushed + i + 1, group_of(x, me, true));
    }
  };
  auto issue_unstage_after = [&](int i) {  // move slab i into place, then the passes after the exchange on it
    const int b = i % nbuf;
    for (int me : S.local) {
      Shard &m = S.sh[me];
      sel(m);
      const uint32_t my_g = gval_of(x, me);
      const char *src = m.staging + (size_t)b * buf_bytes;
      if (S.unstage_dma || S.emulate) {
        // copy engines again (unstage stream): the compute stream only waits for the result (the self-test takes
        // this branch too: the same strided copies, as memcpy)
        wait(S, me, m.us, F_PUSHED, c0_pushed + i + 1, group_of(x, me, true));
        ProfScope ps_un(S, m, m.us, PR_UNSTAGE);
        for (uint32_t u = 0, slot = 0; u < (1u << k); u++) {
          if (u == my_g) continue;
          push_subblock(S, m, fixed_pos, fixed_mask(u, i), const_cast<char *>(src) + (size_t)slot * slot_bytes, true, m.us);
          slot++;
        }
        ps_un.end();
        signal(S, me, m.us, F_UNSTAGED, c0_unstaged + i + 1, all_others(S, me));
        wait(S, me, m.st->stream, F_UNSTAGED, c0_unstaged + i + 1, std::vector<int>(1, me));
      } else {
        wait(S, me, m.st->stream, F_PUSHED, c0_pushed + i + 1, group_of(x, me, true));
        UnstageParams up;
        up.sub = sub;
        up.nslots = (1 << k) - 1;
        up.ins.n = (int)fixed_pos.size();
        for (size_t u = 0; u < fixed_pos.size(); u++) up.ins.pos[u] = (uint8_t)fixed_pos[u];
        for (uint32_t u = 0, slot = 0; u < (1u << k); u++) {
          if (u == my_g) continue;
          up.fixed[slot++] = fixed_mask(u, i);
        }
        const unsigned gx = (unsigned)std::min<uint64_t>((sub + 1023) / 1024, std::max<uint64_t>(1, (uint64_t)m.st->num_sms * 8 / up.nslots));
        dim3 grid(std::max(gx, 1u), (unsigned)up.nslots);
        ProfScope ps_un(S, m, m.st->stream, PR_UNSTAGE);
        if (S.precision == B200SV_F64) unstage_kernel<uint4><<<grid, 256, 0, m.st->stream>>>((uint4 *)m.data, (const uint4 *)src, up);
        else unstage_kernel<uint2><<<grid, 256, 0, m.st->stream>>>((uint2 *)m.data, (const uint2 *)src, up);
        B200_CUDA(cudaGetLastError());
        ps_un.end();
        S.stat_launches++;
        // to every shard: the next exchange may pair this shard with different partners
        signal(S, me, m.st->stream, F_UNSTAGED, c0_unstaged + i + 1, all_others(S, me));
      }
      for (int j = 0; after.valid() && j < after.count; j++) {
        spec.value = (uint32_t)i;
        ProfScope ps(S, m, m.st->stream, PR_SLAB_PASS);
        launch_pass(S, m, *after.step, after.index(me, j), s ? &spec : nullptr);
        if (i == 0) S.stat_overlapped_passes++;
      }
    }
  };
  for (int i = 0; i < std::min(nbuf, nslab); i++) { issue_before(i); issue_push(i); }
  for (int i = 0; i < nslab; i++) {
    if (i + nbuf < nslab) issue_before(i + nbuf);
    issue_unstage_after(i);
    if (i + nbuf < nslab) issue_push(i + nbuf);
  }
  S.seq[F_PUSHED] = c0_pushed + nslab;
  S.seq[F_UNSTAGED] = c0_unstaged + nslab;
  S.stat_staged++;
}

// How each exchange of a program runs -- decided from the plans of ALL ranks and the handle's configuration only, so
// that every process of a multi-process register takes the same decisions.
struct XSched {
  bool staged = false;
  int n_before = 0, n_after = 0;   // passes of the neighbouring tile steps that ride along, slab by slab
  std::vector<int> slab_bits;
  int nbuf = 1;
};
static std::vector<XSched> schedule(const Sharded &S, const Program &prog, std::vector<int> &first_taken,
                                    std::vector<int> &last_taken) {
  const size_t ns = prog.steps.size();
  std::vector<XSched> xs(ns);
  first_taken.assign(ns, 0);
  last_taken.assign(ns, 0);
  for (size_t si = 0; si < ns; si++) {
    const Step &st = prog.steps[si];
    if (st.type != 2) continue;
    const Exchange &x = st.x;
    XSched &d = xs[si];
    d.staged = S.allow_staged && S.staging_bytes > 0;
    for (int i = 0; i < x.k; i++)
      if (x.lpos[i] == 0) d.staged = false;  // runs of one amplitude: not a DMA shape
    if (!d.staged) continue;
    // how many passes are there to take along on either side (identical decision on every rank: all plans are known)
    int avail_before = 0, avail_after = 0;
    if (si > 0 && prog.steps[si - 1].type == 0) {
      avail_before = 1 << 30;
      for (int r = 0; r < S.world; r++)
        avail_before = std::min(avail_before, tile_plan_passes(prog.steps[si - 1].plans[r]) - first_taken[si - 1]);
    }
    if (si + 1 < ns && prog.steps[si + 1].type == 0) {
      const bool next_next_x = si + 2 < ns && prog.steps[si + 2].type == 2;
      avail_after = 1 << 30;
      for (int r = 0; r < S.world; r++)  // keep one pass for the next exchange's "before" side
        avail_after = std::min(avail_after, tile_plan_passes(prog.steps[si + 1].plans[r]) - (next_next_x ? 1 : 0));
    }
    avail_before = std::max(0, std::min(avail_before, S.max_ride));
    avail_after = std::max(0, std::min(avail_after, S.max_ride));
    const double out_bytes = (double)(1ull << S.nl) * S.amp_bytes() * (1.0 - 1.0 / (1 << x.k));
    auto fits = [&](size_t nbits) { return out_bytes / (double)(1u << nbits) <= (double)S.staging_bytes; };
    // widest overlap window first; a window needs at least two slabs to be worth its slab launches
    bool chosen = false;
    for (int total = avail_before + avail_after; total >= 1 && !chosen; total--)
      for (int nb = std::min(avail_before, total); nb >= 0 && !chosen; nb--) {
        const int na = total - nb;
        if (na > avail_after || na < 0) continue;
        PassRef before, after;
        if (nb) { before.step = &prog.steps[si - 1]; before.last = true; before.count = nb; }
        if (na) { after.step = &prog.steps[si + 1]; after.last = false; after.count = na; }
        std::vector<int> bits = pick_slab_bits(S, x, before, after, S.want_slab_bits);
        const size_t need_bits = total >= 3 ? 2 : 1;
        if (bits.size() < need_bits || !fits(bits.size())) continue;
        d.slab_bits = bits;
        d.n_before = nb;
        d.n_after = na;
        chosen = true;
      }
    if (!chosen) {
      // nothing rides along: pipeline pushes against unstages only, with as many slabs as the staging area needs
      d.slab_bits = pick_slab_bits(S, x, PassRef(), PassRef(), 0);
      if (!fits(0)) d.slab_bits = pick_slab_bits(S, x, PassRef(), PassRef(), 4);
      if (!fits(d.slab_bits.size())) { d.staged = false; continue; }
    }
    d.nbuf = (2.0 * out_bytes / (double)(1u << d.slab_bits.size()) <= (double)S.staging_bytes) ? 2 : 1;
    if (d.n_before) last_taken[si - 1] = d.n_before;
    if (d.n_after) first_taken[si + 1] = d.n_after;
  }
  return xs;
}

static void run_program(Sharded &S, const Program &prog, const std::vector<ShOp> &ops) {
  const size_t ns = prog.steps.size();
  std::vector<int> first_taken, last_taken;  // tile steps: how many of their first / last passes ride an exchange
  const std::vector<XSched> xs = schedule(S, prog, first_taken, last_taken);
  for (size_t si = 0; si < ns; si++) {
    const Step &st = prog.steps[si];
    if (st.type == 1) {
      for (int me : S.local) { sel(S.sh[me]); run_direct(S, S.sh[me], ops[st.op], st.pq); }
      continue;
    }
    if (st.type == 0) {
      for (int me : S.local) {
        const int np = tile_plan_passes(st.plans[me]);
        for (int p = first_taken[si]; p < np - last_taken[si]; p++) {
          ProfScope ps(S, S.sh[me], S.sh[me].st->stream, PR_PASS);
          launch_pass(S, S.sh[me], st, p, nullptr);
        }
        if (me == S.local[0]) S.stat_passes += np;
      }
      continue;
    }
    const Exchange &x = st.x;
    const XSched &d = xs[si];
    S.stat_exchanges++;
    S.stat_bytes_exchanged += (double)(1ull << S.nl) * (1.0 - 1.0 / (1 << x.k)) * S.amp_bytes();
    Shard &first = S.sh[S.local[0]];
    if (d.staged) {
      PassRef before, after;
      if (d.n_before) { before.step = &prog.steps[si - 1]; before.last = true; before.count = d.n_before; }
      if (d.n_after) { after.step = &prog.steps[si + 1]; after.last = false; after.count = d.n_after; }
      ProfScope ps(S, first, first.st->stream, PR_REGION);
      exchange_staged(S, x, before, after, d.slab_bits, d.nbuf);
    } else {
      ProfScope ps(S, first, first.st->stream, PR_INPLACE);
      exchange_inplace(S, x);
    }
  }
}

// ------------------------------------------------------------------------------------------ lifetime
static void destroy(Sharded *S) {
  if (!S) return;
  for (Shard &m : S->sh) {
    if (m.local && m.st) {
      cudaSetDevice(m.st->device);
      cudaStreamSynchronize(m.st->stream);
      if (m.xs) { cudaStreamSynchronize(m.xs); cudaStreamDestroy(m.xs); }
      if (m.us) { cudaStreamSynchronize(m.us); cudaStreamDestroy(m.us); }
      for (int j = 0; j < 8; j++) {
        if (m.px[j]) { cudaStreamSynchronize(m.px[j]); cudaStreamDestroy(m.px[j]); }
        if (m.ev_join[j]) cudaEventDestroy(m.ev_join[j]);
      }
      if (m.ev_fork) cudaEventDestroy(m.ev_fork);
      for (int kd = 0; kd < F_KINDS; kd++)
        for (int i = 0; i < kEvRing; i++)
          if (m.ev[kd][i]) cudaEventDestroy(m.ev[kd][i]);
      for (int i = 0; i < kEvRing; i++)
        if (m.ev_pass[i]) cudaEventDestroy(m.ev_pass[i]);
      for (cudaEvent_t e : m.pool) cudaEventDestroy(e);
      if (m.ev_t0) cudaEventDestroy(m.ev_t0);
      if (m.ev_t1) cudaEventDestroy(m.ev_t1);
      if (m.aux) cudaFree(m.aux);
      b200sv_destroy((b200sv_handle)m.st);
    } else if (!m.local) {
      if (m.ipc_data && m.data) cudaIpcCloseMemHandle(m.data);
      if (m.ipc_aux && m.flags) cudaIpcCloseMemHandle(m.flags);
    }
  }
  cudaGetLastError();
  delete S;
}

static void restore_order_impl(Sharded &S) {
  const int n = S.n, nl = S.nl;
  std::vector<int> phys = S.phys, inv(n);
  auto reinv = [&] { for (int q = 0; q < n; q++) inv[phys[q]] = q; };
  Program prog;
  std::vector<ShOp> ops;
  auto add_x = [&](int lpos, int gbit) {
    Step st;
    st.type = 2;
    st.x.k = 1;
    st.x.lpos[0] = lpos;
    st.x.gbit[0] = gbit;
    prog.steps.push_back(std::move(st));
  };
  // global positions: when every logical qubit that belongs on a global position currently sits on a LOCAL one, a single
  // k-qubit all-to-all brings them all home ((1 - 2^-k) of the slice crosses the links once instead of k half-slice
  // exchanges); otherwise one pairwise exchange per position
  {
    std::vector<int> want;
    bool all_local = true;
    for (int g = nl; g < n; g++)
      if (phys[g] != g) { want.push_back(g); all_local = all_local && phys[g] < nl; }
    if (want.size() >= 2 && want.size() <= 4 && all_local) {
      reinv();
      Step st;
      st.type = 2;
      st.x.k = (int)want.size();
      for (size_t i = 0; i < want.size(); i++) {
        const int g = want[i], lpos = phys[g], occupant = inv[g];
        st.x.lpos[i] = lpos;
        st.x.gbit[i] = g - nl;
        phys[occupant] = lpos;
        phys[g] = g;
      }
      prog.steps.push_back(std::move(st));
    }
  }
  for (int g = nl; g < n; g++) {
    reinv();
    if (inv[g] == g) continue;
    int p = phys[g];  // where logical qubit g lives now
    if (p >= nl) {    // on another global position: pull it to a local one first
      const int lpos = nl - 1, victim = inv[lpos];
      add_x(lpos, p - nl);
      phys[victim] = p;
      phys[g] = lpos;
      p = lpos;
      reinv();
    }
    const int occupant = inv[g];
    add_x(p, g - nl);
    phys[occupant] = p;
    phys[g] = g;
  }
  // the local permutation as SWAP gates riding tile passes (~9 transpositions per HBM pass instead of one mcswap
  // pass each; 0/1 matrices: the amplitudes move bit-exactly)
  std::vector<RankQueue> Q(S.world);
  static const cd SWAPM[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
  for (int q = 0; q < nl; q++) {
    const int p = phys[q];
    if (p == q) continue;
    reinv();
    const int other = inv[q];
    const int pq[2] = {p, q};
    for (int r = 0; r < S.world; r++) Q[r].push(2, pq, SWAPM);
    phys[q] = q;
    phys[other] = p;
  }
  if (!Q[0].nq.empty()) {
    State proto;
    proto.nq = S.nl;
    proto.nstates = 1;
    proto.precision = S.precision;
    Step st;
    st.type = 0;
    st.plans.resize(S.world, nullptr);
    for (int r = 0; r < S.world; r++)
      st.plans[r] = tile_plan_build(proto, (int)Q[r].nq.size(), Q[r].nq.data(), Q[r].qubits.data(), Q[r].mats.data());
    prog.steps.push_back(std::move(st));
  }
  run_program(S, prog, ops);
  S.phys = phys;
}

}  // namespace b200sv

using namespace b200sv;

template <typename F> static int sguard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return 1;
  }
}
#define SH ((Sharded *)h)

extern "C" {

int b200sv_sharded_create(b200sv_sharded_handle *out, int num_qubits, int precision, int world, int nlocal,
                          const int *local_ranks, const int *devices, uint64_t staging_bytes) {
  return sguard([&] {
    if (!out || world < 1 || world > kMaxWorld || (world & (world - 1))) throw Error("sharded_create: the number of shards must be a power of two (<= 64)");
    int g = 0;
    while ((1 << g) < world) g++;
    if (num_qubits - g < 1 || num_qubits > 62) throw Error("sharded_create: num_qubits out of range");
    if (nlocal < 1 || nlocal > world || !local_ranks || !devices) throw Error("sharded_create: bad local shard list");
    std::unique_ptr<Sharded, void (*)(Sharded *)> S(new Sharded(), destroy);
    S->n = num_qubits; S->gbits = g; S->nl = num_qubits - g; S->world = world; S->precision = precision;
    S->sh.resize(world);
    for (int r = 0; r < world; r++) S->sh[r].rank = r;
    S->phys.resize(num_qubits);
    for (int q = 0; q < num_qubits; q++) S->phys[q] = q;
    if (const char *e = getenv("B200SV_SHARD_MIN_RUN_BITS")) S->min_run_bits = atoi(e);
    if (const char *e = getenv("B200SV_SHARD_SLAB_BITS")) S->want_slab_bits = std::max(0, std::min(4, atoi(e)));
    if (const char *e = getenv("B200SV_SHARD_STAGED")) S->allow_staged = atoi(e) != 0;
    if (const char *e = getenv("B200SV_SHARD_UNSTAGE")) S->unstage_dma = !strcmp(e, "dma");
    if (const char *e = getenv("B200SV_SHARD_MAX_RIDE")) S->max_ride = std::max(0, std::min(4, atoi(e)));
    if (const char *e = getenv("B200SV_SHARD_PUSH_STREAMS")) S->push_streams = std::max(1, std::min(8, atoi(e)));
    S->min_run_bits = std::max(1, std::min(S->min_run_bits, std::max(S->nl - 1, 1)));
    const size_t slice_bytes = ((size_t)1 << S->nl) * S->amp_bytes();
    for (int i = 0; i < nlocal; i++) {
      const int r = local_ranks[i];
      if (r < 0 || r >= world || S->sh[r].local) throw Error("sharded_create: bad / duplicate local rank");
      Shard &m = S->sh[r];
      b200sv_handle hh = nullptr;
      if (b200sv_create(&hh, S->nl, 1, precision, devices[i])) throw Error(b200sv_last_error());
      m.st = (State *)hh;
      m.local = true;
      m.data = (char *)m.st->data;
      if (b200sv_set_chunk(hh, num_qubits, (uint64_t)r)) throw Error(b200sv_last_error());
      S->local.push_back(r);
    }
    // staging area: an explicit size, or what the device can spare next to the slice (same rule on every rank, so
    // that all processes take the same staged / in-place decisions): 92 % of the device minus the slice minus 2 GiB,
    // at most the slice itself
    if (staging_bytes == (uint64_t)-1) {
      double best = (double)slice_bytes;
      for (int r : S->local) {
        const int dev = S->sh[r].st->device;
        int on_dev = 0;
        for (int r2 : S->local) on_dev += S->sh[r2].st->device == dev;
        size_t f = 0, t = 0;
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaMemGetInfo(&f, &t));
        const double spare = 0.92 * (double)t - (double)slice_bytes * on_dev - (double)(2ull << 30);
        best = std::min(best, std::max(0.0, spare / on_dev));
      }
      staging_bytes = (uint64_t)best & ~(uint64_t)4095;
    }
    if (const char *e = getenv("B200SV_SHARD_STAGING_MB")) staging_bytes = (uint64_t)atoll(e) << 20;
    S->staging_bytes = world > 1 ? (size_t)staging_bytes : 0;
    for (int r : S->local) {
      Shard &m = S->sh[r];
      B200_CUDA(cudaSetDevice(m.st->device));
      m.aux = (char *)device_alloc(m.st->device, kFlagBytes + S->staging_bytes, false);
      B200_CUDA(cudaMemset(m.aux, 0, kFlagBytes));
      m.flags = (uint64_t *)m.aux;
      m.staging = m.aux + kFlagBytes;
      B200_CUDA(cudaStreamCreateWithFlags(&m.xs, cudaStreamNonBlocking));
      B200_CUDA(cudaStreamCreateWithFlags(&m.us, cudaStreamNonBlocking));
      for (int j = 0; j < S->push_streams; j++) {
        B200_CUDA(cudaStreamCreateWithFlags(&m.px[j], cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&m.ev_join[j], cudaEventDisableTiming));
      }
      B200_CUDA(cudaEventCreateWithFlags(&m.ev_fork, cudaEventDisableTiming));
      for (int kd = 0; kd < F_KINDS; kd++)
        for (int i = 0; i < kEvRing; i++) B200_CUDA(cudaEventCreateWithFlags(&m.ev[kd][i], cudaEventDisableTiming));
      for (int i = 0; i < kEvRing; i++) B200_CUDA(cudaEventCreateWithFlags(&m.ev_pass[i], cudaEventDisableTiming));
      B200_CUDA(cudaEventCreate(&m.ev_t0));
      B200_CUDA(cudaEventCreate(&m.ev_t1));
      B200_CUDA(cudaDeviceSynchronize());
    }
    // peer access between the devices of this process (chunk_manager.hpp:129-135)
    for (int a : S->local)
      for (int b : S->local) {
        const int da = S->sh[a].st->device, db = S->sh[b].st->device;
        if (da == db) continue;
        int can = 0;
        B200_CUDA(cudaDeviceCanAccessPeer(&can, da, db));
        if (!can) throw Error("sharded_create: GPU " + std::to_string(da) + " cannot access GPU " + std::to_string(db));
        B200_CUDA(cudaSetDevice(da));
        cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) B200_CUDA(e);
        cudaGetLastError();
      }
    *out = (b200sv_sharded_handle)S.release();
  });
}

int b200sv_sharded_destroy(b200sv_sharded_handle h) {
  return sguard([&] { destroy(SH); });
}

int b200sv_sharded_ipc_export(b200sv_sharded_handle h, int rank, void *blob128) {
  return sguard([&] {
    if (!h || rank < 0 || rank >= SH->world || !SH->sh[rank].local) throw Error("sharded_ipc_export: not a local shard");
    Shard &m = SH->sh[rank];
    B200_CUDA(cudaSetDevice(m.st->device));
    B200_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)blob128, m.data));
    B200_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)((char *)blob128 + 64), m.aux));
  });
}

int b200sv_sharded_ipc_attach(b200sv_sharded_handle h, int rank, const void *blob128) {
  return sguard([&] {
    if (!h || rank < 0 || rank >= SH->world || SH->sh[rank].local) throw Error("sharded_ipc_attach: bad rank");
    Shard &m = SH->sh[rank];
    B200_CUDA(cudaSetDevice(SH->sh[SH->local[0]].st->device));
    cudaIpcMemHandle_t hd, ha;
    memcpy(&hd, blob128, 64);
    memcpy(&ha, (const char *)blob128 + 64, 64);
    void *pd = nullptr, *pa = nullptr;
    B200_CUDA(cudaIpcOpenMemHandle(&pd, hd, cudaIpcMemLazyEnablePeerAccess));
    B200_CUDA(cudaIpcOpenMemHandle(&pa, ha, cudaIpcMemLazyEnablePeerAccess));
    m.data = (char *)pd;
    m.flags = (uint64_t *)pa;
    m.staging = (char *)pa + kFlagBytes;
    m.ipc_data = m.ipc_aux = true;
  });
}

int b200sv_sharded_shard_handle(b200sv_sharded_handle h, int rank, b200sv_handle *out) {
  return sguard([&] {
    if (!h || rank < 0 || rank >= SH->world || !SH->sh[rank].local) throw Error("sharded_shard_handle: not a local shard");
    *out = (b200sv_handle)SH->sh[rank].st;
  });
}

int b200sv_sharded_initialize(b200sv_sharded_handle h) {
  return sguard([&] {
    for (int q = 0; q < SH->n; q++) SH->phys[q] = q;
    for (int r : SH->local) {
      B200_CUDA(cudaSetDevice(SH->sh[r].st->device));
      launch_init(*SH->sh[r].st, r == 0);
    }
  });
}

int b200sv_sharded_synchronize(b200sv_sharded_handle h) {
  return sguard([&] {
    for (int r : SH->local) {
      Shard &m = SH->sh[r];
      B200_CUDA(cudaSetDevice(m.st->device));
      B200_CUDA(cudaStreamSynchronize(m.xs));
      B200_CUDA(cudaStreamSynchronize(m.us));
      B200_CUDA(cudaStreamSynchronize(m.st->stream));
      uint32_t err = 0;
      B200_CUDA(cudaMemcpy(&err, m.flags + (size_t)F_KINDS * SH->world, 4, cudaMemcpyDeviceToHost));
      if (err) throw Error("sharded: a partner shard did not arrive at an exchange (timeout)");
    }
  });
}

static std::vector<ShOp> parse_ops(int n, int nops, const int *kinds, const int *op_off, const int *op_qubits,
                                   const int64_t *data_off, const double *data) {
  if (nops < 0 || (nops > 0 && (!kinds || !op_off || !op_qubits || !data_off))) throw Error("sharded ops: bad arguments");
  std::vector<ShOp> ops(nops);
  for (int i = 0; i < nops; i++) {
    ShOp &op = ops[i];
    op.kind = kinds[i];
    if (op.kind < OP_MATRIX || op.kind > OP_MCU) throw Error("sharded ops: unknown op kind");
    op.q.assign(op_qubits + op_off[i], op_qubits + op_off[i + 1]);
    for (size_t a = 0; a < op.q.size(); a++) {
      if (op.q[a] < 0 || op.q[a] >= n) throw Error("sharded ops: qubit out of range");
      for (size_t b = 0; b < a; b++)
        if (op.q[a] == op.q[b]) throw Error("sharded ops: duplicate qubit");
    }
    const int k = (int)op.q.size();
    size_t want = 0;
    switch (op.kind) {
    case OP_MATRIX: want = (size_t)2 << (2 * k); break;
    case OP_DIAGONAL: want = (size_t)2 << k; break;
    case OP_MCPHASE: want = 2; break;
    case OP_MCU: want = 8; break;
    default: want = 0; break;
    }
    if ((size_t)(data_off[i + 1] - data_off[i]) != want) throw Error("sharded ops: wrong data size for op " + std::to_string(i));
    if (k < 1 || (op.kind == OP_MCSWAP && k < 2) || (op.kind == OP_MATRIX && k > kMaxDenseQubits)) throw Error("sharded ops: bad qubit count");
    if (want) op.data.assign(data + data_off[i], data + data_off[i + 1]);
  }
  return ops;
}

int b200sv_sharded_apply_ops(b200sv_sharded_handle h, int nops, const int *kinds, const int *op_off, const int *op_qubits,
                             const int64_t *data_off, const double *data) {
  return sguard([&] {
    if (!h) throw Error("sharded_apply_ops: null handle");
    std::vector<ShOp> ops = parse_ops(SH->n, nops, kinds, op_off, op_qubits, data_off, data);
    SH->stat_passes = SH->stat_exchanges = SH->stat_staged = SH->stat_inplace = SH->stat_launches = SH->stat_copies = 0;
    SH->stat_overlapped_passes = 0;
    SH->stat_bytes_exchanged = 0;
    std::unique_ptr<Program> prog = compile(*SH, ops);
    for (int r : SH->local) { sel(SH->sh[r]); B200_CUDA(cudaEventRecord(SH->sh[r].ev_t0, SH->sh[r].st->stream)); }
    run_program(*SH, *prog, ops);
    for (int r : SH->local) { sel(SH->sh[r]); B200_CUDA(cudaEventRecord(SH->sh[r].ev_t1, SH->sh[r].st->stream)); }
  });
}

// Host-only view of what apply_ops would do (no device): the compile + schedule steps on a register description.
// out8 = {tile passes of shard 0, exchanges, staged, in place, passes taken along by exchanges, finest slab count,
// coarsest slab count, qubit swaps}.  Test and sizing aid (CPU test-suite; "how many passes / exchanges will QV-36 on
// 8 GPUs take?").
int b200sv_sharded_plan_only(int num_qubits, int precision, int world, uint64_t staging_bytes, int nops, const int *kinds,
                             const int *op_off, const int *op_qubits, const int64_t *data_off, const double *data,
                             double *out8) {
  return sguard([&] {
    if (world < 1 || world > kMaxWorld || (world & (world - 1))) throw Error("sharded_plan_only: bad world");
    Sharded S;
    int g = 0;
    while ((1 << g) < world) g++;
    S.n = num_qubits; S.gbits = g; S.nl = num_qubits - g; S.world = world; S.precision = precision;
    if (S.nl < 1) throw Error("sharded_plan_only: num_qubits out of range");
    S.sh.resize(world);
    S.phys.resize(num_qubits);
    for (int q = 0; q < num_qubits; q++) S.phys[q] = q;
    if (const char *e = getenv("B200SV_SHARD_MIN_RUN_BITS")) S.min_run_bits = atoi(e);
    if (const char *e = getenv("B200SV_SHARD_SLAB_BITS")) S.want_slab_bits = std::max(0, std::min(4, atoi(e)));
    if (const char *e = getenv("B200SV_SHARD_STAGED")) S.allow_staged = atoi(e) != 0;
    if (const char *e = getenv("B200SV_SHARD_MAX_RIDE")) S.max_ride = std::max(0, std::min(4, atoi(e)));
    S.min_run_bits = std::max(1, std::min(S.min_run_bits, std::max(S.nl - 1, 1)));
    S.staging_bytes = world > 1 ? (size_t)staging_bytes : 0;
    std::vector<ShOp> ops = parse_ops(num_qubits, nops, kinds, op_off, op_qubits, data_off, data);
    std::unique_ptr<Program> prog = compile(S, ops);
    std::vector<int> ft, lt;
    const std::vector<XSched> xs = schedule(S, *prog, ft, lt);
    double passes = 0, nx = 0, staged = 0, inplace = 0, taken = 0, fine = 0, coarse = 1e9, swaps = 0;
    for (size_t si = 0; si < prog->steps.size(); si++) {
      const Step &st = prog->steps[si];
      if (st.type == 0) passes += tile_plan_passes(st.plans[0]);
      if (st.type != 2) continue;
      nx++;
      swaps += st.x.k;
      if (xs[si].staged) {
        staged++;
        taken += xs[si].n_before + xs[si].n_after;
        fine = std::max(fine, (double)(1u << xs[si].slab_bits.size()));
        coarse = std::min(coarse, (double)(1u << xs[si].slab_bits.size()));
      } else inplace++;
    }
    out8[0] = passes; out8[1] = nx; out8[2] = staged; out8[3] = inplace; out8[4] = taken; out8[5] = fine;
    out8[6] = staged ? coarse : 0; out8[7] = swaps;
  });
}

// Executor self-test -- TEST INFRASTRUCTURE, no device involved: the register lives in `host_state` (2^n amplitudes,
// shard r = the r-th contiguous slice), the program is compiled and scheduled exactly as apply_ops does and then
// INTERPRETED in issue order: tile-pass parameter blocks (whole and slab-patched) by the tile engine's host
// interpreter, pushes / unstages as the same strided copies with memcpy, in-place exchanges as host swaps; finally the
// qubit order is restored (exchanges + SWAP passes).  Covers, on the CPU test tier, the index algebra of slabs, staging
// slots and exchange groups and the issue order of the pipeline (a buffer reused too early shows up as wrong data).
int b200sv_sharded_selftest(int num_qubits, int precision, int world, uint64_t staging_bytes, int nops, const int *kinds,
                            const int *op_off, const int *op_qubits, const int64_t *data_off, const double *data,
                            void *host_state, double *out8) {
  return sguard([&] {
    if (world < 1 || world > kMaxWorld || (world & (world - 1)) || !host_state) throw Error("sharded_selftest: bad arguments");
    Sharded S;
    int g = 0;
    while ((1 << g) < world) g++;
    S.n = num_qubits; S.gbits = g; S.nl = num_qubits - g; S.world = world; S.precision = precision;
    if (S.nl < 4 || S.nl > 26) throw Error("sharded_selftest: 4..26 qubits per shard");
    S.emulate = true;
    S.sh.resize(world);
    S.phys.resize(num_qubits);
    for (int q = 0; q < num_qubits; q++) S.phys[q] = q;
    if (const char *e = getenv("B200SV_SHARD_MIN_RUN_BITS")) S.min_run_bits = atoi(e);
    if (const char *e = getenv("B200SV_SHARD_SLAB_BITS")) S.want_slab_bits = std::max(0, std::min(4, atoi(e)));
    if (const char *e = getenv("B200SV_SHARD_STAGED")) S.allow_staged = atoi(e) != 0;
    if (const char *e = getenv("B200SV_SHARD_MAX_RIDE")) S.max_ride = std::max(0, std::min(4, atoi(e)));
    S.min_run_bits = std::max(1, std::min(S.min_run_bits, std::max(S.nl - 1, 1)));
    S.staging_bytes = world > 1 ? (size_t)staging_bytes : 0;
    const size_t slice_bytes = ((size_t)1 << S.nl) * S.amp_bytes();
    std::vector<State> states(world);
    std::vector<std::vector<char>> staging(world);
    for (int r = 0; r < world; r++) {
      Shard &m = S.sh[r];
      m.rank = r;
      m.local = true;
      states[r].nq = S.nl;
      states[r].precision = precision;
      states[r].global_nq = num_qubits;
      states[r].chunk_index = (uint64_t)r;
      states[r].selftest_host = (char *)host_state + (size_t)r * slice_bytes;
      m.st = &states[r];
      m.data = (char *)states[r].selftest_host;
      staging[r].resize(S.staging_bytes + 16);
      m.staging = staging[r].data();
      S.local.push_back(r);
    }
    std::vector<ShOp> ops = parse_ops(num_qubits, nops, kinds, op_off, op_qubits, data_off, data);
    {
      std::unique_ptr<Program> prog = compile(S, ops);
      run_program(S, *prog, ops);
    }
    if (out8) {
      out8[0] = (double)S.stat_passes; out8[1] = (double)S.stat_exchanges; out8[2] = (double)S.stat_staged;
      out8[3] = (double)S.stat_inplace; out8[4] = 0; out8[5] = (double)S.stat_copies; out8[6] = S.stat_bytes_exchanged;
      out8[7] = (double)S.stat_overlapped_passes;
    }
    // restore the qubit order with the same machinery (what b200sv_sharded_restore_order does on a handle)
    restore_order_impl(S);
    for (int r = 0; r < world; r++) S.sh[r].st = nullptr;  // the States are stack objects
    S.sh.clear();
  });
}

int b200sv_sharded_stats(b200sv_sharded_handle h, double *out8) {
  return sguard([&] {
    out8[0] = (double)SH->stat_passes;
    out8[1] = (double)SH->stat_exchanges;
    out8[2] = (double)SH->stat_staged;
    out8[3] = (double)SH->stat_inplace;
    out8[4] = (double)SH->stat_launches;
    out8[5] = (double)SH->stat_copies;
    out8[6] = SH->stat_bytes_exchanged;
    out8[7] = (double)SH->stat_overlapped_passes;
  });
}

// Profiling of the first local shard: on = 1 starts collecting CUDA-event pairs (cleared), on = 0 stops.  After a
// synchronize, b200sv_sharded_profile_read sums them: out12 = {count, ms} for whole-state tile passes, exchange
// regions (staged: from the first slab pass before to the last slab pass after), pushes (copy stream), unstage
// kernels, slab passes, in-place exchanges.
int b200sv_sharded_profile(b200sv_sharded_handle h, int on) {
  return sguard([&] {
    Shard &m = SH->sh[SH->local[0]];
    B200_CUDA(cudaSetDevice(m.st->device));
    B200_CUDA(cudaDeviceSynchronize());
    m.prof.clear();
    m.pool_used = 0;
    SH->profile = on != 0;
  });
}
int b200sv_sharded_profile_read(b200sv_sharded_handle h, double *out12) {
  return sguard([&] {
    Shard &m = SH->sh[SH->local[0]];
    B200_CUDA(cudaSetDevice(m.st->device));
    B200_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 2 * PR_KINDS; i++) out12[i] = 0.0;
    for (const Shard::Rec &r : m.prof) {
      float ms = 0;
      B200_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      out12[2 * r.kind] += 1.0;
      out12[2 * r.kind + 1] += ms;
    }
  });
}

int b200sv_sharded_qubit_map(b200sv_sharded_handle h, int *phys) {
  return sguard([&] { for (int q = 0; q < SH->n; q++) phys[q] = SH->phys[q]; });
}

int b200sv_sharded_elapsed_ms(b200sv_sharded_handle h, double *ms) {
  return sguard([&] {
    double worst = 0;
    for (int r : SH->local) {
      Shard &m = SH->sh[r];
      B200_CUDA(cudaSetDevice(m.st->device));
      B200_CUDA(cudaEventSynchronize(m.ev_t1));
      float t = 0;
      B200_CUDA(cudaEventElapsedTime(&t, m.ev_t0, m.ev_t1));
      worst = std::max(worst, (double)t);
    }
    *ms = worst;
  });
}

// Bring every logical qubit back to its own physical position (what CacheBlocking's restore_qubit_map does before
// measure / save ops so that chunked and unchunked runs sample identically; test/terra/backends/aer_simulator/
// test_chunk.py:31-168 asserts exact equality): global positions first (pairwise exchanges), then the local
// permutation by transpositions (mcswap passes).
int b200sv_sharded_restore_order(b200sv_sharded_handle h) {
  return sguard([&] { restore_order_impl(*SH); });
}

// per-shard squared norms (entries of shards hosted elsewhere stay 0: sum / gather over processes)
int b200sv_sharded_norms(b200sv_sharded_handle h, double *out_world) {
  return sguard([&] {
    for (int r = 0; r < SH->world; r++) out_world[r] = 0.0;
    for (int r : SH->local) {
      sel(SH->sh[r]);
      reduce_norm(*SH->sh[r].st, &out_world[r]);
    }
  });
}

// Executor::expval_pauli over chunks (statevector_executor.hpp:551-720): Pauli factors on local positions go to the
// shard kernels, Z factors on global positions become a sign per shard, X / Y factors on global positions pair each
// shard with the shard whose global bits differ by the X mask (the pair's slice is read in place over peer access --
// no exchange).  Returns the contribution of the shards hosted by this process (sum over processes).  Every shard's
// earlier work must be complete (b200sv_sharded_synchronize; plus a process barrier in the multi-process layout).
int b200sv_sharded_expval_pauli(b200sv_sharded_handle h, const uint64_t *qubits, int k, const char *pauli, double *partial) {
  return sguard([&] {
    Sharded &S = *SH;
    if (!pauli || (int)strlen(pauli) != k) throw Error("Pauli string length must equal the number of qubits");
    std::vector<uint64_t> q_in;
    std::string p_in_rev;                      // factors on local positions, in qubit order
    uint64_t gx = 0, gz = 0;                   // masks over the shard index
    int num_y = 0;
    for (int i = 0; i < k; i++) {
      if (qubits[i] >= (uint64_t)S.n) throw Error("qubit index out of range");
      const char c = pauli[k - 1 - i];
      const int p = S.phys[qubits[i]];
      if (p < S.nl) { q_in.push_back((uint64_t)p); p_in_rev.push_back(c); continue; }
      const uint64_t bit = 1ull << (p - S.nl);
      switch (c) {
      case 'I': break;
      case 'X': gx |= bit; break;
      case 'Z': gz |= bit; break;
      case 'Y': gx |= bit; gz |= bit; num_y++; break;
      default: throw Error(std::string("Invalid Pauli \"") + c + "\".");
      }
    }
    std::string p_in(p_in_rev.rbegin(), p_in_rev.rend());
    double pre = 1.0, pim = 0.0;               // add_y_phase (qubitvector.hpp:2275-2298) of the global factors
    switch (num_y & 3) {
    case 1: pre = 0; pim = -1; break;
    case 2: pre = -1; pim = 0; break;
    case 3: pre = 0; pim = 1; break;
    default: break;
    }
    double total = 0.0;
    for (int r : S.local) {
      Shard &m = S.sh[r];
      b200sv_handle hh = (b200sv_handle)m.st;
      double v = 0.0;
      if (gx) {
        const int pair = r ^ (int)gx;
        if (r > pair) continue;  // each pair once, by its lower shard
        if (!S.sh[pair].data) throw Error("sharded: shard " + std::to_string(pair) + " is not attached");
        const uint64_t zc = (uint64_t)__builtin_popcountll((uint64_t)r & gz), zcp = (uint64_t)__builtin_popcountll((uint64_t)pair & gz);
        if (b200sv_expval_pauli_pair(hh, q_in.data(), (int)q_in.size(), p_in.c_str(), S.sh[pair].data, zc, zcp, pre, pim, &v))
          throw Error(b200sv_last_error());
      } else {
        const double sign = (__builtin_popcountll((uint64_t)r & gz) & 1) ? -1.0 : 1.0;
        if (b200sv_expval_pauli(hh, q_in.data(), (int)q_in.size(), p_in.c_str(), 1.0, 0.0, &v)) throw Error(b200sv_last_error());
        v *= sign;
      }
      total += v;
    }
    *partial = total;
  });
}

// Executor::sample_measure over chunks (statevector_executor.hpp:1149-1227): every draw belongs to the shard whose
// cumulative-norm interval contains it; that shard samples it locally.  norms_world = the squared norms of ALL shards
// (b200sv_sharded_norms, gathered over processes).  out[i] = the sampled LOGICAL basis state of draw i if it fell on a
// shard of this process, else 0 (sum over processes).
int b200sv_sharded_sample_measure(b200sv_sharded_handle h, const double *rnds, int64_t shots, const double *norms_world,
                                  uint64_t *out) {
  return sguard([&] {
    Sharded &S = *SH;
    std::vector<double> cum(S.world + 1, 0.0);
    for (int r = 0; r < S.world; r++) cum[r + 1] = cum[r] + norms_world[r];
    for (int64_t i = 0; i < shots; i++) out[i] = 0;
    for (int r : S.local) {
      std::vector<double> mine;
      std::vector<int64_t> where;
      for (int64_t i = 0; i < shots; i++) {
        const bool in = r == S.world - 1 ? rnds[i] >= cum[r] : (rnds[i] >= cum[r] && rnds[i] < cum[r + 1]);
        if (in) { mine.push_back(rnds[i] - cum[r]); where.push_back(i); }
      }
      if (mine.empty()) continue;
      std::vector<uint64_t> smp(mine.size());
      sel(S.sh[r]);
      sample_measure(*S.sh[r].st, mine.data(), (int64_t)mine.size(), smp.data());
      for (size_t j = 0; j < mine.size(); j++) {
        const uint64_t physical = smp[j] | ((uint64_t)r << S.nl);
        uint64_t logical = 0;
        for (int q = 0; q < S.n; q++) logical |= ((physical >> S.phys[q]) & 1ull) << q;
        out[where[j]] = logical;
      }
    }
  });
}

}  // extern "C"
