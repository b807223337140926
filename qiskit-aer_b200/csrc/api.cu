// b200sv C ABI (include/b200sv.h): handle management, argument validation,
// the reference's host-side routing rules (exact-== special cases, Pauli mask
// construction, global-qubit resolution for sharded chunks) and error
// translation.  Kernels live in gates.cu / reduce.cu.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <random>

#include "common.cuh"

namespace b200sv {

static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

// ---- device allocations.  The slice of a destroyed handle (>= 256 MiB) is kept, one block per device, for the next
// handle of the same size: Aer's executors create and destroy the register once per circuit / experiment
// (Executor::run_circuit_*; the reference's ChunkManager likewise keeps its chunks when the shape repeats,
// parallel_state_executor.hpp:318-326 "can reuse allocated chunks"), and a 128 GiB cudaMalloc + cudaFree costs ~0.2 s.
// Every allocation that fails first releases the cache and retries; b200sv_trim() releases it on demand;
// B200SV_ALLOC_CACHE=0 switches it off.
namespace {
struct CachedBlock { void *ptr = nullptr; size_t bytes = 0; };
std::mutex g_cache_mu;
CachedBlock g_cache[64];
CachedBlock g_scratch_cache[64];  // a destroyed handle's scratch buffer (reductions, sampler), per device
CachedBlock g_pinned_cache;       // ... and its pinned host staging (cudaMallocHost / cudaFreeHost are slow and synchronise)
bool cache_enabled() {
  static const bool on = [] { const char *e = getenv("B200SV_ALLOC_CACHE"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace
void trim_alloc_cache() {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 64; d++) {
    if (g_cache[d].ptr) {
      cudaSetDevice(d);
      cudaFree(g_cache[d].ptr);
      g_cache[d] = CachedBlock();
    }
    if (g_scratch_cache[d].ptr) {
      cudaSetDevice(d);
      cudaFree(g_scratch_cache[d].ptr);
      g_scratch_cache[d] = CachedBlock();
    }
  }
  cudaSetDevice(cur);
  cudaGetLastError();
}
void *device_alloc(int device, size_t bytes, bool may_reuse) {
  if (may_reuse && cache_enabled() && device >= 0 && device < 64) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache[device].ptr && g_cache[device].bytes == bytes) {
      void *p = g_cache[device].ptr;
      g_cache[device] = CachedBlock();
      return p;
    }
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    trim_alloc_cache();
    e = cudaMalloc(&p, bytes);
  }
  B200_CUDA(e);
  return p;
}
void device_free(int device, void *p, size_t bytes, bool may_cache) {
  if (!p) return;
  if (may_cache && cache_enabled() && bytes >= ((size_t)256 << 20) && device >= 0 && device < 64) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache[device].ptr) cudaFree(g_cache[device].ptr);
    g_cache[device].ptr = p;
    g_cache[device].bytes = bytes;
    return;
  }
  cudaFree(p);
}

// small buffers of destroyed handles, kept for the next handle (one each)
static void stash_small(State &s) {
  if (!cache_enabled() || s.device < 0 || s.device >= 64) {
    if (s.scratch) cudaFree(s.scratch);
    if (s.pinned) cudaFreeHost(s.pinned);
    s.scratch = s.pinned = nullptr;
    return;
  }
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (s.scratch) {
    CachedBlock &c = g_scratch_cache[s.device];
    if (c.ptr && c.bytes >= s.scratch_bytes) cudaFree(s.scratch);
    else { if (c.ptr) cudaFree(c.ptr); c.ptr = s.scratch; c.bytes = s.scratch_bytes; }
  }
  if (s.pinned) {
    CachedBlock &c = g_pinned_cache;
    if (c.ptr && c.bytes >= s.pinned_bytes) cudaFreeHost(s.pinned);
    else { if (c.ptr) cudaFreeHost(c.ptr); c.ptr = s.pinned; c.bytes = s.pinned_bytes; }
  }
  s.scratch = s.pinned = nullptr;
  s.scratch_bytes = s.pinned_bytes = 0;
}
static bool take_cached(CachedBlock &c, size_t need, void **ptr, size_t *bytes) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (!c.ptr || c.bytes < need) return false;
  *ptr = c.ptr;
  *bytes = c.bytes;
  c = CachedBlock();
  return true;
}

void *State::ensure_scratch(size_t bytes) {
  if (bytes > scratch_bytes && !scratch && cache_enabled() && device >= 0 && device < 64)
    take_cached(g_scratch_cache[device], bytes, &scratch, &scratch_bytes);
  if (bytes > scratch_bytes) {
    if (scratch) {
      B200_CUDA(cudaStreamSynchronize(stream));
      B200_CUDA(cudaFree(scratch));
      scratch = nullptr;
    }
    size_t want = std::max<size_t>(bytes, 1 << 20);
    scratch = device_alloc(device, want, false);
    scratch_bytes = want;
  }
  return scratch;
}
void *State::ensure_pinned(size_t bytes) {
  if (bytes > pinned_bytes && !pinned && cache_enabled()) take_cached(g_pinned_cache, bytes, &pinned, &pinned_bytes);
  if (bytes > pinned_bytes) {
    if (pinned) {
      B200_CUDA(cudaStreamSynchronize(stream));
      B200_CUDA(cudaFreeHost(pinned));
      pinned = nullptr;
    }
    size_t want = std::max<size_t>(bytes, 1 << 16);
    B200_CUDA(cudaMallocHost(&pinned, want));
    pinned_bytes = want;
  }
  return pinned;
}

std::vector<int> checked_qubits(const State &s, const uint64_t *qubits, int k, bool allow_global) {
  if (k < 0 || (k > 0 && !qubits)) throw Error("invalid qubit list");
  std::vector<int> q(k);
  const int limit = allow_global && s.global_nq > s.nq ? s.global_nq : s.nq;
  for (int i = 0; i < k; i++) {
    if (qubits[i] >= (uint64_t)limit) throw Error("qubit index " + std::to_string(qubits[i]) + " out of range");
    q[i] = (int)qubits[i];
    for (int j = 0; j < i; j++)
      if (q[j] == q[i]) throw Error("duplicate qubit " + std::to_string(q[i]));
  }
  return q;
}

// pauli_masks_and_phase + add_y_phase (qubitvector.hpp:2236-2298)
struct PauliMasks {
  uint64_t x = 0, z = 0;
  int num_y = 0, x_max = 0;
};
static PauliMasks pauli_masks(const std::vector<int> &q, const char *pauli) {
  const size_t N = q.size();
  if (!pauli || strlen(pauli) != N) throw Error("Pauli string length must equal the number of qubits");
  PauliMasks m;
  for (size_t i = 0; i < N; i++) {
    const uint64_t bit = 1ull << q[i];
    switch (pauli[N - 1 - i]) {
    case 'I': break;
    case 'X': m.x += bit; m.x_max = std::max(m.x_max, q[i]); break;
    case 'Z': m.z += bit; break;
    case 'Y': m.x += bit; m.x_max = std::max(m.x_max, q[i]); m.z += bit; m.num_y++; break;
    default: throw Error(std::string("Invalid Pauli \"") + pauli[N - 1 - i] + "\".");
    }
  }
  return m;
}
static void add_y_phase(int num_y, double &re, double &im) {
  const double r = re, i = im;
  switch (num_y & 3) {
  case 1: re = i; im = -r; break;
  case 2: re = -r; im = -i; break;
  case 3: re = -i; im = r; break;
  default: break;
  }
}

static void select(State *s) {
  if (!s) throw Error("null handle");
  B200_CUDA(cudaSetDevice(s->device));
}

// Split a control list into local controls and the verdict of global ones
// (chunk_utils.hpp:55-80 / thrust base_index_ masks): returns false when a global
// control bit is 0 for this chunk (gate is the identity here).
static bool resolve_controls(const State &s, std::vector<int> &controls) {
  std::vector<int> local;
  for (int c : controls) {
    if (c < s.nq) local.push_back(c);
    else if (!((s.chunk_index >> (c - s.nq)) & 1ull)) return false;
  }
  controls.swap(local);
  return true;
}

template <typename F> static int guard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return 1;
  }
}

// Peer access for pointers that live on another GPU of this process (chunk_manager.hpp:129-135,
// device_chunk_container.hpp:336-351: cudaDeviceEnablePeerAccess between the devices that hold chunks), enabled
// lazily the first time a handle is handed such a pointer.  IPC-mapped pointers were opened with
// cudaIpcMemLazyEnablePeerAccess and need nothing.
void ensure_peer(const State &s, const void *ptr) {
  if (!ptr) return;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return; }
  if (attr.type != cudaMemoryTypeDevice || attr.device == s.device) return;
  static bool enabled[64][64] = {};
  if (attr.device < 0 || attr.device >= 64 || s.device >= 64 || enabled[s.device][attr.device]) return;
  int can = 0;
  B200_CUDA(cudaDeviceCanAccessPeer(&can, s.device, attr.device));
  if (!can) throw Error("GPU " + std::to_string(s.device) + " cannot access GPU " + std::to_string(attr.device) + " (no peer path)");
  cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);  // current device = s.device (select() ran)
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) B200_CUDA(e);
  cudaGetLastError();
  enabled[s.device][attr.device] = true;
}

// Per-state kernels index the state with blockIdx.y (<= 65535): containers with more states run them slab by slab on
// views of <= 65535 states (the reference sizes shot groups by memory only, chunk_manager.hpp:223-264).
template <typename F> static void for_state_slabs(State *s, F f) {
  constexpr int64_t kMaxY = 65535;
  if (s->nstates <= kMaxY) { f(*s, (int64_t)0); return; }
  for (int64_t s0 = 0; s0 < s->nstates; s0 += kMaxY) {
    State v = *s;
    v.nstates = std::min<int64_t>(kMaxY, s->nstates - s0);
    v.data = (char *)s->data + ((uint64_t)s0 << s->nq) * s->amp_bytes();
    v.owns_data = false;
    if (v.checkpoint) v.checkpoint = (char *)s->checkpoint + ((uint64_t)s0 << s->nq) * s->amp_bytes();
    f(v, s0);
    s->scratch = v.scratch; s->scratch_bytes = v.scratch_bytes;  // the view may have grown the shared buffers
    s->pinned = v.pinned; s->pinned_bytes = v.pinned_bytes;
  }
}

static State *make_state(int nq, int64_t nstates, int precision, int device) {
  if (nq < 0 || nq > 40) throw Error("num_qubits out of range");
  if (nstates < 1) throw Error("num_states must be >= 1");
  if (precision != B200SV_F64 && precision != B200SV_F32) throw Error("precision must be 64 or 32");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error("No CUDA device available! b200sv has no CPU fallback (cf. aer_controller.hpp:306-310)");
  if (device < 0 || device >= count) throw Error("device index out of range");
  State *s = new State();
  s->device = device; s->nq = nq; s->nstates = nstates; s->precision = precision; s->global_nq = nq;
  B200_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    delete s;
    throw Error("b200sv kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
  }
  s->num_sms = prop.multiProcessorCount;
  return s;
}

}  // namespace b200sv

using namespace b200sv;
#define H ((State *)h)

extern "C" {

int b200sv_version(void) { return 1; }
const char *b200sv_last_error(void) { return g_last_error.c_str(); }

int b200sv_device_count(int *count) {
  return guard([&] {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) c = 0;
    *count = c;
  });
}

int b200sv_mem_info(int device, uint64_t *free_bytes, uint64_t *total_bytes) {
  return guard([&] {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) throw Error("mem_info: device index out of range");
    size_t f = 0, t = 0;
    B200_CUDA(cudaSetDevice(device));
    B200_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
  });
}

int b200sv_measure_fp64_peak(int device, double duration_ms, double *burst_tflops, double *sustained_tflops) {
  return guard([&] {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) throw Error("measure_fp64_peak: device index out of range");
    if (!burst_tflops || !sustained_tflops) throw Error("measure_fp64_peak: null output");
    measure_fp64_peak(device, duration_ms, burst_tflops, sustained_tflops);
  });
}

int b200sv_trim(void) { return guard([&] { trim_alloc_cache(); }); }

int b200sv_create(b200sv_handle *out, int num_qubits, int64_t num_states, int precision, int device) {
  return guard([&] {
    State *s = make_state(num_qubits, num_states, precision, device);
    try {
      s->data = device_alloc(device, s->total_amps() * s->amp_bytes(), true);
      s->owns_data = true;
      B200_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
      s->owns_stream = true;
      launch_init(*s, true);
    } catch (...) {
      if (s->data) cudaFree(s->data);
      delete s;
      throw;
    }
    *out = (b200sv_handle)s;
  });
}

int b200sv_create_external(b200sv_handle *out, int num_qubits, int64_t num_states, int precision, int device,
                           void *dev_ptr, void *cuda_stream) {
  return guard([&] {
    if (!dev_ptr || ((uintptr_t)dev_ptr & 15)) throw Error("external device pointer must be non-null and 16-byte aligned");
    // (kernels with 256-bit accesses check for 32-byte alignment themselves and take their 128-bit variants otherwise)
    State *s = make_state(num_qubits, num_states, precision, device);
    s->data = dev_ptr;
    s->stream = (cudaStream_t)cuda_stream;
    *out = (b200sv_handle)s;
  });
}

int b200sv_destroy(b200sv_handle h) {
  return guard([&] {
    if (!h) return;
    select(H);
    cudaStreamSynchronize(H->stream);
    if (H->owns_data && H->data) device_free(H->device, H->data, H->total_amps() * H->amp_bytes(), true);
    stash_small(*H);
    if (H->checkpoint) cudaFree(H->checkpoint);
    if (H->owns_stream && H->stream) cudaStreamDestroy(H->stream);
    delete H;
  });
}

int b200sv_num_qubits(b200sv_handle h, int *n) { return guard([&] { select(H); *n = H->nq; }); }
int b200sv_device_ptr(b200sv_handle h, void **p) { return guard([&] { select(H); *p = H->data; }); }
int b200sv_stream(b200sv_handle h, void **p) { return guard([&] { select(H); *p = (void *)H->stream; }); }

int b200sv_set_chunk(b200sv_handle h, int global_num_qubits, uint64_t chunk_index) {
  return guard([&] {
    select(H);
    if (global_num_qubits < H->nq || global_num_qubits > 63) throw Error("global_num_qubits out of range");
    if (global_num_qubits - H->nq < 63 && (chunk_index >> (global_num_qubits - H->nq)) != 0)
      throw Error("chunk_index out of range");
    H->global_nq = global_num_qubits;
    H->chunk_index = chunk_index;
  });
}

int b200sv_synchronize(b200sv_handle h) {
  return guard([&] { select(H); B200_CUDA(cudaStreamSynchronize(H->stream)); });
}

int b200sv_initialize(b200sv_handle h) { return guard([&] { select(H); launch_init(*H, true); }); }
int b200sv_zero(b200sv_handle h) { return guard([&] { select(H); launch_init(*H, false); }); }

int b200sv_upload(b200sv_handle h, const void *host, uint64_t offset, uint64_t count) {
  return guard([&] {
    select(H);
    if (offset + count > H->total_amps()) throw Error("upload range exceeds the state");
    B200_CUDA(cudaMemcpyAsync((char *)H->data + offset * H->amp_bytes(), host, count * H->amp_bytes(),
                              cudaMemcpyHostToDevice, H->stream));
    B200_CUDA(cudaStreamSynchronize(H->stream));
  });
}
int b200sv_download(b200sv_handle h, void *host, uint64_t offset, uint64_t count) {
  return guard([&] {
    select(H);
    if (offset + count > H->total_amps()) throw Error("download range exceeds the state");
    B200_CUDA(cudaMemcpyAsync(host, (char *)H->data + offset * H->amp_bytes(), count * H->amp_bytes(),
                              cudaMemcpyDeviceToHost, H->stream));
    B200_CUDA(cudaStreamSynchronize(H->stream));
  });
}

int b200sv_download_line(b200sv_handle h, int row_bits, uint64_t xor_mask, void *host_out) {
  return guard([&] {
    select(H);
    if (2 * row_bits != H->nq || H->nstates != 1) throw Error("download_line: the state is not a 2^m x 2^m matrix");
    if (row_bits < 63 && (xor_mask >> row_bits)) throw Error("download_line: mask out of range");
    launch_gather_line(*H, row_bits, xor_mask, host_out);
  });
}

int b200sv_dm_expval_pauli(b200sv_handle h, int row_bits, const uint64_t *qubits, int k, const char *pauli, double pre,
                           double pim, double *out) {
  return guard([&] {
    select(H);
    if (2 * row_bits != H->nq || H->nstates != 1) throw Error("dm_expval_pauli: the state is not a 2^m x 2^m matrix");
    std::vector<int> q(k);
    for (int i = 0; i < k; i++) {
      if (qubits[i] >= (uint64_t)row_bits) throw Error("qubit index " + std::to_string(qubits[i]) + " out of range");
      q[i] = (int)qubits[i];
    }
    PauliMasks m = pauli_masks(q, pauli);
    add_y_phase(m.num_y, pre, pim);
    reduce_dm_expval(*H, row_bits, m.x, m.z, pre, pim, out);  // identity string: the trace (densitymatrix.hpp:463-466)
  });
}
int b200sv_dm_probabilities(b200sv_handle h, int row_bits, const uint64_t *qubits, int k, double *out) {
  return guard([&] {
    select(H);
    if (2 * row_bits != H->nq || H->nstates != 1) throw Error("dm_probabilities: the state is not a 2^m x 2^m matrix");
    std::vector<int> q(k);
    for (int i = 0; i < k; i++) {
      if (qubits[i] >= (uint64_t)row_bits) throw Error("qubit index " + std::to_string(qubits[i]) + " out of range");
      q[i] = (int)qubits[i];
      for (int j = 0; j < i; j++)
        if (q[j] == q[i]) throw Error("duplicate qubit " + std::to_string(q[i]));
    }
    reduce_dm_probabilities(*H, row_bits, q.data(), k, out);
  });
}

int b200sv_initialize_component(b200sv_handle h, const uint64_t *qubits, int k, const double *state) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    launch_init_component(*H, q.data(), k, state);
  });
}

int b200sv_checkpoint(b200sv_handle h) {
  return guard([&] {
    select(H);
    const size_t bytes = H->total_amps() * H->amp_bytes();
    if (!H->checkpoint) H->checkpoint = device_alloc(H->device, bytes, false);
    B200_CUDA(cudaMemcpyAsync(H->checkpoint, H->data, bytes, cudaMemcpyDeviceToDevice, H->stream));
  });
}
int b200sv_revert(b200sv_handle h, int keep) {
  return guard([&] {
    select(H);
    if (!H->checkpoint) throw Error("revert: no checkpoint");
    const size_t bytes = H->total_amps() * H->amp_bytes();
    B200_CUDA(cudaMemcpyAsync(H->data, H->checkpoint, bytes, cudaMemcpyDeviceToDevice, H->stream));
    if (!keep) {
      B200_CUDA(cudaStreamSynchronize(H->stream));
      B200_CUDA(cudaFree(H->checkpoint));
      H->checkpoint = nullptr;
    }
  });
}
int b200sv_inner_product(b200sv_handle h, double *re, double *im) {
  return guard([&] {
    select(H);
    if (!H->checkpoint) throw Error("inner_product: no checkpoint");
    for_state_slabs(H, [&](State &v, int64_t s0) { reduce_inner_product(v, v.checkpoint, re + s0, im + s0); });
  });
}

// ------------------------------------------------------------------ gates
int b200sv_apply_matrix(b200sv_handle h, const uint64_t *qubits, int k, const double *mat) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("apply_matrix: empty qubit list");
    auto q = checked_qubits(*H, qubits, k);
    if (k <= kMaxRegQubits) launch_dense(*H, q.data(), k, nullptr, 0, mat);
    else launch_dense_generic(*H, q.data(), k, mat);
  });
}

int b200sv_apply_diagonal(b200sv_handle h, const uint64_t *qubits, int k, const double *diag) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("apply_diagonal_matrix: empty qubit list");
    auto q = checked_qubits(*H, qubits, k, true);
    // restrict to local qubits given this chunk's global bits (chunk_utils.hpp:82-118 block_diagonal_matrix)
    std::vector<int> lq;
    std::vector<int> lbit;
    uint64_t fixed = 0;
    for (int j = 0; j < k; j++) {
      if (q[j] < H->nq) { lq.push_back(q[j]); lbit.push_back(j); }
      else if ((H->chunk_index >> (q[j] - H->nq)) & 1ull) fixed |= 1ull << j;
    }
    if ((int)lq.size() == k) { launch_diagonal(*H, q.data(), k, diag); return; }
    const int kl = (int)lq.size();
    std::vector<double> d2(2ull << kl);
    for (uint64_t i = 0; i < (1ull << kl); i++) {
      uint64_t src = fixed;
      for (int b = 0; b < kl; b++)
        if ((i >> b) & 1) src |= 1ull << lbit[b];
      d2[2 * i] = diag[2 * src];
      d2[2 * i + 1] = diag[2 * src + 1];
    }
    if (kl == 0) {  // scalar on this chunk: apply as a 1-qubit diagonal {d,d} (cf. global phase, statevector_state.hpp:446-450)
      int q0 = 0;
      double dd[4] = {d2[0], d2[1], d2[0], d2[1]};
      if (H->nq == 0) throw Error("apply_diagonal_matrix: zero-qubit chunk");
      launch_diagonal(*H, &q0, 1, dd);
    } else {
      launch_diagonal(*H, lq.data(), kl, d2.data());
    }
  });
}

int b200sv_apply_diagonal_layer(b200sv_handle h, int ngates, const int *nq, const uint64_t *qubits, const double *diags) {
  return guard([&] {
    select(H);
    if (ngates < 0 || (ngates > 0 && (!nq || !qubits || !diags))) throw Error("apply_diagonal_layer: bad arguments");
    if (H->global_nq > H->nq) throw Error("apply_diagonal_layer: not available on a chunk of a sharded register");
    for (int g = 0; g < ngates; g++) {
      if (nq[g] != 1 && nq[g] != 2) throw Error("apply_diagonal_layer: gates must act on 1 or 2 qubits");
      for (int j = 0; j < nq[g]; j++)
        if (qubits[2 * g + j] >= (uint64_t)H->nq) throw Error("qubit index " + std::to_string(qubits[2 * g + j]) + " out of range");
      if (nq[g] == 2 && qubits[2 * g] == qubits[2 * g + 1]) throw Error("duplicate qubit " + std::to_string(qubits[2 * g]));
    }
    if (ngates == 0) return;
    if (H->nq < 4) {  // tiny registers: gate by gate
      for (int g = 0; g < ngates; g++) {
        int q[2] = {(int)qubits[2 * g], (int)qubits[2 * g + 1]};
        launch_diagonal(*H, q, nq[g], diags + 8 * (size_t)g);
      }
      return;
    }
    launch_diag_layer(*H, ngates, nq, qubits, diags);
  });
}

int b200sv_apply_multiplexer(b200sv_handle h, const uint64_t *ctrl, int nc, const uint64_t *tgt, int nt,
                             const double *mat) {
  return guard([&] {
    select(H);
    // qubits = targets ++ controls; block b acts on the targets (qubitvector.hpp:1305-1340).
    std::vector<uint64_t> all(tgt, tgt + nt);
    all.insert(all.end(), ctrl, ctrl + nc);
    const int k = nc + nt;
    auto q = checked_qubits(*H, all.data(), k);
    const uint64_t DIM = 1ull << k, columns = 1ull << nt, blocks = 1ull << nc;
    std::vector<double> full(2 * DIM * DIM, 0.0);  // block-diagonal expansion, column major
    for (uint64_t b = 0; b < blocks; b++)
      for (uint64_t i = 0; i < columns; i++)
        for (uint64_t j = 0; j < columns; j++) {
          const uint64_t src = i + b * columns + DIM * j;
          const uint64_t dst = (i + b * columns) + DIM * (j + b * columns);
          full[2 * dst] = mat[2 * src];
          full[2 * dst + 1] = mat[2 * src + 1];
        }
    if (k <= kMaxRegQubits) launch_dense(*H, q.data(), k, nullptr, 0, full.data());
    else launch_dense_generic(*H, q.data(), k, full.data());
  });
}

int b200sv_apply_permutation(b200sv_handle h, const uint64_t *qubits, int k, const uint64_t *pairs, int npairs) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    launch_permutation(*H, q.data(), k, pairs, npairs);
  });
}

int b200sv_apply_mcx(b200sv_handle h, const uint64_t *qubits, int k) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("apply_mcx: empty qubit list");
    auto q = checked_qubits(*H, qubits, k, true);
    const int t = q.back();
    if (t >= H->nq) throw Error("apply_mcx: target on a global qubit needs a chunk swap first");
    std::vector<int> c(q.begin(), q.end() - 1);
    if (!resolve_controls(*H, c)) return;
    launch_mcx(*H, c.data(), (int)c.size(), t);
  });
}
int b200sv_apply_mcy(b200sv_handle h, const uint64_t *qubits, int k) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("apply_mcy: empty qubit list");
    auto q = checked_qubits(*H, qubits, k, true);
    const int t = q.back();
    if (t >= H->nq) throw Error("apply_mcy: target on a global qubit needs a chunk swap first");
    std::vector<int> c(q.begin(), q.end() - 1);
    if (!resolve_controls(*H, c)) return;
    launch_mcy(*H, c.data(), (int)c.size(), t);
  });
}
int b200sv_apply_mcswap(b200sv_handle h, const uint64_t *qubits, int k) {
  return guard([&] {
    select(H);
    if (k < 2) throw Error("apply_mcswap: needs at least two qubits");
    auto q = checked_qubits(*H, qubits, k, true);
    const int t0 = q[k - 2], t1 = q[k - 1];
    if (t0 >= H->nq || t1 >= H->nq) throw Error("apply_mcswap: target on a global qubit needs a chunk swap first");
    std::vector<int> c(q.begin(), q.end() - 2);
    if (!resolve_controls(*H, c)) return;
    launch_mcswap(*H, c.data(), (int)c.size(), t0, t1);
  });
}
int b200sv_apply_mcphase(b200sv_handle h, const uint64_t *qubits, int k, double re, double im) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("apply_mcphase: empty qubit list");
    auto q = checked_qubits(*H, qubits, k, true);
    if (!resolve_controls(*H, q)) return;  // every listed qubit acts as a control of the phase
    if (q.empty()) {                       // all listed qubits are global and set: scalar phase on this chunk
      int q0 = 0;
      double dd[4] = {re, im, re, im};
      launch_diagonal(*H, &q0, 1, dd);
      return;
    }
    launch_mcphase(*H, q.data(), (int)q.size(), re, im);
  });
}

int b200sv_apply_mcu(b200sv_handle h, const uint64_t *qubits, int k, const double *m) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("apply_mcu: empty qubit list");
    auto q = checked_qubits(*H, qubits, k, true);
    // reference routing on exact equality (qubitvector.hpp:1626-1633)
    const bool offdiag_zero = m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0;
    if (offdiag_zero && m[0] == 1.0 && m[1] == 0.0) {
      std::vector<int> all = q;
      if (!resolve_controls(*H, all)) return;
      if (all.empty()) { int q0 = 0; double dd[4] = {m[6], m[7], m[6], m[7]}; launch_diagonal(*H, &q0, 1, dd); return; }
      launch_mcphase(*H, all.data(), (int)all.size(), m[6], m[7]);
      return;
    }
    const int t = q.back();
    std::vector<int> c(q.begin(), q.end() - 1);
    if (t >= H->nq) {
      if (!offdiag_zero) throw Error("apply_mcu: non-diagonal target on a global qubit needs a chunk swap first");
      // diagonal on a global target: scalar d[bit] under the (local) controls
      if (!resolve_controls(*H, c)) return;
      const int bit = (int)((H->chunk_index >> (t - H->nq)) & 1ull);
      const double re = m[6 * bit], im = m[6 * bit + 1];
      if (c.empty()) { int q0 = 0; double dd[4] = {re, im, re, im}; launch_diagonal(*H, &q0, 1, dd); }
      else launch_mcphase(*H, c.data(), (int)c.size(), re, im);
      return;
    }
    if (!resolve_controls(*H, c)) return;
    if (offdiag_zero && c.empty()) {
      double dd[4] = {m[0], m[1], m[6], m[7]};
      launch_diagonal(*H, &t, 1, dd);
      return;
    }
    launch_dense(*H, &t, 1, c.data(), (int)c.size(), m);
  });
}

int b200sv_apply_pauli(b200sv_handle h, const uint64_t *qubits, int k, const char *pauli, double cre, double cim) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    PauliMasks m = pauli_masks(q, pauli);
    if (m.x + m.z == 0) return;  // identity string: no-op even with coeff (qubitvector.hpp:2400-2403)
    add_y_phase(m.num_y, cre, cim);
    launch_pauli(*H, m.x, m.z, m.x_max, cre, cim);
  });
}

int b200sv_apply_gate_sequence(b200sv_handle h, int ngates, const int *nq, const uint64_t *qubits,
                               const double *mats, int *passes_out) {
  return guard([&] {
    select(H);
    if (ngates < 0 || (ngates > 0 && (!nq || !qubits || !mats))) throw Error("apply_gate_sequence: bad arguments");
    const int passes = ngates ? apply_gate_sequence(*H, ngates, nq, qubits, mats, 3) : 0;
    if (passes_out) *passes_out = passes;
  });
}

int b200sv_apply_op_sequence(b200sv_handle h, int nops, const int *kind, const uint64_t *qubits, const double *mats,
                             const int *slot, const uint8_t *codes, int nslots, int *passes_out) {
  return guard([&] {
    select(H);
    if (nops < 0 || (nops > 0 && (!kind || !qubits || !mats))) throw Error("apply_op_sequence: bad arguments");
    const int passes = nops ? apply_gate_sequence(*H, nops, kind, qubits, mats, 3, slot, codes, nslots) : 0;
    if (passes_out) *passes_out = passes;
  });
}

int b200sv_selftest_op_sequence(int num_qubits, int64_t num_states, int precision, void *host_state, int nops,
                                const int *kind, const uint64_t *qubits, const double *mats, const int *slot,
                                const uint8_t *codes, int nslots, int *passes_out) {
  return guard([&] {
    // host_state == NULL: plan only (the pass count for any register size, e.g. a 33-qubit slice)
    if (nops < 1 || !kind || !qubits || !mats || num_qubits < 12 || num_qubits > (host_state ? 24 : 40) || num_states < 1)
      throw Error("selftest_op_sequence: bad arguments");
    State st;  // no device, no stream: the tile passes are interpreted on the host array
    st.nq = num_qubits;
    st.nstates = num_states;
    st.precision = precision;
    st.selftest_host = host_state ? host_state : (void *)&st;
    st.plan_only = host_state == nullptr;
    const int passes = apply_gate_sequence(st, nops, kind, qubits, mats, 3, slot, codes, nslots);
    if (passes_out) *passes_out = passes;
  });
}

int b200sv_plan_epochs(int num_qubits, int local_qubits, int nops, const int *op_off, const int *op_qubits,
                       const uint8_t *need_local, int min_run_bits, int multi_swap, int *phys, int64_t *plan_out,
                       int64_t plan_cap, int64_t *plan_len) {
  return guard([&] {
    if (num_qubits < 1 || num_qubits > 62 || local_qubits < 1 || local_qubits > num_qubits || nops < 0 || !phys ||
        !plan_len || (nops > 0 && (!op_off || !op_qubits || !need_local)))
      throw Error("plan_epochs: bad arguments");
    std::vector<char> seen(num_qubits, 0);
    for (int q = 0; q < num_qubits; q++) {
      if (phys[q] < 0 || phys[q] >= num_qubits || seen[phys[q]]) throw Error("plan_epochs: phys is not a permutation");
      seen[phys[q]] = 1;
    }
    for (int k = 0; k < (nops ? op_off[nops] : 0); k++)
      if (op_qubits[k] < 0 || op_qubits[k] >= num_qubits) throw Error("plan_epochs: qubit out of range");
    std::vector<int64_t> out;
    plan_epochs(num_qubits, local_qubits, num_qubits - local_qubits, nops, op_off, op_qubits, need_local, min_run_bits,
                multi_swap != 0, phys, out);
    *plan_len = (int64_t)out.size();
    if ((int64_t)out.size() > plan_cap || (out.size() && !plan_out)) throw Error("plan_epochs: plan buffer too small");
    std::copy(out.begin(), out.end(), plan_out);
  });
}

int b200sv_fuse_assign(int nops, const int *op_off, const int *op_qubits, const uint8_t *op_is_diag, int max_qubit,
                       int window, int max_diag_qubit, int *block_of_op, int *nblocks) {
  return guard([&] {
    if (nops < 0 || !nblocks || (nops > 0 && (!op_off || !op_qubits || !op_is_diag || !block_of_op)) || max_qubit < 1 ||
        max_qubit > 10 || max_diag_qubit < 1 || max_diag_qubit > 62 || window < 1)
      throw Error("fuse_assign: bad arguments");
    for (int k = 0; k < (nops ? op_off[nops] : 0); k++)
      if (op_qubits[k] < 0 || op_qubits[k] > 62) throw Error("fuse_assign: qubit out of range");
    fuse_assign(nops, op_off, op_qubits, op_is_diag, max_qubit, window, max_diag_qubit, block_of_op, nblocks);
  });
}

int b200sv_fuse_block_matrix(int k, const int *block_qubits, int ngates, const int *gate_off, const int *gate_qubits,
                             const int64_t *gate_moff, const double *gate_mats, int diag, double *out) {
  return guard([&] {
    if (k < 1 || k > (diag ? 28 : 10) || ngates < 1 || !block_qubits || !gate_off || !gate_qubits || !gate_moff ||
        !gate_mats || !out)
      throw Error("fuse_block_matrix: bad arguments");
    fuse_block_matrix(k, block_qubits, ngates, gate_off, gate_qubits, gate_moff, gate_mats, diag, out);
  });
}

int b200sv_apply_batched_pauli(b200sv_handle h, const uint64_t *masks4) {
  return guard([&] {
    select(H);
    for_state_slabs(H, [&](State &v, int64_t s0) { launch_batched_pauli(v, masks4 + 4 * s0); });
  });
}

int b200sv_collapse(b200sv_handle h, const uint64_t *qubits, int k, const uint64_t *outcomes, const double *scales,
                    const uint8_t *active) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    if (!outcomes || !scales || !active) throw Error("collapse: null argument");
    for_state_slabs(H, [&](State &v, int64_t s0) { launch_collapse(v, q.data(), k, outcomes + s0, scales + s0, active + s0); });
  });
}

int b200sv_apply_batched_matrix(b200sv_handle h, const uint64_t *qubits, int k, const double *mats, int nmats,
                                const int *index, const double *scale) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    if (!mats || !index || !scale || nmats < 1) throw Error("apply_batched_matrix: bad arguments");
    for (int64_t st = 0; st < H->nstates; st++)
      if (index[st] >= nmats) throw Error("apply_batched_matrix: matrix index out of range");
    if (k >= 1 && k <= 3) {
      for_state_slabs(H, [&](State &v, int64_t s0) { launch_batched_matrix(v, q.data(), k, mats, nmats, index + s0, scale + s0); });
      return;
    }
    // wider blocks: one launch per state through a view (rare: Kraus channels act on 1-2 qubits)
    const size_t msz = (size_t)2 << (2 * k);
    std::vector<double> scaled(msz);
    for (int64_t st = 0; st < H->nstates; st++) {
      if (index[st] < 0) continue;
      State v = *H;
      v.nstates = 1;
      v.data = (char *)H->data + ((uint64_t)st << H->nq) * H->amp_bytes();
      v.owns_data = false;
      for (size_t i = 0; i < msz; i++) scaled[i] = mats[(size_t)index[st] * msz + i] * scale[st];
      if (k <= kMaxRegQubits) launch_dense(v, q.data(), k, nullptr, 0, scaled.data());
      else launch_dense_generic(v, q.data(), k, scaled.data());
      H->scratch = v.scratch; H->scratch_bytes = v.scratch_bytes; H->pinned = v.pinned; H->pinned_bytes = v.pinned_bytes;
    }
  });
}

int b200sv_create_view(b200sv_handle *out, b200sv_handle parent, int64_t first_state, int64_t num_states) {
  return guard([&] {
    State *P = (State *)parent;
    select(P);
    if (first_state < 0 || num_states < 1 || first_state + num_states > P->nstates) throw Error("create_view: state range out of bounds");
    State *v = new State();
    v->device = P->device; v->nq = P->nq; v->nstates = num_states; v->precision = P->precision;
    v->global_nq = P->nq; v->num_sms = P->num_sms;
    v->data = (char *)P->data + ((uint64_t)first_state << P->nq) * P->amp_bytes();
    v->stream = P->stream;  // same stream: ordered with the parent's work
    *out = (b200sv_handle)v;
  });
}

// ------------------------------------------------------------------ reductions
int b200sv_norm(b200sv_handle h, double *out) {
  return guard([&] { select(H); for_state_slabs(H, [&](State &v, int64_t s0) { reduce_norm(v, out + s0); }); });
}

int b200sv_norm_matrix(b200sv_handle h, const uint64_t *qubits, int k, const double *mat, double *out) {
  return guard([&] {
    select(H);
    if (k < 1) throw Error("norm(qubits, mat): empty qubit list");
    auto q = checked_qubits(*H, qubits, k);
    for_state_slabs(H, [&](State &v, int64_t s0) { reduce_norm_matrix(v, q.data(), k, mat, out + s0); });
  });
}

int b200sv_probabilities(b200sv_handle h, const uint64_t *qubits, int k, double *out) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    for_state_slabs(H, [&](State &v, int64_t s0) {
      if (k == 0) reduce_norm(v, out + s0);
      else reduce_probabilities(v, q.data(), k, out + ((uint64_t)s0 << k));
    });
  });
}

int b200sv_sample_measure(b200sv_handle h, const double *rnds, int64_t shots, uint64_t *out) {
  return guard([&] {
    select(H);
    for_state_slabs(H, [&](State &v, int64_t s0) { sample_measure(v, rnds + s0 * shots, shots, out + s0 * shots); });
  });
}

int b200sv_expval_pauli(b200sv_handle h, const uint64_t *qubits, int k, const char *pauli, double pre, double pim,
                        double *out) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    PauliMasks m = pauli_masks(q, pauli);
    if (m.x + m.z == 0) {  // qubitvector.hpp:2309-2311
      for_state_slabs(H, [&](State &v, int64_t s0) { reduce_norm(v, out + s0); });
      return;
    }
    add_y_phase(m.num_y, pre, pim);
    for_state_slabs(H, [&](State &v, int64_t s0) { reduce_expval_pauli(v, m.x, m.z, m.x_max, pre, pim, nullptr, 0, 0, out + s0); });
  });
}

int b200sv_expval_pauli_pair(b200sv_handle h, const uint64_t *qubits, int k, const char *pauli,
                             const void *pair_dev_ptr, uint64_t z_count, uint64_t z_count_pair, double pre,
                             double pim, double *out) {
  return guard([&] {
    select(H);
    auto q = checked_qubits(*H, qubits, k);
    PauliMasks m = pauli_masks(q, pauli);
    add_y_phase(m.num_y, pre, pim);
    ensure_peer(*H, pair_dev_ptr);
    reduce_expval_pauli(*H, m.x, m.z, m.x_max, pre, pim, pair_dev_ptr ? pair_dev_ptr : H->data, z_count, z_count_pair,
                        out);
  });
}

// ------------------------------------------------------------------ exchange
int b200sv_chunk_swap_peer(b200sv_handle h, int local_q, void *peer, int upper, int half) {
  return guard([&] {
    select(H);
    if (local_q < 0 || local_q >= H->nq) throw Error("chunk swap: local qubit out of range");
    if (H->nstates != 1) throw Error("chunk swap: not available on batched containers");
    if (H->nq < 2) throw Error("chunk swap: chunk too small");
    ensure_peer(*H, peer);
    launch_chunk_swap_peer(*H, local_q, peer, upper, half);
  });
}
int b200sv_multi_swap_peer(b200sv_handle h, int k, const int *local_q, uint32_t my_g, void *const *peers) {
  return guard([&] {
    select(H);
    if (H->nstates != 1) throw Error("multi swap: not available on batched containers");
    if (k < 1 || k > 4 || H->nq <= k) throw Error("multi swap: bad qubit count");
    for (int b = 0; b < k; b++) {
      if (local_q[b] < 0 || local_q[b] >= H->nq) throw Error("multi swap: local qubit out of range");
      for (int c = 0; c < b; c++)
        if (local_q[c] == local_q[b]) throw Error("multi swap: duplicate local qubit");
    }
    if (my_g >> k) throw Error("multi swap: my_g out of range");
    for (uint32_t v = 0; v < (1u << k); v++)
      if (v != my_g) ensure_peer(*H, peers[v]);
    launch_multi_swap_peer(*H, k, local_q, my_g, peers);
  });
}
int b200sv_swap_range_peer(b200sv_handle h, uint64_t dest_offset, void *peer, uint64_t src_offset, uint64_t count) {
  return guard([&] {
    select(H);
    if (!peer) throw Error("swap range: null peer pointer");
    if (dest_offset + count > H->total_amps()) throw Error("swap range: range exceeds the chunk");
    ensure_peer(*H, peer);
    launch_swap_range_peer(*H, dest_offset, peer, src_offset, count);
  });
}
int b200sv_copy_range_peer(b200sv_handle h, uint64_t dest_offset, const void *peer, uint64_t src_offset, uint64_t count) {
  return guard([&] {
    select(H);
    if (!peer) throw Error("copy range: null peer pointer");
    if (dest_offset + count > H->total_amps()) throw Error("copy range: range exceeds the chunk");
    ensure_peer(*H, peer);
    B200_CUDA(cudaMemcpyAsync((char *)H->data + dest_offset * H->amp_bytes(), (const char *)peer + src_offset * H->amp_bytes(),
                              count * H->amp_bytes(), cudaMemcpyDeviceToDevice, H->stream));
  });
}
int b200sv_pack_half(b200sv_handle h, int local_q, int bit, uint64_t begin, uint64_t count, void *buf) {
  return guard([&] {
    select(H);
    if (local_q < 0 || local_q >= H->nq) throw Error("pack_half: local qubit out of range");
    if (begin + count > (H->total_amps() >> 1)) throw Error("pack_half: range exceeds half the chunk");
    launch_pack_half(*H, local_q, bit, begin, count, buf, false);
  });
}
int b200sv_unpack_half(b200sv_handle h, int local_q, int bit, uint64_t begin, uint64_t count, const void *buf) {
  return guard([&] {
    select(H);
    if (local_q < 0 || local_q >= H->nq) throw Error("unpack_half: local qubit out of range");
    if (begin + count > (H->total_amps() >> 1)) throw Error("unpack_half: range exceeds half the chunk");
    launch_pack_half(*H, local_q, bit, begin, count, (void *)buf, true);
  });
}

int b200sv_ipc_export(b200sv_handle h, void *handle64) {
  return guard([&] {
    select(H);
    if (!H->owns_data) throw Error("ipc_export: only library-owned allocations can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    B200_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, H->data));
  });
}
int b200sv_ipc_open(b200sv_handle h, const void *handle64, void **peer) {
  return guard([&] {
    select(H);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, sizeof(hd));
    B200_CUDA(cudaIpcOpenMemHandle(peer, hd, cudaIpcMemLazyEnablePeerAccess));
  });
}
int b200sv_ipc_close(b200sv_handle h, void *peer) {
  return guard([&] { select(H); B200_CUDA(cudaIpcCloseMemHandle(peer)); });
}
int b200sv_set_stream(b200sv_handle h, void *cuda_stream) {
  return guard([&] {
    select(H);
    B200_CUDA(cudaStreamSynchronize(H->stream));
    if (H->owns_stream && H->stream) B200_CUDA(cudaStreamDestroy(H->stream));
    H->stream = (cudaStream_t)cuda_stream;
    H->owns_stream = false;
  });
}

// ------------------------------------------------------------------ RNG (host)
int b200sv_rng_uniform(uint64_t seed, int64_t n, double *out) {
  return guard([&] {
    std::mt19937_64 rng(seed);  // RngEngine::set_seed / rand(0,1) (framework/rng.hpp:45-70)
    for (int64_t i = 0; i < n; i++) out[i] = std::uniform_real_distribution<double>(0.0, 1.0)(rng);
  });
}

}  // extern "C"
