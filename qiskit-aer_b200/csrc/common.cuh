// b200sv internal header: handle layout, error plumbing, complex/index helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>
#include <cstring>

#include <nvtx3/nvToolsExt.h>  // header-only (the tools library is loaded on demand by an attached profiler)

#include "../../include/b200sv.h"

namespace b200sv {

constexpr int kMaxDenseQubits = 10;   // generic (smem) dense kernel limit, like the reference's K3 (chunk_container.hpp:850)
constexpr int kMaxRegQubits = 5;      // register-resident dense kernel (fusion_max_qubit default, fusion.hpp:762)
constexpr int kMaxDiagQubits = 10;
constexpr int kMaxInsert = 40;        // zero-insert positions (targets + controls)

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

void set_last_error(const std::string &msg);
// device memory with the one-block-per-device reuse cache of api.cu (OOM releases the cache and retries)
void *device_alloc(int device, size_t bytes, bool may_reuse);
void device_free(int device, void *p, size_t bytes, bool may_cache);
void trim_alloc_cache();

#define B200_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw ::b200sv::Error(std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

// NVTX range around a host-side phase (tile pass, exchange, reduction): shows up in Nsight Systems / ncu --nvtx
// timelines (SURVEY section 5 tracing); costs a function-pointer test when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

template <typename T> struct cx_of;
template <> struct cx_of<double> { using type = double2; };
template <> struct cx_of<float> { using type = float2; };
template <typename T> using cx = typename cx_of<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T re, T im) {
  cx<T> r; r.x = re; r.y = im; return r;
}
// a*b
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// acc += a*b   (4 FMAs)
template <typename C> __device__ __forceinline__ void cfma(C &acc, C a, C b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// insert a zero bit at position `pos` (index0 building block, indexes.hpp:212-222)
__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t v, int pos) {
  const uint64_t low = v & ((1ull << pos) - 1);
  return ((v >> pos) << (pos + 1)) | low;
}

struct InsertList {  // sorted ascending positions
  int n;
  uint8_t pos[kMaxInsert];
};
__host__ __device__ __forceinline__ uint64_t insert_zeros(uint64_t v, const InsertList &l) {
  for (int i = 0; i < l.n; i++) v = insert_zero(v, l.pos[i]);
  return v;
}

struct TilePlan;  // tile.cu: captured tile passes (parameter blocks + kernel variants)
// a sub-cube of the state: `nbits` global index positions held at the bits of `value`
struct SlabSpec {
  int nbits = 0;
  int pos[4] = {0, 0, 0, 0};
  uint32_t value = 0;
  int sm_limit = 0;   // > 0: use at most this many SMs (leave the rest to a concurrent kernel)
};

// ---------------------------------------------------------------------------
struct State {
  int device = 0;
  int nq = 0;                 // local qubits per state
  int64_t nstates = 1;
  int precision = B200SV_F64;
  void *data = nullptr;       // nstates << nq amplitudes
  bool owns_data = false;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  // sharding (b200sv_set_chunk)
  int global_nq = 0;
  uint64_t chunk_index = 0;
  // scratch
  void *scratch = nullptr;    // device scratch for reductions / sampler
  size_t scratch_bytes = 0;
  void *pinned = nullptr;     // pinned host staging
  size_t pinned_bytes = 0;
  void *checkpoint = nullptr;
  int num_sms = 148;
  // scheduler self-test only (b200sv_selftest_op_sequence): a HOST array the tile-pass parameter blocks are
  // interpreted on instead of being launched; never set on a handle
  void *selftest_host = nullptr;
  bool plan_only = false;  // selftest without a host array: count the passes only
  const uint8_t *selftest_codes = nullptr;
  TilePlan *capture = nullptr;  // set while tile_plan_build runs: passes are recorded instead of launched
  bool in_layer_split = false;  // apply_gate_sequence is running the dense part of a diagonal-layer split

  uint64_t amps_per_state() const { return 1ull << nq; }
  uint64_t total_amps() const { return (uint64_t)nstates << nq; }
  size_t amp_bytes() const { return precision == B200SV_F64 ? 16 : 8; }
  void *ensure_scratch(size_t bytes);
  void *ensure_pinned(size_t bytes);
};

// sorted copy + validation of a qubit list against nq (throws)
std::vector<int> checked_qubits(const State &s, const uint64_t *qubits, int k, bool allow_global = false);

// ---- kernel launchers (gates.cu) -------------------------------------------
void launch_dense(State &s, const int *targets, int k, const int *controls, int nc, const double *mat_colmajor);
void launch_dense_generic(State &s, const int *targets, int k, const double *mat_colmajor);
void launch_diagonal(State &s, const int *qubits, int k, const double *diag);
void launch_diag_layer(State &s, int ngates, const int *nq, const uint64_t *qubits, const double *diags);
void launch_mcphase(State &s, const int *qubits, int k, double re, double im);
void launch_mcx(State &s, const int *controls, int nc, int target);
void launch_mcy(State &s, const int *controls, int nc, int target);
void launch_mcswap(State &s, const int *controls, int nc, int t0, int t1);
void launch_permutation(State &s, const int *qubits, int k, const uint64_t *pairs, int npairs);
void launch_pauli(State &s, uint64_t x_mask, uint64_t z_mask, int x_max, double pre, double pim);
void launch_batched_pauli(State &s, const uint64_t *masks4_host);
void launch_collapse(State &s, const int *qubits, int k, const uint64_t *outcomes, const double *scales,
                     const uint8_t *active);
void launch_batched_matrix(State &s, const int *qubits, int k, const double *mats, int nmats, const int *index,
                           const double *scale);
void launch_gather_line(State &s, int row_bits, uint64_t xor_mask, void *host_out);
void launch_init(State &s, bool ket0);
void launch_init_component(State &s, const int *qubits, int k, const double *state);
void launch_pack_half(State &s, int q, int bit, uint64_t begin, uint64_t count, void *buf, bool unpack);
void launch_chunk_swap_peer(State &s, int q, void *peer, int upper, int half);
void launch_multi_swap_peer(State &s, int k, const int *local_q, uint32_t my_g, void *const *peers);
void launch_swap_range_peer(State &s, uint64_t dest_offset, void *peer, uint64_t src_offset, uint64_t count);

// ---- tile-blocked multi-gate passes (tile.cu)
int apply_gate_sequence(State &s, int ngates, const int *nq, const uint64_t *qubits, const double *mats, int low_bits,
                        const int *slot = nullptr, const uint8_t *codes_host = nullptr, int nslots = 0);

void measure_fp64_peak(int device, double duration_ms, double *burst_tflops, double *sustained_tflops);

// plan once, launch later (whole state or slab by slab): the sharded executor's view of the tile engine
TilePlan *tile_plan_build(State &s, int ngates, const int *nq, const uint64_t *qubits, const double *mats);
void tile_plan_free(TilePlan *plan);
int tile_plan_passes(const TilePlan *plan);
uint64_t tile_plan_pass_mask(const TilePlan *plan, int pass);
void tile_plan_launch(State &s, const TilePlan *plan, int pass, const SlabSpec *slab);

// ---- epoch planner for sharded registers (planner.cu, host only)
void plan_epochs(int n, int nl, int gbits, int nops, const int *op_off, const int *op_qubits, const uint8_t *need_local,
                 int min_run_bits, bool multi_swap, int *phys, std::vector<int64_t> &out);

void fuse_assign(int nops, const int *op_off, const int *op_qubits, const uint8_t *op_is_diag, int max_qubit, int window,
                 int max_diag_qubit, int *block_of_op, int *nblocks_out);
void fuse_block_matrix(int k, const int *block_qubits, int ngates, const int *gate_off, const int *gate_qubits,
                       const int64_t *gate_moff, const double *gate_mats, int diag, double *out);

// ---- reductions (reduce.cu) -------------------------------------------------
void reduce_norm(State &s, double *out);
void reduce_norm_matrix(State &s, const int *qubits, int k, const double *mat, double *out);
void reduce_probabilities(State &s, const int *qubits, int k, double *out);
void reduce_expval_pauli(State &s, uint64_t x_mask, uint64_t z_mask, int x_max, double pre, double pim,
                         const void *pair, uint64_t zc, uint64_t zcp, double *out);
void reduce_inner_product(State &s, const void *other, double *re, double *im);
void sample_measure(State &s, const double *rnds, int64_t shots, uint64_t *out);
void reduce_dm_expval(State &s, int m, uint64_t x_mask, uint64_t z_mask, double pre, double pim, double *out);
void reduce_dm_probabilities(State &s, int m, const int *qubits, int k, double *out);

}  // namespace b200sv
