/* b200sv -- B200-native statevector engine: the thin C ABI.
 *
 * This is the drop-in boundary behind Qiskit Aer's C++ `QubitVector` method
 * set (the duck-typed `statevec_t` of `Statevector::State`,
 * /root/reference/src/simulators/statevector/statevector_state.hpp:100-101).
 * Every entry point names the reference method it replaces (file:line refer
 * to /root/reference/src/simulators/statevector/ unless stated otherwise).
 * Style precedent: the reference's own C facade,
 * contrib/runtime/aer_runtime.cpp:19-240 (opaque handle, plain scalars).
 *
 * Conventions (identical to the reference, qubitvector.hpp:225-411):
 *   - qubit q <-> bit q of the amplitude index (little endian);
 *   - qubit lists are uint64_t arrays, controls first, target(s) last;
 *   - matrices are complex<double>, interleaved (re,im), COLUMN-major
 *     vectorised: mat[i + dim*j] = M[i][j]; qubits[0] = least significant
 *     matrix bit; diagonals have 2^k entries;
 *   - Pauli strings: pauli[k-1-i] acts on qubits[i];
 *   - random draws (rnds) are uniform in [0,1) and come from the HOST
 *     (Aer's RngEngine); only draws cross this ABI.
 * One handle = `num_states` statevectors of `num_qubits` qubits, stored
 * contiguously on ONE device (state s occupies amplitudes [s<<n, (s+1)<<n)).
 * num_states > 1 is the batched-shot container (qubitvector_thrust.hpp:
 * 1184-1214): gates act on every state, reductions return one value per state.
 * A sharded state uses one handle per GPU (= per process) with
 * b200sv_set_chunk() describing which slice of the global register it holds.
 *
 * All functions return 0 on success, non-zero on error (then
 * b200sv_last_error() describes it; the C++ adapter re-throws, matching the
 * reference's exception behaviour, circuit_executor.hpp:574,726).
 * There is NO CPU fallback: without a CUDA device every call fails loudly.
 * Thread-safety: a handle may be used by one host thread at a time; every call
 * selects the handle's device first (as the reference does,
 * chunk/device_chunk_container.hpp:131-135).
 */
#ifndef B200SV_H
#define B200SV_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200sv_state *b200sv_handle;

enum { B200SV_F64 = 64, B200SV_F32 = 32 };

/* ---- lifetime / configuration ------------------------------------------ */
/* ABI version (bumped on any signature change). */
int b200sv_version(void);
const char *b200sv_last_error(void);
/* number of visible CUDA devices (chunk_manager.hpp:166-190) */
int b200sv_device_count(int *count);

/* free / total memory of one device: what DeviceChunkContainer::Allocate sizes its chunk count from
 * (chunk/device_chunk_container.hpp:391-404 cudaMemGetInfo; circuit_executor.hpp:378-392 get_gpu_memory_mb) */
int b200sv_mem_info(int device, uint64_t *free_bytes, uint64_t *total_bytes);

/* Measurement aid (no reference counterpart): the FP64 FMA rate this device sustains with constant-bank multiplier
 * operands -- the denominator of the tile passes' FP64 roofline, measured in the same run as the numbers it bounds.
 * burst = best single launch, sustained = mean of the second half of ~duration_ms of back-to-back launches. */
int b200sv_measure_fp64_peak(int device, double duration_ms, double *burst_tflops, double *sustained_tflops);

/* Releases the device memory the library keeps for reuse: the slice of the last destroyed handle (one block per
 * device; Aer's executors re-create the register for every circuit, and a 128 GiB cudaMalloc/cudaFree pair costs
 * ~0.2 s).  Allocations that fail release it by themselves; call this before handing the GPU to another process. */
int b200sv_trim(void);

/* QubitVector(), set_num_qubits (qubitvector.hpp:922; thrust :860 chunk_setup).
 * Allocates num_states << num_qubits amplitudes on `device` and a private stream. */
int b200sv_create(b200sv_handle *out, int num_qubits, int64_t num_states, int precision, int device);
/* Same, but adopts device memory + stream owned by the caller (e.g. a torch
 * tensor / torch stream used for NCCL plumbing).  dev_ptr must hold
 * num_states << num_qubits amplitudes of the given precision, 16-byte aligned. */
int b200sv_create_external(b200sv_handle *out, int num_qubits, int64_t num_states, int precision, int device,
                           void *dev_ptr, void *cuda_stream);
int b200sv_destroy(b200sv_handle h);
int b200sv_num_qubits(b200sv_handle h, int *n);
/* raw device pointer / stream (for peer exchange and torch interop) */
int b200sv_device_ptr(b200sv_handle h, void **dev_ptr);
int b200sv_stream(b200sv_handle h, void **cuda_stream);
/* chunk_setup(chunk_bits, num_qubits, chunk_index, ...) (qubitvector.hpp:1045;
 * thrust :860): this handle is chunk `chunk_index` of a register of
 * `global_num_qubits` qubits; local qubits are [0, num_qubits). Controls /
 * diagonal qubits >= num_qubits are then resolved from chunk_index without
 * data movement (thrust_kernels.hpp:1190,1342,2004 base_index_). */
int b200sv_set_chunk(b200sv_handle h, int global_num_qubits, uint64_t chunk_index);
/* synchronize() (qubitvector_thrust.hpp:1088-1099) */
int b200sv_synchronize(b200sv_handle h);
/* set_sample_measure_index_size is accepted for API parity and ignored. */

/* ---- data -------------------------------------------------------------- */
int b200sv_initialize(b200sv_handle h);                 /* initialize(): |0..0> in every state (qubitvector.hpp:1088) */
int b200sv_zero(b200sv_handle h);                       /* zero() (qubitvector.hpp:903) */
/* initialize_from_data / copy_to_vector (qubitvector.hpp:1153; thrust CopyIn/CopyOut):
 * host <-> device copy of `count` amplitudes starting at amplitude `offset`
 * (over the whole batch), in the handle's precision. */
int b200sv_upload(b200sv_handle h, const void *host, uint64_t offset, uint64_t count);
int b200sv_download(b200sv_handle h, void *host, uint64_t offset, uint64_t count);
/* Density-matrix helper: the state is vec(rho) of a 2^m x 2^m matrix (index = row + col * 2^m,
 * densitymatrix.hpp:292-343).  Copies the 2^m entries rho[i ^ xor_mask, i] to the host: xor_mask = 0 is the
 * diagonal used by probability()/probabilities()/sample_measure (densitymatrix.hpp:590-593) and the other
 * masks are the lines DensityMatrix::expval_pauli walks (:470-520). */
int b200sv_download_line(b200sv_handle h, int row_bits, uint64_t xor_mask, void *host_out);
/* Density-matrix reductions on the device (state = vec(rho), as above): DensityMatrix::expval_pauli
 * (densitymatrix.hpp:455-520; GPU twin density_expval_pauli_func, densitymatrix_thrust.hpp:1011-1188) =
 * sum_i Re(phase * rho[i ^ x, i]) (-1)^popcount(i & z) with the Pauli given like b200sv_expval_pauli (qubits < row_bits;
 * the identity string returns the trace), and the marginal probabilities of k <= 12 qubits from the diagonal
 * (densitymatrix.hpp:590-593 -> qubitvector.hpp:2108; out has 2^k entries). */
int b200sv_dm_expval_pauli(b200sv_handle h, int row_bits, const uint64_t *qubits, int k, const char *pauli, double pre,
                           double pim, double *out);
int b200sv_dm_probabilities(b200sv_handle h, int row_bits, const uint64_t *qubits, int k, double *out);
/* initialize_component(qubits, state) (qubitvector.hpp:879-900) */
int b200sv_initialize_component(b200sv_handle h, const uint64_t *qubits, int k, const double *state);
/* checkpoint()/revert(keep)/inner_product() (qubitvector.hpp:995-1041) -- device-resident copy */
int b200sv_checkpoint(b200sv_handle h);
int b200sv_revert(b200sv_handle h, int keep);
int b200sv_inner_product(b200sv_handle h, double *re, double *im);

/* ---- gates ------------------------------------------------------------- */
/* apply_matrix (qubitvector.hpp:1299 -> transformer.hpp:90) */
int b200sv_apply_matrix(b200sv_handle h, const uint64_t *qubits, int k, const double *mat);
/* apply_diagonal_matrix (qubitvector.hpp:1343 -> transformer.hpp:235) */
int b200sv_apply_diagonal(b200sv_handle h, const uint64_t *qubits, int k, const double *diag);
/* A LAYER of commuting diagonal 1-/2-qubit gates (cp, cz, rz, p, rzz, ... on any qubits) in one streaming pass, with no
 * 2^k table: nq[g] in {1, 2}, qubits = 2 entries per gate, diags = 4 complex<double> per gate (index = bit(q0) +
 * 2 bit(q1); 1-qubit gates use the first two).  Same result as one apply_diagonal_matrix per gate
 * (qubitvector.hpp:1343; DiagonalMult*, thrust_kernels.hpp:1318-1444); entries must be non-zero (projectors go through
 * apply_diagonal_matrix). */
int b200sv_apply_diagonal_layer(b200sv_handle h, int ngates, const int *nq, const uint64_t *qubits, const double *diags);
/* apply_multiplexer (qubitvector.hpp:1305) */
int b200sv_apply_multiplexer(b200sv_handle h, const uint64_t *ctrl, int nc, const uint64_t *tgt, int nt,
                             const double *mat);
/* apply_permutation_matrix (qubitvector.hpp:1354) pairs = 2*npairs matrix indices */
int b200sv_apply_permutation(b200sv_handle h, const uint64_t *qubits, int k, const uint64_t *pairs, int npairs);
/* apply_mcx / apply_mcy / apply_mcswap (qubitvector.hpp:1447,1489,1540) -- bit exact */
int b200sv_apply_mcx(b200sv_handle h, const uint64_t *qubits, int k);
int b200sv_apply_mcy(b200sv_handle h, const uint64_t *qubits, int k);
int b200sv_apply_mcswap(b200sv_handle h, const uint64_t *qubits, int k);
/* apply_mcphase (qubitvector.hpp:1576) */
int b200sv_apply_mcphase(b200sv_handle h, const uint64_t *qubits, int k, double re, double im);
/* apply_mcu (qubitvector.hpp:1615): 2x2 column-major on the last qubit, incl. the
 * reference's exact-== routing to phase / diagonal */
int b200sv_apply_mcu(b200sv_handle h, const uint64_t *qubits, int k, const double *mat);
/* apply_pauli (qubitvector.hpp:2393) */
int b200sv_apply_pauli(b200sv_handle h, const uint64_t *qubits, int k, const char *pauli, double cre, double cim);
/* apply_batched_pauli_ops (qubitvector_thrust.hpp:2892, batched_pauli_func :2819):
 * per-state Pauli given as num_states x {x_mask, z_mask, num_y, apply(0/1)} */
int b200sv_apply_batched_pauli(b200sv_handle h, const uint64_t *masks4);

/* Gate-queue flush: apply a SEQUENCE of 1- and 2-qubit dense gates (same semantics as calling
 * apply_matrix once per gate, in order) in as few HBM passes as possible: gates whose qubits fit a
 * 12-bit shared-memory tile ride on one pass.  nq[i] in {1,2}; qubits = 2 entries per gate
 * (second ignored for nq=1); mats = 16 complex<double> slots per gate, column-major like apply_matrix.
 * Reference counterpart: the blocked-gate queue, chunk/device_chunk_container.hpp:999-1108
 * (queue_blocked_gate) + :1208 (dev_apply_shared_memory_blocked_gates), entered through
 * QubitVectorThrust::apply_matrix / apply_mcx while register blocking is active
 * (qubitvector_thrust.hpp:1511-1512,1628-1634).  passes_out (optional) receives the pass count. */
int b200sv_apply_gate_sequence(b200sv_handle h, int ngates, const int *nq, const uint64_t *qubits,
                               const double *mats, int *passes_out);

/* Batched noisy shots: the same queue flush, with sampled Pauli noise riding on the passes.  kind[i] = 1 / 2
 * (dense 1-/2-qubit gate, as above) or 3 = a per-state Pauli on qubits[2*i]: state s applies
 * codes[slot[i]*num_states + s] (0..3 = I, X, Y, Z).  One launch covers every shot of the container; the Pauli
 * costs no extra HBM traffic.  Replaces apply_batched_pauli_ops + the per-op launches of
 * BatchShotsExecutor::apply_ops_batched_shots_for_group (src/simulators/batch_shots_executor.hpp:500-603;
 * qubitvector_thrust.hpp:2892, batched_pauli_func :2819).  The draws stay on the host (Aer's RngEngine). */
int b200sv_apply_op_sequence(b200sv_handle h, int nops, const int *kind, const uint64_t *qubits, const double *mats,
                             const int *slot, const uint8_t *codes, int nslots, int *passes_out);

/* Self-test of the pass / round scheduler behind the two calls above -- TEST INFRASTRUCTURE, not a compute path: no
 * handle and no device are involved.  The op list is partitioned exactly as b200sv_apply_op_sequence does, but each
 * tile-pass parameter block is INTERPRETED on `host_state` (num_states << num_qubits complex<double> or
 * complex<float> amplitudes, 12 <= num_qubits <= 24) with the kernels' addressing (staging maps, swizzled slots,
 * per-thread round blocks, gate forms), and the warp-local-segment invariant is checked.  host_state == NULL plans
 * only (returns the pass count; num_qubits up to 40).  Lets the CPU-only
 * test-suite cover host logic that otherwise runs only in front of a GPU.  No reference counterpart. */
int b200sv_selftest_op_sequence(int num_qubits, int64_t num_states, int precision, void *host_state, int nops,
                                const int *kind, const uint64_t *qubits, const double *mats, const int *slot,
                                const uint8_t *codes, int nslots, int *passes_out);

/* Epoch planner for a register sharded over 2^(num_qubits - local_qubits) devices by its top qubits (host code, no
 * device).  Ops are given by their logical qubits (op i = op_qubits[op_off[i] .. op_off[i+1]), need_local[] flags the
 * qubits that must be chunk-local for the op: targets of non-diagonal ops; controls and diagonal ops are resolved from
 * the chunk index).  phys[] (logical -> physical position, >= local_qubits means "selects the device") is updated.
 * plan_out receives int64 records: {0, op, nq, physical qubits...} | {1, local position, global bit} (pairwise exchange)
 * | {2, k, local positions..., global bits...} (one all-to-all pass, multi_swap != 0).  Every op runs as soon as its
 * qubits are local under the current map; at each stall the wanted global qubits replace the local ones whose next use is
 * farthest away; evicted positions >= min_run_bits are preferred (long contiguous runs for the exchange).
 * Role of CacheBlocking::optimize_circuit + swap insertion (src/transpile/cacheblocking.hpp:52, block_circuit :82, insert_swap :182) and of
 * ParallelStateExecutor::apply_chunk_swap scheduling (src/simulators/parallel_state_executor.hpp:1134). */
int b200sv_plan_epochs(int num_qubits, int local_qubits, int nops, const int *op_off, const int *op_qubits,
                       const uint8_t *need_local, int min_run_bits, int multi_swap, int *phys, int64_t *plan_out,
                       int64_t plan_cap, int64_t *plan_len);

/* The engine's gate fusion (host code, no device; role of Fusion::optimize_circuit, src/transpile/fusion.hpp:849, with
 * a B200 cost model instead of the CPU one of :1002-1136): dense blocks up to max_qubit qubits, purely diagonal blocks
 * up to max_diag_qubit, diagonal gates commute.  b200sv_fuse_assign maps every op (given by its qubits and a "is
 * diagonal" flag) to a block id (blocks are numbered in execution order); b200sv_fuse_block_matrix multiplies the gates
 * of one block (row-major 2^m x 2^m complex<double> each, bit j of a gate index <-> its j-th qubit) into the block's
 * 2^k x 2^k row-major matrix, or its 2^k diagonal when diag != 0 (bit i of the index <-> block_qubits[i]). */
int b200sv_fuse_assign(int nops, const int *op_off, const int *op_qubits, const uint8_t *op_is_diag, int max_qubit,
                       int window, int max_diag_qubit, int *block_of_op, int *nblocks);
int b200sv_fuse_block_matrix(int k, const int *block_qubits, int ngates, const int *gate_off, const int *gate_qubits,
                             const int64_t *gate_moff, const double *gate_mats, int diag, double *out);

/* Per-state measurement collapse for batched containers (apply_batched_measure / apply_batched_reset,
 * qubitvector_thrust.hpp:2251-2460: check_measure_probability_func + reset_after_measure_func): for every
 * state s with active[s] != 0, amplitudes whose `qubits` bits differ from outcomes[s] are zeroed and the
 * rest are multiplied by scales[s] (= 1/sqrt(p_outcome); scales[s] <= 0 means "normalise the survivor by its
 * own modulus", for measurements of every qubit).  One launch for all states. */
int b200sv_collapse(b200sv_handle h, const uint64_t *qubits, int k, const uint64_t *outcomes, const double *scales,
                    const uint8_t *active);
/* Per-state matrices in ONE launch: state s applies the column-major 2^k x 2^k matrix mats[index[s]] times scale[s]
 * (index[s] < 0: state s is left alone).  Batched Kraus channels (apply_batched_kraus + MatrixMultNxN_conditional,
 * qubitvector_thrust.hpp:2996-3177: index = the operator each shot's draw selected, scale = 1/sqrt(p)) and
 * per-parameter matrices of bound circuits (apply_batched_matrix, :1578-1611). */
int b200sv_apply_batched_matrix(b200sv_handle h, const uint64_t *qubits, int k, const double *mats, int nmats,
                                const int *index, const double *scale);
/* A handle onto states [first_state, first_state + num_states) of a batched container, sharing its memory
 * and stream (per-shot fallbacks of the batched executor, batch_shots_executor.hpp:594-603; per-parameter
 * matrices, apply_batched_matrix qubitvector_thrust.hpp:1578-1611).  Destroy it before the parent. */
int b200sv_create_view(b200sv_handle *out, b200sv_handle parent, int64_t first_state, int64_t num_states);

/* ---- reductions (out has num_states entries unless noted) --------------- */
int b200sv_norm(b200sv_handle h, double *out);          /* norm() (qubitvector.hpp:1879) */
/* norm(qubits, mat) (qubitvector.hpp:1889) -- Kraus probability ||M psi||^2 */
int b200sv_norm_matrix(b200sv_handle h, const uint64_t *qubits, int k, const double *mat, double *out);
/* probabilities(qubits) (qubitvector.hpp:2108): out[num_states][2^k], ALL outcomes in one pass */
int b200sv_probabilities(b200sv_handle h, const uint64_t *qubits, int k, double *out);
/* sample_measure(rnds) (qubitvector.hpp:2149): out[shots] (state 0 only when num_states == 1;
 * for batches rnds/out are [num_states][shots]).  Non-destructive. */
int b200sv_sample_measure(b200sv_handle h, const double *rnds, int64_t shots, uint64_t *out);
/* expval_pauli (qubitvector.hpp:2300) */
int b200sv_expval_pauli(b200sv_handle h, const uint64_t *qubits, int k, const char *pauli, double pre, double pim,
                        double *out);
/* expval_pauli with a pair chunk (qubitvector.hpp:2350; thrust_kernels.hpp:2413):
 * the X-part crosses a global qubit, pair_dev_ptr is the partner chunk's device memory. */
int b200sv_expval_pauli_pair(b200sv_handle h, const uint64_t *qubits, int k, const char *pauli,
                             const void *pair_dev_ptr, uint64_t z_count, uint64_t z_count_pair, double pre,
                             double pim, double *out);

/* ---- global-qubit exchange (sharded states) ------------------------------ */
/* apply_chunk_swap(qubits, chunk) (qubitvector.hpp:1753; thrust :1677, CSwapChunk_func
 * thrust_kernels.hpp:1884): swap local qubit `local_q` with the global qubit that
 * distinguishes this chunk from `peer_dev_ptr` (peer-mapped device memory of the partner
 * chunk).  `this_is_upper` = this chunk has the global bit set.  Each side calls it once and
 * moves half of the pairs (`half` = 0/1) so both NVLink directions are used. */
int b200sv_chunk_swap_peer(b200sv_handle h, int local_q, void *peer_dev_ptr, int this_is_upper, int half);
/* Several global qubits at once (apply_multi_chunk_swap, parallel_state_executor.hpp:1339-1552, the
 * "all-to-all shuffle of 2^nswap sub-blocks"): local qubits local_q[0..k) trade places with the k global bits
 * that select among 2^k ranks.  An amplitude whose local bits read l and whose rank's global bits read g moves
 * to the rank whose global bits read l, local bits g.  `my_g` is this rank's value of the k global bits;
 * peer_dev_ptrs[v] is the peer-mapped slice of the rank whose global bits read v (entry my_g is ignored).
 * Every rank calls it once; each unordered pair of sub-blocks is moved by exactly one of its two owners
 * (balanced), in place, both NVLink directions busy: (1 - 2^-k) of the slice crosses the links once instead
 * of k half-slice exchanges. */
int b200sv_multi_swap_peer(b200sv_handle h, int k, const int *local_q, uint32_t my_g, void *const *peer_dev_ptrs);
/* apply_chunk_swap(chunk, dest_offset, src_offset, size) (qubitvector.hpp:1824-1840; the sub-block shuffle of
 * apply_multi_chunk_swap, parallel_state_executor.hpp:1474-1500) and the whole-chunk exchange of a swap between two
 * global qubits (qubitvector.hpp:1765-1776): `count` amplitudes at this[dest_offset] trade places with
 * peer_dev_ptr[src_offset].  The peer chunk may live on another GPU of this process (peer access is enabled on first
 * use, chunk_manager.hpp:129-135) or be an IPC mapping. */
int b200sv_swap_range_peer(b200sv_handle h, uint64_t dest_offset, void *peer_dev_ptr, uint64_t src_offset, uint64_t count);
/* one-way variant (write_back = false, qubitvector.hpp:1771-1775; initialize(const QubitVector&) :1121): this[dest_offset ..)
 * <- peer_dev_ptr[src_offset ..), stream-ordered device-to-device copy */
int b200sv_copy_range_peer(b200sv_handle h, uint64_t dest_offset, const void *peer_dev_ptr, uint64_t src_offset, uint64_t count);
/* staging variant for NCCL send/recv: gather/scatter the half-slice of amplitudes whose
 * local qubit `local_q` equals `bit`, slice [begin, begin+count) of that half, to/from a
 * contiguous device buffer (send_buffer/recv_buffer, qubitvector.hpp:1061-1081). */
int b200sv_pack_half(b200sv_handle h, int local_q, int bit, uint64_t begin, uint64_t count, void *dev_buf);
int b200sv_unpack_half(b200sv_handle h, int local_q, int bit, uint64_t begin, uint64_t count, const void *dev_buf);

/* CUDA-IPC plumbing for one-process-per-GPU sharding: export this handle's allocation (64-byte
 * cudaIpcMemHandle_t; only for memory the library allocated itself), map a partner process's
 * allocation into this process (peer access over NVLink is enabled lazily) and unmap it.  The mapped
 * pointer is what b200sv_chunk_swap_peer takes.  Replaces the in-process peer pointers of
 * chunk/device_chunk_container.hpp:336-351 (cudaDeviceEnablePeerAccess) for the multi-process layout. */
int b200sv_ipc_export(b200sv_handle h, void *handle64);
int b200sv_ipc_open(b200sv_handle h, const void *handle64, void **peer_dev_ptr);
int b200sv_ipc_close(b200sv_handle h, void *peer_dev_ptr);
/* run this handle's kernels on a caller-owned stream from now on (e.g. the torch stream NCCL is ordered on) */
int b200sv_set_stream(b200sv_handle h, void *cuda_stream);

/* ---- sharded registers: the executor (csrc/sharded.cu) ------------------------------------------------------------
 * One register of num_qubits qubits split over `world` = 2^g shards by its top g qubits, one shard per GPU -- the
 * B200 counterpart of ParallelStateExecutor + ChunkManager (src/simulators/parallel_state_executor.hpp:318-376 chunk
 * placement, :772 apply_ops_chunks, :1134-1552 chunk swaps; statevector/chunk/chunk_manager.hpp:129-135,166-396
 * devices, peer access, memory sizing) and of Statevector::Executor's cross-chunk reductions
 * (statevector/statevector_executor.hpp:551 expval_pauli, :1149 sample_measure), with the host side in C++:
 * epoch planning, gate queues, tile passes, the exchange and its overlap with the passes next to it all run behind
 * these calls.  The shards listed in local_ranks live in THIS process on devices[i] (one process may drive all GPUs --
 * the layout of Aer's Controller -- or one GPU each under torchrun); shards of other processes are attached from the
 * 128-byte blobs their owners export (CUDA IPC: slice + staging / flag area).  staging_bytes = size of each shard's
 * staging area for the pipelined exchange ((uint64_t)-1: what the device can spare; 0: in-place exchange only). */
typedef struct b200sv_sharded *b200sv_sharded_handle;
enum { B200SV_OP_MATRIX = 0, B200SV_OP_DIAGONAL = 1, B200SV_OP_MCX = 2, B200SV_OP_MCY = 3, B200SV_OP_MCPHASE = 4,
       B200SV_OP_MCSWAP = 5, B200SV_OP_MCU = 6 };
int b200sv_sharded_create(b200sv_sharded_handle *out, int num_qubits, int precision, int world, int nlocal,
                          const int *local_ranks, const int *devices, uint64_t staging_bytes);
int b200sv_sharded_destroy(b200sv_sharded_handle h);
int b200sv_sharded_ipc_export(b200sv_sharded_handle h, int rank, void *blob128);
int b200sv_sharded_ipc_attach(b200sv_sharded_handle h, int rank, const void *blob128);
/* the plain handle of a local shard (download, norm, ... through the calls above; owned by the sharded handle) */
int b200sv_sharded_shard_handle(b200sv_sharded_handle h, int rank, b200sv_handle *out);
/* |0...0> of the whole register (Executor::initialize_qreg, statevector_executor.hpp:437-470); resets the qubit map */
int b200sv_sharded_initialize(b200sv_sharded_handle h);
/* waits for all local shards; reports a partner that never arrived at an exchange */
int b200sv_sharded_synchronize(b200sv_sharded_handle h);
/* Apply a circuit given by LOGICAL qubits: op i has kind kinds[i], qubits op_qubits[op_off[i] .. op_off[i+1]) (controls
 * first, targets last) and data[data_off[i] .. data_off[i+1]) doubles (matrix: 2*4^k column-major like
 * b200sv_apply_matrix; diagonal: 2*2^k; mcphase: re, im; mcu: 8; mcx / mcy / mcswap: none).  Asynchronous. */
int b200sv_sharded_apply_ops(b200sv_sharded_handle h, int nops, const int *kinds, const int *op_off, const int *op_qubits,
                             const int64_t *data_off, const double *data);
/* host-only: what apply_ops would do on a register of `world` shards with the given staging size -- out8 = {tile
 * passes of shard 0, exchanges, staged, in place, passes taken along by exchanges, finest slab count, coarsest slab
 * count, qubit swaps} (test and sizing aid; no device involved) */
int b200sv_sharded_plan_only(int num_qubits, int precision, int world, uint64_t staging_bytes, int nops, const int *kinds,
                             const int *op_off, const int *op_qubits, const int64_t *data_off, const double *data,
                             double *out8);
/* Executor self-test -- TEST INFRASTRUCTURE, no device: the register lives in host_state (2^n amplitudes of the given
 * precision, shard r = the r-th contiguous slice); the program is compiled and scheduled as apply_ops does and then
 * interpreted on the host in issue order (tile-pass parameter blocks whole and slab-patched, pushes / unstages as the
 * same strided copies, in-place exchanges as swaps), and the qubit order is restored.  out8 as b200sv_sharded_stats.
 * Lets the CPU-only test-suite cover the slab / staging / exchange-group index algebra and the pipeline's issue order. */
int b200sv_sharded_selftest(int num_qubits, int precision, int world, uint64_t staging_bytes, int nops, const int *kinds,
                            const int *op_off, const int *op_qubits, const int64_t *data_off, const double *data,
                            void *host_state, double *out8);
/* of the last apply_ops: {tile passes, exchanges, staged, in place, kernel launches, DMA copies, bytes sent per shard,
 * passes that ran slab-wise next to an exchange} */
int b200sv_sharded_stats(b200sv_sharded_handle h, double *out8);
/* CUDA-event profiling of the first local shard: profile(1) starts collecting, profile(0) stops; profile_read sums
 * {count, ms} pairs for: whole-state tile passes, staged exchange regions, pushes (copy stream), unstage kernels,
 * slab passes, in-place exchanges */
int b200sv_sharded_profile(b200sv_sharded_handle h, int on);
int b200sv_sharded_profile_read(b200sv_sharded_handle h, double *out12);
/* device time of the last apply_ops (CUDA events on the shards' compute streams, max over local shards) */
int b200sv_sharded_elapsed_ms(b200sv_sharded_handle h, double *ms);
/* logical qubit -> physical position (>= local qubits: selects the shard) */
int b200sv_sharded_qubit_map(b200sv_sharded_handle h, int *phys);
int b200sv_sharded_restore_order(b200sv_sharded_handle h);
int b200sv_sharded_norms(b200sv_sharded_handle h, double *out_world);
int b200sv_sharded_expval_pauli(b200sv_sharded_handle h, const uint64_t *qubits, int k, const char *pauli, double *partial);
int b200sv_sharded_sample_measure(b200sv_sharded_handle h, const double *rnds, int64_t shots, const double *norms_world,
                                  uint64_t *out);

/* ---- host RNG identical to Aer's RngEngine (framework/rng.hpp:31-99) ------ */
/* n draws of rand(0,1) from std::mt19937_64 seeded with `seed` (what
 * State::sample_measure consumes, statevector_state.hpp:1026-1027). */
int b200sv_rng_uniform(uint64_t seed, int64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif /* B200SV_H */
