// QubitVectorB200<data_t>: the reference-side adapter a maintainer adds to Aer.
//
// A header-only `statevec_t` for `AER::Statevector::State<statevec_t>`
// (src/simulators/statevector/statevector_state.hpp:100-101) exposing the same
// method set as AER::QV::QubitVector<data_t> (src/simulators/statevector/
// qubitvector.hpp:62-480 -- signatures mirrored one for one) plus the batched
// multi-shot hooks of QubitVectorThrust (src/simulators/statevector/
// qubitvector_thrust.hpp:1184-1214,2251-2484,2683-2713,2892-3402), forwarding
// every amplitude operation to the C ABI of the B200 engine (include/b200sv.h).
// Errors come back as the std::runtime_error the executors already catch
// (src/simulators/circuit_executor.hpp:574,726).
//
// Needs the Aer source tree on the include path (it uses Aer's own types:
// reg_t, cvector_t, AER::Vector, Operations::Op, RngEngine, ClassicalRegister,
// QV::Rotation).
//
// Three modes, chosen by how the executors call chunk_setup():
//  * single state (Executor::run_circuit_*): one vector = one handle;
//  * cache blocking (ParallelStateExecutor): one vector = one chunk = one handle;
//    the chunks of a register are spread over the target GPUs in contiguous runs
//    (chunk i of C lives on target_gpus[i * G / C], as chunk_manager.hpp:330-353
//    places them), like the CPU class it reports support_global_indexing() ==
//    false, so the State rewrites global-qubit diagonals/controls per chunk on
//    the host (statevector_state.hpp:735-753) and chunk swaps arrive through
//    apply_chunk_swap(qubits, other_chunk, write_back) /
//    apply_chunk_swap(other_chunk, dest_offset, src_offset, size): in-place
//    kernels over NVLink peer access when the two chunks sit on different GPUs;
//  * batched shots (BatchShotsExecutor, `batched_shots_gpu`): every vector of a
//    group is a view on ONE container handle with num_states = shots.  With
//    (one container per target GPU, shots split evenly: the executor sees one
//    group per GPU and drives them from parallel OpenMP threads).  With
//    enable_batch(true) the first vector executes for all states in one launch
//    and the others return immediately (the get_chunk_count() contract,
//    qubitvector_thrust.hpp:1201-1214); with enable_batch(false) a vector acts
//    on its own state only.  Classical registers live on the host, one
//    AER::ClassicalRegister per shot.
//
// Gate queue: dense 1-/2-qubit gates (and, in batched mode, cx / small
// diagonals / sampled Pauli noise) are queued and flushed through
// b200sv_apply_op_sequence, which packs them into as few HBM passes as possible;
// every other method flushes first, so observable semantics are those of
// immediate application.  Disable with B200SV_GATE_QUEUE=0.
#ifndef _qv_qubit_vector_b200_hpp_
#define _qv_qubit_vector_b200_hpp_

#include <complex>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200sv.h"
#include "framework/creg.hpp"
#include "framework/json.hpp"
#include "framework/linalg/vector.hpp"
#include "framework/operations.hpp"
#include "framework/rng.hpp"
#include "framework/types.hpp"
#include "framework/utils.hpp"
#include "simulators/statevector/qubitvector.hpp"  // QV::Rotation, pauli_masks_and_phase, Linalg::VMatrix

namespace AER {
namespace QV {

namespace b200detail {
inline void ck(int rc) {
  if (rc) throw std::runtime_error(std::string("b200sv: ") + b200sv_last_error());
}
// ops waiting for the next flush: dense 1-/2-qubit gates and per-state Pauli ops
struct OpQueue {
  std::vector<int> kind, slot;
  std::vector<uint64_t> qubits;
  std::vector<double> mats;
  std::vector<uint8_t> codes;  // [nslots][nstates]
  int nslots = 0;
  bool empty() const { return kind.empty(); }
  void clear() { kind.clear(); slot.clear(); qubits.clear(); mats.clear(); codes.clear(); nslots = 0; }
  void push_dense(const uint64_t *q, int k, const std::complex<double> *m) {
    kind.push_back(k);
    slot.push_back(0);
    qubits.push_back(q[0]);
    qubits.push_back(k == 2 ? q[1] : 0);
    const size_t off = mats.size();
    mats.resize(off + 32, 0.0);
    std::memcpy(&mats[off], m, ((size_t)1 << (2 * k)) * sizeof(std::complex<double>));
  }
  void flush(b200sv_handle h) {
    if (kind.empty()) return;
    const int rc = b200sv_apply_op_sequence(h, (int)kind.size(), kind.data(), qubits.data(), mats.data(), slot.data(),
                                            codes.empty() ? nullptr : codes.data(), nslots, nullptr);
    clear();
    ck(rc);
  }
};
// one batched-shot container shared by the vectors of a group
struct Batch {
  b200sv_handle h = nullptr;
  size_t nq = 0, nstates = 0;
  uint_t first_index = 0;
  std::vector<ClassicalRegister> cregs;
  int_t cond_reg = -1;
  bool on = true;
  OpQueue queue;
  ~Batch() { if (h) b200sv_destroy(h); }
};
// the containers of one multi-shot allocation, one per target GPU (chunk_manager.hpp:223-264 spreads the shots
// over the devices); part p holds the states [first[p], first[p + 1]) of the allocation
struct BatchSet {
  std::vector<std::shared_ptr<Batch>> parts;
  std::vector<uint_t> first;
};
// the chunks of one cache-blocked register: which GPU holds which chunk
struct ChunkGroup {
  std::vector<int> devices;
  uint_t first_index = 0, num_chunks = 1;
  int device_of(uint_t chunk_index) const {
    const uint_t local = chunk_index - first_index;
    return devices[(size_t)(local * devices.size() / num_chunks)];
  }
};
// target GPUs of an allocation: the executor's list (circuit_executor.hpp:340-359), else every visible device.
// B200SV_VIRTUAL_GPUS=k (test knob) repeats the list k times so that a one-GPU box exercises the multi-device
// placement, grouping and peer-swap paths.
inline std::vector<int> resolve_devices(const reg_t &targets, int fallback) {
  std::vector<int> d;
  for (auto t : targets) d.push_back((int)t);
  if (d.empty()) d.push_back(fallback);
  if (const char *e = getenv("B200SV_VIRTUAL_GPUS")) {
    const int k = atoi(e);
    std::vector<int> rep;
    for (int i = 0; i < k; i++) rep.insert(rep.end(), d.begin(), d.end());
    if (!rep.empty()) d.swap(rep);
  }
  return d;
}
}  // namespace b200detail

template <typename data_t = double> class QubitVectorB200 {
public:
  QubitVectorB200() = default;
  explicit QubitVectorB200(size_t num_qubits) { set_num_qubits(num_qubits); }
  virtual ~QubitVectorB200() { release(); }
  QubitVectorB200(const QubitVectorB200 &) {}  // like the reference: copies start empty (qubitvector.hpp:75)
  QubitVectorB200 &operator=(const QubitVectorB200 &) = delete;
  QubitVectorB200 &operator=(QubitVectorB200 &&o) {
    if (this != &o) {
      if (o.h_) o.flush();
      release();
      h_ = o.h_; num_qubits_ = o.num_qubits_; data_size_ = o.data_size_; chunk_index_ = o.chunk_index_;
      batch_ = std::move(o.batch_); batch_set_ = std::move(o.batch_set_); group_ = std::move(o.group_);
      batch_pos_ = o.batch_pos_; view_ = o.view_; tmp_view_ = o.tmp_view_; tmp_view_state_ = o.tmp_view_state_;
      device_ = o.device_; target_gpus_ = o.target_gpus_; queue_ = std::move(o.queue_);
      o.h_ = nullptr; o.view_ = nullptr; o.tmp_view_ = nullptr; o.num_qubits_ = 0; o.data_size_ = 0;
    }
    return *this;
  }

  static std::string name() { return "statevector_b200"; }
  static int &device() { static int d = 0; return d; }

  //---------------------------------------------------------------- size / config
  virtual void set_num_qubits(size_t num_qubits) {
    if (batch_) {
      if (num_qubits != batch_->nq) throw std::runtime_error("QubitVectorB200: batched container holds a different qubit count");
      num_qubits_ = num_qubits;
      data_size_ = 1ull << num_qubits;
      return;
    }
    if (h_ && num_qubits == num_qubits_) return;
    release();
    ck(b200sv_create(&h_, (int)num_qubits, 1, sizeof(data_t) == 8 ? B200SV_F64 : B200SV_F32, my_device()));
    num_qubits_ = num_qubits;
    data_size_ = 1ull << num_qubits;
  }
  // GPU this vector's amplitudes live on: its chunk's place in a cache-blocked register, else the first target GPU
  int my_device() const { return group_ ? group_->device_of(chunk_index_) : (device_ >= 0 ? device_ : device()); }
  virtual uint_t num_qubits() const { return num_qubits_; }
  uint_t size() const { return data_size_; }
  size_t required_memory_mb(uint_t num_qubits) const {
    size_t unit = std::log2(sizeof(std::complex<data_t>));
    size_t shift_mb = std::max<int_t>(0, num_qubits + unit - 20);
    return 1ULL << shift_mb;
  }
  // first vector of a container (= of a GPU's share of the shots) in batched mode; every chunk is its own group in
  // cache-blocking mode (one handle, one stream and one gate queue per chunk)
  bool top_of_group() { return batch_ ? batch_pos_ == 0 : true; }
  std::complex<data_t> *data() const { return nullptr; }  // amplitudes live in HBM
  void *device_data() const { void *p = nullptr; flush(); ck(b200sv_device_ptr(Hs(), &p)); return p; }

  void set_json_chop_threshold(double t) { json_chop_threshold_ = t; }
  double get_json_chop_threshold() { return json_chop_threshold_; }
  void set_omp_threads(int n) { omp_threads_ = n > 0 ? n : 1; }
  uint_t get_omp_threads() { return omp_threads_; }
  void set_omp_threshold(int n) { omp_threshold_ = n; }
  uint_t get_omp_threshold() { return omp_threshold_; }
  void set_num_threads_per_group(int) {}
  void cuStateVec_enable(bool) {}
  void set_target_gpus(reg_t &t) {
    target_gpus_ = t;
    if (!t.empty()) { device() = (int)t[0]; device_ = (int)t[0]; }
  }
  void set_sample_measure_index_size(int n) { sample_measure_index_size_ = n; }
  int get_sample_measure_index_size() { return sample_measure_index_size_; }
  void set_max_matrix_bits(int_t) {}
  void set_max_sampling_shots(int_t) {}
  void synchronize(void) { if (batch_ || h_) { flush(); ck(b200sv_synchronize(Hs())); } }
  bool support_global_indexing(void) { return false; }
  virtual bool batched_optimization_supported(void) { return true; }

  // enable_batch (qubitvector_thrust.hpp:1184-1198): returns the previous setting
  virtual bool enable_batch(bool flg) const {
    if (!batch_) return false;
    const bool prev = batch_->on;
    if (prev != flg) batch_->queue.flush(batch_->h);
    batch_->on = flg;
    return prev;
  }

  //---------------------------------------------------------------- chunks / containers
  // chunk_setup (qubitvector.hpp:1045; thrust :860-935).  chunk_bits == num_qubits with several local
  // "chunks" is the multi-shot container of BatchShotsExecutor (chunk_manager.hpp:223-264).
  uint_t chunk_setup(int chunk_bits, int num_qubits, uint_t chunk_index, uint_t num_local_chunks) {
    chunk_index_ = chunk_index;
    drop_batch();
    group_.reset();
    const std::vector<int> devs = b200detail::resolve_devices(target_gpus_, device());
    if (chunk_bits == num_qubits && num_local_chunks > 1) {
      // multi-shot containers, one per target GPU; single precision included (aer_controller.hpp:642-646 batches
      // QubitVectorThrust<float> too)
      release();
      const size_t G = std::min<size_t>(devs.size(), (size_t)num_local_chunks);
      batch_set_ = std::make_shared<b200detail::BatchSet>();
      for (size_t g = 0; g < G; g++) {
        const uint_t lo = num_local_chunks * g / G, hi = num_local_chunks * (g + 1) / G;
        auto b = std::make_shared<b200detail::Batch>();
        ck(b200sv_create(&b->h, chunk_bits, (int64_t)(hi - lo), sizeof(data_t) == 8 ? B200SV_F64 : B200SV_F32, devs[g]));
        b->nq = chunk_bits;
        b->nstates = hi - lo;
        b->first_index = chunk_index + lo;
        b->cregs.resize(hi - lo);
        batch_set_->parts.push_back(b);
        batch_set_->first.push_back(chunk_index + lo);
      }
      batch_set_->first.push_back(chunk_index + num_local_chunks);
      batch_ = batch_set_->parts[0];
      batch_pos_ = 0;
      num_qubits_ = chunk_bits;
      data_size_ = 1ull << chunk_bits;
    } else if (chunk_bits < num_qubits && num_local_chunks > 1) {
      // cache blocking: chunk i of the register lives on devs[i * G / C]
      auto g = std::make_shared<b200detail::ChunkGroup>();
      g->devices = devs;
      g->first_index = chunk_index;
      g->num_chunks = num_local_chunks;
      if (h_ && g->device_of(chunk_index) != my_device()) release();
      group_ = g;
    }
    return num_local_chunks;
  }
  uint_t chunk_setup(QubitVectorB200<data_t> &base, const uint_t chunk_index) {
    chunk_index_ = chunk_index;
    drop_batch();
    group_.reset();
    if (base.batch_set_) {
      release();
      const auto &bs = *base.batch_set_;
      size_t p = 0;
      while (p + 1 < bs.parts.size() && chunk_index >= bs.first[p + 1]) p++;
      batch_ = bs.parts[p];
      batch_pos_ = chunk_index - batch_->first_index;
      num_qubits_ = batch_->nq;
      data_size_ = 1ull << batch_->nq;
    } else if (base.group_) {
      if (h_ && base.group_->device_of(chunk_index) != my_device()) release();
      group_ = base.group_;
    }
    return 0;
  }
  uint_t chunk_index(void) { return chunk_index_; }
  bool fetch_chunk(void) const { return true; }
  void release_chunk(bool = true) const {}
  void enter_register_blocking(const reg_t &) {}
  void leave_register_blocking(void) {}
  std::complex<data_t> *send_buffer(uint_t &size_in_byte) { size_in_byte = 0; throw std::runtime_error("QubitVectorB200: MPI buffers are not supported (use NCCL sharding)"); }
  std::complex<data_t> *recv_buffer(uint_t &size_in_byte) { size_in_byte = 0; throw std::runtime_error("QubitVectorB200: MPI buffers are not supported (use NCCL sharding)"); }
  void release_send_buffer(void) const {}
  void release_recv_buffer(void) const {}

  // apply_chunk_swap(qubits, chunk, write_back): qubitvector.hpp:1753-1790.  The partner chunk may live on another
  // GPU: the kernels reach it over NVLink peer access (enabled by the library on first use).
  void apply_chunk_swap(const reg_t &qubits, QubitVectorB200<data_t> &src, bool write_back = true) {
    uint_t q0 = qubits[qubits.size() - 2], q1 = qubits[qubits.size() - 1];
    if (q0 > q1) std::swap(q0, q1);
    synchronize();
    src.synchronize();
    void *peer = src.device_data();
    if (q0 >= num_qubits_) {  // both global: exchange whole chunks on the device
      if (write_back) ck(b200sv_swap_range_peer(Hs(), 0, peer, 0, data_size_));
      else ck(b200sv_copy_range_peer(Hs(), 0, peer, 0, data_size_));  // copy only (qubitvector.hpp:1771-1775)
      synchronize();
      return;
    }
    // this (lower chunk: its q0=1 half) <-> src (q0=0 half); the kernel moves both directions
    const bool this_is_upper = !(chunk_index_ < src.chunk_index_);
    ck(b200sv_chunk_swap_peer(Hs(), (int)q0, peer, this_is_upper ? 1 : 0, 0));
    ck(b200sv_chunk_swap_peer(Hs(), (int)q0, peer, this_is_upper ? 1 : 0, 1));
    synchronize();
  }
  // MPI-only overload (parallel_state_executor.hpp:1331, inside #ifdef AER_MPI): inter-process exchange is NCCL / NVLink
  // here (b200sv_sharded_*), so this entry point is unreachable in a build without AER_MPI.
  void apply_chunk_swap(const reg_t &, uint_t) { throw std::runtime_error("QubitVectorB200: remote (MPI) chunk swap is not supported"); }
  // apply_chunk_swap(chunk, dest_offset, src_offset, size) (qubitvector.hpp:1824-1840): the sub-block shuffle of
  // ParallelStateExecutor::apply_multi_chunk_swap (:1474-1500) -- `size` amplitudes of this chunk trade places with
  // `size` amplitudes of `src`, in place on the device(s).
  void apply_chunk_swap(QubitVectorB200<data_t> &src, uint_t dest_offset, uint_t src_offset, uint_t size) {
    if (src.chunk_index_ == chunk_index_)
      throw std::runtime_error("QubitVectorB200: receive-buffer copy belongs to the MPI path, which is not supported");
    synchronize();
    src.synchronize();
    ck(b200sv_swap_range_peer(Hs(), dest_offset, src.device_data(), src_offset, size));
    synchronize();
  }

  //---------------------------------------------------------------- data
  void zero() { if (idle()) return; drop_queue(); ck(b200sv_zero(H())); }
  void initialize() { if (idle()) return; drop_queue(); ck(b200sv_initialize(H())); }
  void initialize(const QubitVectorB200<data_t> &obj) {
    set_num_qubits(obj.num_qubits_);
    drop_queue();  // whatever was queued would act on data that is overwritten now
    const_cast<QubitVectorB200<data_t> &>(obj).synchronize();
    ck(b200sv_copy_range_peer(Hs(), 0, obj.device_data(), 0, data_size_));
    ck(b200sv_synchronize(Hs()));
  }
  template <typename list_t> void initialize_from_vector(const list_t &vec) {
    if (data_size_ != vec.size()) throw std::runtime_error("QubitVector::initialize input vector is incorrect length");
    std::vector<std::complex<data_t>> tmp(vec.size());
    for (size_t i = 0; i < vec.size(); i++) tmp[i] = std::complex<data_t>(vec[i]);
    upload_all(tmp.data());
  }
  void initialize_from_vector(std::vector<std::complex<data_t>> &&vec) { initialize_from_data(vec.data(), vec.size()); }
  void initialize_from_vector(AER::Vector<std::complex<data_t>> &&vec) { initialize_from_data(vec.data(), vec.size()); }
  virtual void move_from_vector(AER::Vector<std::complex<data_t>> &&vec) { initialize_from_data(vec.data(), vec.size()); }
  void initialize_from_data(const std::complex<data_t> *data, const size_t num_states) {
    if (data_size_ != num_states) throw std::runtime_error("QubitVector::initialize input vector is incorrect length");
    upload_all(data);
  }
  void initialize_component(const reg_t &qubits, const cvector_t<double> &state) {
    for_target([&](b200sv_handle h) { ck(b200sv_initialize_component(h, qubits.data(), (int)qubits.size(), (const double *)state.data())); });
  }
  cvector_t<data_t> vector() const {
    cvector_t<data_t> ret(data_size_);
    flush(); ck(b200sv_download(Hs(), ret.data(), 0, data_size_));
    return ret;
  }
  AER::Vector<std::complex<data_t>> copy_to_vector() const {
    AER::Vector<std::complex<data_t>> ret(data_size_, false);
    flush(); ck(b200sv_download(Hs(), ret.data(), 0, data_size_));
    return ret;
  }
  AER::Vector<std::complex<data_t>> move_to_vector() { return copy_to_vector(); }
  cdict_t<data_t> vector_ket(double epsilon = 0) const { return Utils::vec2ket(vector(), epsilon, 16); }
  json_t json() const {
    auto v = vector();
    json_t js = json_t(data_size_, json_t(2, 0.));
    for (size_t j = 0; j < data_size_; j++) {
      if (std::abs(v[j].real()) > json_chop_threshold_) js[j][0] = v[j].real();
      if (std::abs(v[j].imag()) > json_chop_threshold_) js[j][1] = v[j].imag();
    }
    return js;
  }
  std::complex<data_t> get_state(uint_t pos) const {
    std::complex<data_t> v;
    flush(); ck(b200sv_download(Hs(), &v, pos, 1));
    return v;
  }
  void set_state(uint_t pos, std::complex<data_t> &val) { flush(); ck(b200sv_upload(Hs(), &val, pos, 1)); }
  void checkpoint() { flush(); ck(b200sv_checkpoint(Hs())); }
  void revert(bool keep) { flush(); ck(b200sv_revert(Hs(), keep ? 1 : 0)); }
  std::complex<double> inner_product() const {
    double re, im;
    flush(); ck(b200sv_inner_product(Hs(), &re, &im));
    return {re, im};
  }

  //---------------------------------------------------------------- classical registers (batched mode keeps them per shot)
  virtual void initialize_creg(uint_t num_memory, uint_t num_register) {
    if (batch_) batch_->cregs[batch_pos_].initialize(num_memory, num_register);
  }
  virtual void initialize_creg(uint_t num_memory, uint_t num_register, const std::string &memory_hex,
                               const std::string &register_hex) {
    if (batch_) batch_->cregs[batch_pos_].initialize(num_memory, num_register, memory_hex, register_hex);
  }
  // read_measured_data (qubitvector_thrust.hpp:2463-2484): this shot's bits into the State's creg
  template <typename storage_t> void read_measured_data(storage_t &creg) { read_creg(creg); }
  virtual void set_conditional(int_t reg) { if (batch_ && !idle()) batch_->cond_reg = reg; }
  virtual int_t set_batched_system_conditional(int_t, reg_t &) { return -1; }
  virtual void apply_bfunc(const Operations::Op &op) {  // bfunc_kernel (:3180)
    if (!batch_ || idle()) return;
    for_states([&](size_t s) { batch_->cregs[s].apply_bfunc(op); });
  }
  virtual void apply_roerror(const Operations::Op &op, std::vector<RngEngine> &rng) {  // roerror_kernel (:3302)
    if (!batch_ || idle()) return;
    for_states([&](size_t s) { batch_->cregs[s].apply_roerror(op, rng[s]); });
  }

  //---------------------------------------------------------------- gates (qubitvector.hpp:225-294)
  void apply_matrix(const reg_t &qubits, const cvector_t<double> &mat) {
    if (idle()) return;
    if (enqueue(qubits, mat.data(), mat.size())) return;
    for_target([&](b200sv_handle h) { ck(b200sv_apply_matrix(h, qubits.data(), (int)qubits.size(), (const double *)mat.data())); });
  }
  void apply_multiplexer(const reg_t &control_qubits, const reg_t &target_qubits, const cvector_t<double> &mat) {
    for_target([&](b200sv_handle h) {
      ck(b200sv_apply_multiplexer(h, control_qubits.data(), (int)control_qubits.size(), target_qubits.data(),
                                  (int)target_qubits.size(), (const double *)mat.data()));
    });
  }
  void apply_diagonal_matrix(const reg_t &qubits, const cvector_t<double> &mat) {
    if (idle()) return;
    if (ride_queue(qubits) && mat.size() == (1ull << qubits.size())) {  // ride on the tile passes
      const size_t dim = mat.size();
      std::vector<std::complex<double>> full(dim * dim, 0.0);
      for (size_t i = 0; i < dim; i++) full[i + dim * i] = mat[i];
      if (enqueue(qubits, full.data(), full.size())) return;
    }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_diagonal(h, qubits.data(), (int)qubits.size(), (const double *)mat.data())); });
  }
  void apply_permutation_matrix(const reg_t &qubits, const std::vector<std::pair<uint_t, uint_t>> &pairs) {
    std::vector<uint64_t> flat;
    for (auto &p : pairs) { flat.push_back(p.first); flat.push_back(p.second); }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_permutation(h, qubits.data(), (int)qubits.size(), flat.data(), (int)pairs.size())); });
  }
  void apply_mcx(const reg_t &qubits) {
    if (idle()) return;
    if (ride_queue(qubits)) {  // x / cx as dense gates share HBM passes with their neighbours (and the sampled noise)
      static const std::complex<double> X[4] = {0, 1, 1, 0};
      static const std::complex<double> CX[16] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0};  // control = qubits[0]
      if (enqueue(qubits, qubits.size() == 1 ? X : CX, qubits.size() == 1 ? 4 : 16)) return;
    }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_mcx(h, qubits.data(), (int)qubits.size())); });
  }
  void apply_mcy(const reg_t &qubits) {
    if (idle()) return;
    if (ride_queue(qubits)) {
      const std::complex<double> I(0, 1);
      const std::complex<double> Y[4] = {0, I, -I, 0};
      std::complex<double> CY[16] = {};  // control = qubits[0], target = qubits[1]; column-major, index = q0 + 2 q1
      CY[0] = 1; CY[2 + 4 * 2] = 1; CY[3 + 4 * 1] = I; CY[1 + 4 * 3] = -I;
      if (enqueue(qubits, qubits.size() == 1 ? Y : CY, qubits.size() == 1 ? 4 : 16)) return;
    }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_mcy(h, qubits.data(), (int)qubits.size())); });
  }
  void apply_mcphase(const reg_t &qubits, const std::complex<double> phase) {
    if (idle()) return;
    if (ride_queue(qubits)) {
      std::complex<double> d[16] = {};
      const size_t dim = 1ull << qubits.size();
      for (size_t i = 0; i < dim; i++) d[i + dim * i] = 1.0;
      d[dim * dim - 1] = phase;
      if (enqueue(qubits, d, dim * dim)) return;
    }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_mcphase(h, qubits.data(), (int)qubits.size(), phase.real(), phase.imag())); });
  }
  void apply_mcu(const reg_t &qubits, const cvector_t<double> &mat) {
    if (idle()) return;
    // an uncontrolled 2x2 is a plain 1-qubit matrix (qubitvector.hpp:1676-1680), a singly-controlled one a 4x4 with
    // U in the control = 1 block: both ride on the gate queue
    if (qubits.size() == 1 && (ride_queue(qubits) || !(mat[1] == 0.0 && mat[2] == 0.0)) && enqueue(qubits, mat.data(), mat.size())) return;
    if (qubits.size() == 2 && mat.size() == 4 && ride_queue(qubits)) {
      std::complex<double> CU[16] = {};  // control = qubits[0], target = qubits[1]; column-major, index = q0 + 2 q1
      CU[0] = 1; CU[2 + 4 * 2] = 1;
      for (int tr = 0; tr < 2; tr++)
        for (int tc = 0; tc < 2; tc++) CU[(1 + 2 * tr) + 4 * (1 + 2 * tc)] = mat[tr + 2 * tc];
      if (enqueue(qubits, CU, 16)) return;
    }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_mcu(h, qubits.data(), (int)qubits.size(), (const double *)mat.data())); });
  }
  void apply_mcswap(const reg_t &qubits) {
    if (idle()) return;
    if (qubits.size() == 2 && ride_queue(qubits)) {
      static const std::complex<double> SWAP[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
      if (enqueue(qubits, SWAP, 16)) return;
    }
    for_target([&](b200sv_handle h) { ck(b200sv_apply_mcswap(h, qubits.data(), (int)qubits.size())); });
  }
  void apply_multi_swaps(const reg_t &qubits) {  // pairs of qubits, qubitvector.hpp:1843-1876
    for (size_t i = 0; i + 1 < qubits.size(); i += 2) apply_mcswap({qubits[i], qubits[i + 1]});
  }
  void apply_rotation(const reg_t &qubits, const Rotation r, const double theta) {  // qubitvector.hpp:1723-1750
    switch (r) {
    case Rotation::x: apply_mcu(qubits, Linalg::VMatrix::rx(theta)); break;
    case Rotation::y: apply_mcu(qubits, Linalg::VMatrix::ry(theta)); break;
    case Rotation::z: apply_mcu(qubits, Linalg::VMatrix::rz(theta)); break;
    case Rotation::xx: apply_matrix(qubits, Linalg::VMatrix::rxx(theta)); break;
    case Rotation::yy: apply_matrix(qubits, Linalg::VMatrix::ryy(theta)); break;
    case Rotation::zz: apply_diagonal_matrix(qubits, Linalg::VMatrix::rzz_diag(theta)); break;
    case Rotation::zx: apply_matrix(qubits, Linalg::VMatrix::rzx(theta)); break;
    default: throw std::invalid_argument("QubitVector::invalid rotation axis.");
    }
  }
  void apply_pauli(const reg_t &qubits, const std::string &pauli, const complex_t &coeff = 1) {
    for_target([&](b200sv_handle h) { ck(b200sv_apply_pauli(h, qubits.data(), (int)qubits.size(), pauli.c_str(), coeff.real(), coeff.imag())); });
  }

  //---------------------------------------------------------------- reductions (qubitvector.hpp:302-411); a vector's own state
  virtual double probability(const uint_t outcome) const { return std::norm(std::complex<double>(get_state(outcome))); }
  virtual std::vector<double> probabilities() const {
    reg_t all(num_qubits_);
    for (size_t i = 0; i < num_qubits_; i++) all[i] = i;
    return probabilities(all);
  }
  virtual std::vector<double> probabilities(const reg_t &qubits) const {
    std::vector<double> p(1ull << qubits.size());
    flush(); ck(b200sv_probabilities(Hs(), qubits.data(), (int)qubits.size(), p.data()));
    return p;
  }
  virtual reg_t sample_measure(const std::vector<double> &rnds) const {
    reg_t s(rnds.size());
    flush(); ck(b200sv_sample_measure(Hs(), rnds.data(), (int64_t)rnds.size(), s.data()));
    return s;
  }
  double norm() const {
    if (batch_ && batch_->on) {  // per-vector partial sums contract: only the first vector reports (thrust :1974-1987)
      if (batch_pos_ != 0) return 0.0;
      std::vector<double> v(batch_->nstates);
      flush(); ck(b200sv_norm(batch_->h, v.data()));
      double s = 0;
      for (double x : v) s += x;
      return s;
    }
    double v; flush(); ck(b200sv_norm(Hs(), &v)); return v;
  }
  double norm(const uint_t qubit, const cvector_t<double> &mat) const { return norm(reg_t({qubit}), mat); }
  double norm(const reg_t &qubits, const cvector_t<double> &mat) const {
    double v;
    flush(); ck(b200sv_norm_matrix(Hs(), qubits.data(), (int)qubits.size(), (const double *)mat.data(), &v));
    return v;
  }
  double norm_diagonal(const uint_t qubit, const cvector_t<double> &mat) const { return norm_diagonal(reg_t({qubit}), mat); }
  double norm_diagonal(const reg_t &qubits, const cvector_t<double> &mat) const {
    const size_t dim = 1ull << qubits.size();
    cvector_t<double> full(dim * dim, 0.);
    for (size_t i = 0; i < dim; i++) full[i + dim * i] = mat[i];
    return norm(qubits, full);
  }
  double expval_pauli(const reg_t &qubits, const std::string &pauli, const complex_t initial_phase = 1.0) const {
    double v;
    flush(); ck(b200sv_expval_pauli(Hs(), qubits.data(), (int)qubits.size(), pauli.c_str(), initial_phase.real(),
                                    initial_phase.imag(), &v));
    return v;
  }
  double expval_pauli(const reg_t &qubits, const std::string &pauli, const QubitVectorB200<data_t> &pair_chunk,
                      const uint_t z_count, const uint_t z_count_pair, const complex_t initial_phase = 1.0) const {
    double v;
    const_cast<QubitVectorB200<data_t> &>(pair_chunk).synchronize();
    void *pair = pair_chunk.device_data();
    flush(); ck(b200sv_expval_pauli_pair(Hs(), qubits.data(), (int)qubits.size(), pauli.c_str(), pair, z_count,
                                         z_count_pair, initial_phase.real(), initial_phase.imag(), &v));
    return v;
  }

  //---------------------------------------------------------------- batched-shot operations (first vector acts for the group)
  // Sampled Pauli noise, one op list per shot (qubitvector_thrust.hpp:2892-2945): the Paulis become per-state
  // codes of queued ops and ride on the same HBM passes as the surrounding gates.
  virtual void apply_batched_pauli_ops(const std::vector<std::vector<Operations::Op>> &ops) {
    if (!batch_ || idle()) return;
    const size_t S = batch_->nstates;
    std::vector<uint64_t> xm(S, 0), zm(S, 0);
    uint64_t touched = 0;
    for (size_t s = 0; s < ops.size() && s < S; s++) {
      for (const auto &op : ops[s]) {
        if (op.conditional && !batch_->cregs[s].check_conditional(op)) continue;
        if (op.name == "x") xm[s] ^= 1ull << op.qubits[0];
        else if (op.name == "z") zm[s] ^= 1ull << op.qubits[0];
        else if (op.name == "y") { xm[s] ^= 1ull << op.qubits[0]; zm[s] ^= 1ull << op.qubits[0]; }
        else if (op.name == "pauli") {
          uint_t px, pz, ny, xmax;
          std::tie(px, pz, ny, xmax) = pauli_masks_and_phase(op.qubits, op.string_params[0]);
          xm[s] ^= px; zm[s] ^= pz;
        }
      }
      touched |= xm[s] | zm[s];
    }
    if (!touched) return;
    auto &Q = batch_->queue;
    for (uint_t q = 0; q < num_qubits_; q++) {
      if (!((touched >> q) & 1)) continue;
      Q.kind.push_back(3);
      Q.slot.push_back(Q.nslots);
      Q.qubits.push_back(q); Q.qubits.push_back(0);
      Q.mats.resize(Q.mats.size() + 32, 0.0);
      const size_t off = Q.codes.size();
      Q.codes.resize(off + S, 0);
      for (size_t s = 0; s < S; s++) {
        const int x = (int)((xm[s] >> q) & 1), z = (int)((zm[s] >> q) & 1);
        Q.codes[off + s] = (uint8_t)(x ? (z ? 2 : 1) : (z ? 3 : 0));
      }
      Q.nslots++;
    }
    if (!queue_enabled() || Q.kind.size() >= batch_queue_limit()) Q.flush(batch_->h);
  }
  // ops queued before a batched container flushes on its own: the flush is asynchronous, so the executor's host work
  // for the following ops (noise sampling for every shot) runs while the GPU works through the passes of this part
  static size_t batch_queue_limit() {
    static const size_t n = [] { const char *e = getenv("B200SV_BATCH_QUEUE_OPS"); return e && atoi(e) > 0 ? (size_t)atoi(e) : (size_t)512; }();
    return n;
  }
  // apply_batched_measure (qubitvector_thrust.hpp:2251-2330): r = rng[s].rand(); outcome = first i with
  // r < cumulative probability; collapse + renormalise; store bits.
  virtual void apply_batched_measure(const reg_t &qubits, std::vector<RngEngine> &rng, const reg_t &cmemory,
                                     const reg_t &cregs) {
    if (!batch_ || idle()) return;
    measure_impl(qubits, rng, &cmemory, &cregs, false);
  }
  // apply_batched_reset (qubitvector_thrust.hpp:2386-2460): measure without storing, then flip back to |0>
  virtual void apply_batched_reset(const reg_t &qubits, std::vector<RngEngine> &rng) {
    if (!batch_ || idle()) return;
    measure_impl(qubits, rng, nullptr, nullptr, true);
  }
  // apply_batched_kraus (qubitvector_thrust.hpp:3096-3177): per shot, r = rng.rand(); accumulate
  // p_j = ||K_j psi||^2 until it exceeds r; apply K_j / sqrt(p_j).
  void apply_batched_kraus(const reg_t &qubits, const std::vector<cmatrix_t> &kmats, std::vector<RngEngine> &rng) {
    if (!batch_ || idle()) return;
    const size_t S = batch_->nstates;
    std::vector<uint8_t> active = take_conditional();
    batch_->queue.flush(batch_->h);
    std::vector<double> r(S), accum(S, 0.0), p(S), scale(S, 1.0);
    std::vector<int> chosen(S, -1);
    for (size_t s = 0; s < S; s++) r[s] = rng[s].rand(0., 1.);
    const size_t msz = (size_t)1 << (2 * qubits.size());
    cvector_t<double> table(msz * kmats.size());
    for (size_t j = 0; j < kmats.size(); j++) {
      cvector_t<double> vmat = Utils::vectorize_matrix(kmats[j]);
      std::copy(vmat.begin(), vmat.end(), table.begin() + j * msz);
      const bool last = j + 1 == kmats.size();
      bool undecided = false;
      for (size_t s = 0; s < S && !undecided; s++) undecided = active[s] && chosen[s] < 0;
      if (!undecided) continue;  // every shot has its operator: no need to evaluate the remaining probabilities
      if (!last) ck(b200sv_norm_matrix(batch_->h, qubits.data(), (int)qubits.size(), (const double *)vmat.data(), p.data()));
      for (size_t s = 0; s < S; s++) {
        if (!active[s] || chosen[s] >= 0) continue;
        const double pj = last ? 1.0 - accum[s] : p[s];
        accum[s] += pj;
        if (last || accum[s] > r[s]) {
          chosen[s] = (int)j;
          scale[s] = 1.0 / std::sqrt(pj);
        }
      }
    }
    // one launch: every shot applies the operator its draw selected, renormalised
    ck(b200sv_apply_batched_matrix(batch_->h, qubits.data(), (int)qubits.size(), (const double *)table.data(),
                                   (int)kmats.size(), chosen.data(), scale.data()));
  }
  // per-parameter matrices (runtime parameter binding): apply_batched_matrix (qubitvector_thrust.hpp:1578-1611)
  void apply_batched_matrix(const reg_t &qubits, const cvector_t<double> &mat, const uint_t num_matrices,
                            const uint_t num_shots_per_matrix) {
    if (!batch_ || idle()) return;
    batch_->queue.flush(batch_->h);
    const size_t S = batch_->nstates;
    std::vector<int> index(S, -1);
    std::vector<double> scale(S, 1.0);
    for (size_t s = 0; s < S; s++) {
      const uint_t m = s / num_shots_per_matrix;
      if (m < num_matrices) index[s] = (int)m;
    }
    ck(b200sv_apply_batched_matrix(batch_->h, qubits.data(), (int)qubits.size(), (const double *)mat.data(),
                                   (int)num_matrices, index.data(), scale.data()));
  }
  void apply_batched_diagonal_matrix(const reg_t &qubits, const cvector_t<double> &mat, const uint_t num_matrices,
                                     const uint_t num_shots_per_matrix) {
    if (!batch_ || idle()) return;
    batch_->queue.flush(batch_->h);
    const size_t msize = 1ull << qubits.size();
    for (uint_t m = 0; m < num_matrices; m++) {
      b200sv_handle v = nullptr;
      ck(b200sv_create_view(&v, batch_->h, (int64_t)(m * num_shots_per_matrix), (int64_t)num_shots_per_matrix));
      const int rc = b200sv_apply_diagonal(v, qubits.data(), (int)qubits.size(), (const double *)(mat.data() + m * msize));
      b200sv_destroy(v);
      ck(rc);
    }
  }
  // batched_expval_pauli (qubitvector_thrust.hpp:2683-2713; batched_expval_*_func thrust_kernels.hpp:2472-2600):
  // val[s] += Re(param) * <P>_s  (and val[2s+1] += Im(param) * <P>_s when the variance is requested)
  void batched_expval_pauli(std::vector<double> &val, const reg_t &qubits, const std::string &pauli, bool variance,
                            std::complex<double> param, bool /*last*/, const complex_t initial_phase = 1.0) const {
    if (batch_ && batch_->on && batch_pos_ != 0) return;
    const size_t S = (batch_ && batch_->on) ? batch_->nstates : 1;
    if (val.empty()) val.assign(variance ? 2 * S : S, 0.0);
    std::vector<double> e(S);
    flush();
    ck(b200sv_expval_pauli((batch_ && batch_->on) ? batch_->h : Hs(), qubits.data(), (int)qubits.size(), pauli.c_str(),
                           initial_phase.real(), initial_phase.imag(), e.data()));
    for (size_t s = 0; s < S; s++) {
      if (variance) { val[2 * s] += param.real() * e[s]; val[2 * s + 1] += param.imag() * e[s]; }
      else val[s] += param.real() * e[s];
    }
  }

  //---------------------------------------------------------------- gate queue
  static bool queue_enabled() {
    static const bool on = [] { const char *e = getenv("B200SV_GATE_QUEUE"); return !(e && e[0] == '0'); }();
    return on;
  }
  void flush() const {
    if (batch_) batch_->queue.flush(batch_->h);
    else if (h_) queue_.flush(h_);
  }

protected:
  static void ck(int rc) { b200detail::ck(rc); }
  // handle the next call acts on: the container (batched, first vector), this vector's state view, or its own handle
  b200sv_handle H() const { return batch_ ? (batch_->on ? batch_->h : state_view(batch_pos_)) : h_; }
  // handle for operations that always concern THIS vector's state only (data access, reductions)
  b200sv_handle Hs() const { return batch_ ? state_view(batch_pos_) : h_; }
  bool idle() const { return batch_ && batch_->on && batch_pos_ != 0; }
  bool batch_queueing() const { return batch_ && batch_->on && batch_->cond_reg < 0 && queue_enabled(); }
  // may a <= 2-qubit special gate (x, cx, cy, cz/cp, swap, cu, small diagonal) be rewritten as a dense matrix and
  // queued?  Yes when the flush runs tile passes (>= 12 qubits in double, >= 13 in single precision, all qubits local): the gate then
  // shares an HBM pass with its neighbours instead of costing one of its own.
  bool ride_queue(const reg_t &qubits) const {
    if (!queue_enabled() || qubits.empty() || qubits.size() > 2) return false;
    for (const auto q : qubits)
      if (q >= num_qubits_) return false;
    if (batch_) return batch_queueing();
    return sizeof(data_t) == 8 ? num_qubits_ >= 12 : num_qubits_ >= 13;  // the sizes the tile passes take
  }
  b200sv_handle state_view(size_t s) const {
    if (s == batch_pos_) {
      if (!view_) ck(b200sv_create_view(&view_, batch_->h, (int64_t)s, 1));
      return view_;
    }
    if (tmp_view_ && tmp_view_state_ != s) { b200sv_destroy(tmp_view_); tmp_view_ = nullptr; }
    if (!tmp_view_) { ck(b200sv_create_view(&tmp_view_, batch_->h, (int64_t)s, 1)); tmp_view_state_ = s; }
    return tmp_view_;
  }
  // one-shot conditional (chunk_container.hpp:416-420): which states execute the next batched op
  std::vector<uint8_t> take_conditional() {
    const size_t S = batch_->nstates;
    std::vector<uint8_t> active(S, 1);
    if (batch_->cond_reg >= 0) {
      for (size_t s = 0; s < S; s++) {
        const auto &reg = batch_->cregs[s].creg_register();
        active[s] = reg.size() > (size_t)batch_->cond_reg && reg[reg.size() - batch_->cond_reg - 1] == '1';
      }
      batch_->cond_reg = -1;
    }
    return active;
  }
  template <typename F> void for_states(F f) {
    if (batch_->on) {
      std::vector<uint8_t> active = take_conditional();
      for (size_t s = 0; s < batch_->nstates; s++)
        if (active[s]) f(s);
    } else {
      f(batch_pos_);
    }
  }
  // run f on the handle(s) the call concerns, honouring the batched-mode contracts
  template <typename F> void for_target(F f) {
    if (!batch_) { flush(); f(h_); return; }
    if (!batch_->on) { batch_->queue.flush(batch_->h); f(state_view(batch_pos_)); return; }
    if (batch_pos_ != 0) return;
    batch_->queue.flush(batch_->h);
    if (batch_->cond_reg < 0) { f(batch_->h); return; }
    std::vector<uint8_t> active = take_conditional();
    for (size_t s = 0; s < batch_->nstates; s++)
      if (active[s]) f(state_view(s));
  }
  bool enqueue(const reg_t &qubits, const std::complex<double> *mat, size_t mat_size) {
    if (!queue_enabled() || qubits.size() < 1 || qubits.size() > 2 || mat_size != (1ull << (2 * qubits.size()))) return false;
    if (batch_) {
      if (!batch_->on || batch_->cond_reg >= 0) return false;
      batch_->queue.push_dense(qubits.data(), (int)qubits.size(), mat);
      if (batch_->queue.kind.size() >= batch_queue_limit()) batch_->queue.flush(batch_->h);
      return true;
    }
    queue_.push_dense(qubits.data(), (int)qubits.size(), mat);
    if (queue_.kind.size() >= 4096) queue_.flush(h_);
    return true;
  }
  void drop_queue() {
    if (!batch_) { queue_.clear(); return; }
    if (batch_->on && batch_pos_ == 0) batch_->queue.clear();
    else batch_->queue.flush(batch_->h);
  }
  void upload_all(const std::complex<data_t> *data) {
    flush();
    if (batch_ && batch_->on) {  // set_statevec in batched mode: every shot starts from the same vector
      if (batch_pos_ != 0) return;
      // one host -> device transfer, then stream-ordered device copies into the other shots
      ck(b200sv_upload(batch_->h, data, 0, data_size_));
      void *base = nullptr;
      ck(b200sv_device_ptr(batch_->h, &base));
      for (size_t s = 1; s < batch_->nstates; s++) ck(b200sv_copy_range_peer(batch_->h, s << num_qubits_, base, 0, data_size_));
      return;
    }
    ck(b200sv_upload(Hs(), data, 0, data_size_));
  }
  void read_creg(ClassicalRegister &creg) {
    if (!batch_) return;
    creg.creg_memory() = batch_->cregs[batch_pos_].creg_memory();
    creg.creg_register() = batch_->cregs[batch_pos_].creg_register();
  }
  template <typename T> void read_creg(T &) {}
  void measure_impl(const reg_t &qubits, std::vector<RngEngine> &rng, const reg_t *cmemory, const reg_t *cregs, bool reset) {
    const size_t S = batch_->nstates, DIM = 1ull << qubits.size();
    std::vector<uint8_t> active = take_conditional();
    batch_->queue.flush(batch_->h);
    std::vector<double> scale(S, 1.0);
    std::vector<uint64_t> outcome(S, 0), masks(4 * S, 0);
    bool any_flip = false;
    auto record = [&](size_t s, size_t o) {
      outcome[s] = o;
      if (reset) {
        uint64_t x = 0;
        for (size_t j = 0; j < qubits.size(); j++)
          if ((o >> j) & 1) x |= 1ull << qubits[j];
        masks[4 * s] = x; masks[4 * s + 3] = x != 0;
        any_flip |= x != 0;
      } else {
        reg_t bits(qubits.size());
        for (size_t j = 0; j < qubits.size(); j++) bits[j] = (o >> j) & 1;
        batch_->cregs[s].store_measure(bits, *cmemory, *cregs);
      }
    };
    bool all_in_order = qubits.size() == num_qubits_;
    for (size_t j = 0; j < qubits.size() && all_in_order; j++) all_in_order = qubits[j] == j;
    if (all_in_order && qubits.size() > 10) {
      // every qubit, in order: "first outcome with r < cumulative probability" IS the sampler's definition,
      // so draw through sample_measure (no 2^n-per-shot probability table)
      std::vector<double> r(S, 0.0);
      for (size_t s = 0; s < S; s++) if (active[s]) r[s] = rng[s].rand();
      reg_t smp(S);
      ck(b200sv_sample_measure(batch_->h, r.data(), 1, smp.data()));
      for (size_t s = 0; s < S; s++) {
        if (!active[s]) continue;
        scale[s] = -1.0;  // normalise the surviving amplitude by its own modulus
        record(s, smp[s]);
      }
    } else {
      // marginal probabilities in slabs of states so that the table stays small
      const size_t slab = std::max<size_t>(1, std::min<size_t>(S, ((size_t)32 << 20) / DIM));
      std::vector<double> probs(slab * DIM);
      for (size_t s0 = 0; s0 < S; s0 += slab) {
        const size_t ns = std::min(slab, S - s0);
        if (ns == S) {
          ck(b200sv_probabilities(batch_->h, qubits.data(), (int)qubits.size(), probs.data()));
        } else {
          b200sv_handle v = nullptr;
          ck(b200sv_create_view(&v, batch_->h, (int64_t)s0, (int64_t)ns));
          const int rc = b200sv_probabilities(v, qubits.data(), (int)qubits.size(), probs.data());
          b200sv_destroy(v);
          ck(rc);
        }
        for (size_t i = 0; i < ns; i++) {
          const size_t s = s0 + i;
          if (!active[s]) continue;
          const double r = rng[s].rand();
          double total = 0, cum = 0;
          for (size_t o = 0; o < DIM; o++) total += probs[i * DIM + o];
          size_t pick = DIM - 1;
          for (size_t o = 0; o + 1 < DIM; o++) {
            cum += probs[i * DIM + o] / total;
            if (r < cum) { pick = o; break; }
          }
          scale[s] = 1.0 / std::sqrt(probs[i * DIM + pick]);
          record(s, pick);
        }
      }
    }
    ck(b200sv_collapse(batch_->h, qubits.data(), (int)qubits.size(), outcome.data(), scale.data(), active.data()));
    if (reset && any_flip) ck(b200sv_apply_batched_pauli(batch_->h, masks.data()));
  }
  void drop_batch() {
    if (view_) { b200sv_destroy(view_); view_ = nullptr; }
    if (tmp_view_) { b200sv_destroy(tmp_view_); tmp_view_ = nullptr; }
    batch_.reset();
    batch_set_.reset();
    batch_pos_ = 0;
  }
  void release() {
    queue_.clear();
    drop_batch();
    if (h_) { b200sv_destroy(h_); h_ = nullptr; }
  }
  b200sv_handle h_ = nullptr;
  size_t num_qubits_ = 0;
  size_t data_size_ = 0;
  uint_t chunk_index_ = 0;
  uint_t omp_threads_ = 1, omp_threshold_ = 14;
  int sample_measure_index_size_ = 10;
  double json_chop_threshold_ = 0;
  mutable b200detail::OpQueue queue_;
  std::shared_ptr<b200detail::Batch> batch_;
  std::shared_ptr<b200detail::BatchSet> batch_set_;   // first vector of a multi-shot allocation only
  std::shared_ptr<b200detail::ChunkGroup> group_;     // cache-blocking mode: placement of the register's chunks
  reg_t target_gpus_;
  int device_ = -1;
  size_t batch_pos_ = 0;
  mutable b200sv_handle view_ = nullptr, tmp_view_ = nullptr;
  mutable size_t tmp_view_state_ = 0;
};

}  // namespace QV
}  // namespace AER
#endif
