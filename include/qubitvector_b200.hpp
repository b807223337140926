// QubitVectorB200<data_t>: the reference-side adapter a maintainer adds to Aer.
//
// A header-only `statevec_t` for `AER::Statevector::State<statevec_t>`
// (src/simulators/statevector/statevector_state.hpp:100-101) exposing the same
// method set as AER::QV::QubitVector<data_t> (src/simulators/statevector/
// qubitvector.hpp:62-480 -- signatures mirrored one for one) and forwarding
// every amplitude operation to the C ABI of the B200 engine (include/b200sv.h).
// Errors come back as the std::runtime_error the executors already catch
// (src/simulators/circuit_executor.hpp:574,726).
//
// Needs the Aer source tree on the include path (it uses Aer's own types:
// reg_t, cvector_t, AER::Vector, Operations::Op, RngEngine, QV::Rotation).
// One object = one statevector (= one chunk in cache-blocking mode); like the
// CPU class it reports support_global_indexing() == false, so the State
// rewrites global-qubit diagonals/controls per chunk on the host
// (statevector_state.hpp:735-753) and chunk swaps arrive through
// apply_chunk_swap(qubits, other_chunk, write_back).
#ifndef _qv_qubit_vector_b200_hpp_
#define _qv_qubit_vector_b200_hpp_

#include <complex>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200sv.h"
#include "framework/json.hpp"
#include "framework/linalg/vector.hpp"
#include "framework/operations.hpp"
#include "framework/rng.hpp"
#include "framework/types.hpp"
#include "framework/utils.hpp"
#include "simulators/statevector/qubitvector.hpp"  // QV::Rotation, Linalg::VMatrix helpers

namespace AER {
namespace QV {

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
      o.h_ = nullptr; o.num_qubits_ = 0; o.data_size_ = 0;
    }
    return *this;
  }

  static std::string name() { return "statevector_b200"; }
  static int &device() { static int d = 0; return d; }

  //---------------------------------------------------------------- size / config
  virtual void set_num_qubits(size_t num_qubits) {
    if (h_ && num_qubits == num_qubits_) return;
    release();
    ck(b200sv_create(&h_, (int)num_qubits, 1, sizeof(data_t) == 8 ? B200SV_F64 : B200SV_F32, device()));
    num_qubits_ = num_qubits;
    data_size_ = 1ull << num_qubits;
  }
  virtual uint_t num_qubits() const { return num_qubits_; }
  uint_t size() const { return data_size_; }
  size_t required_memory_mb(uint_t num_qubits) const {
    size_t unit = std::log2(sizeof(std::complex<data_t>));
    size_t shift_mb = std::max<int_t>(0, num_qubits + unit - 20);
    return 1ULL << shift_mb;
  }
  bool top_of_group() { return true; }
  std::complex<data_t> *data() const { return nullptr; }  // amplitudes live in HBM
  void *device_data() const { void *p = nullptr; flush(); ck(b200sv_device_ptr(h_, &p)); return p; }

  void set_json_chop_threshold(double t) { json_chop_threshold_ = t; }
  double get_json_chop_threshold() { return json_chop_threshold_; }
  void set_omp_threads(int n) { omp_threads_ = n > 0 ? n : 1; }
  uint_t get_omp_threads() { return omp_threads_; }
  void set_omp_threshold(int n) { omp_threshold_ = n; }
  uint_t get_omp_threshold() { return omp_threshold_; }
  void set_num_threads_per_group(int) {}
  void cuStateVec_enable(bool) {}
  void set_target_gpus(reg_t &t) { if (!t.empty()) device() = (int)t[0]; }
  void set_sample_measure_index_size(int n) { sample_measure_index_size_ = n; }
  int get_sample_measure_index_size() { return sample_measure_index_size_; }
  void set_max_matrix_bits(int_t) {}
  void set_max_sampling_shots(int_t) {}
  void synchronize(void) { if (h_) { flush(); ck(b200sv_synchronize(h_)); } }
  virtual bool enable_batch(bool) const { return false; }
  bool support_global_indexing(void) { return false; }
  virtual bool batched_optimization_supported(void) { return false; }

  //---------------------------------------------------------------- chunks
  uint_t chunk_setup(int, int, uint_t chunk_index, uint_t num_local_chunks) {
    chunk_index_ = chunk_index;
    return num_local_chunks;
  }
  uint_t chunk_setup(QubitVectorB200<data_t> &, const uint_t chunk_index) {
    chunk_index_ = chunk_index;
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

  // apply_chunk_swap(qubits, chunk, write_back): qubitvector.hpp:1753-1790
  void apply_chunk_swap(const reg_t &qubits, QubitVectorB200<data_t> &src, bool write_back = true) {
    uint_t q0 = qubits[qubits.size() - 2], q1 = qubits[qubits.size() - 1];
    if (q0 > q1) std::swap(q0, q1);
    if (q0 >= num_qubits_) {  // both global: exchange (or copy) whole chunks
      const size_t bytes = data_size_ * sizeof(std::complex<data_t>);
      synchronize(); src.synchronize();
      std::vector<char> a(bytes), b(bytes);  // rare path (X on a global qubit): staged through the host
      ck(b200sv_download(src.h_, b.data(), 0, data_size_));
      if (write_back) { flush(); ck(b200sv_download(h_, a.data(), 0, data_size_)); ck(b200sv_upload(src.h_, a.data(), 0, data_size_)); }
      flush(); ck(b200sv_upload(h_, b.data(), 0, data_size_));
      return;
    }
    // this (lower chunk: its q0=1 half) <-> src (q0=0 half); the kernel moves both directions
    const bool this_is_upper = !(chunk_index_ < src.chunk_index_);
    src.synchronize();
    flush(); ck(b200sv_chunk_swap_peer(h_, (int)q0, src.device_data(), this_is_upper ? 1 : 0, 0));
    flush(); ck(b200sv_chunk_swap_peer(h_, (int)q0, src.device_data(), this_is_upper ? 1 : 0, 1));
    synchronize();
  }
  void apply_chunk_swap(const reg_t &, uint_t) { throw std::runtime_error("QubitVectorB200: remote (MPI) chunk swap is not supported"); }
  void apply_chunk_swap(QubitVectorB200<data_t> &, uint_t, uint_t, uint_t) { throw std::runtime_error("QubitVectorB200: multi chunk swap is not supported"); }

  //---------------------------------------------------------------- data
  void zero() { flush(); ck(b200sv_zero(h_)); }
  void initialize() { flush(); ck(b200sv_initialize(h_)); }
  void initialize(const QubitVectorB200<data_t> &obj) {
    set_num_qubits(obj.num_qubits_);
    auto v = obj.copy_to_vector();
    flush(); ck(b200sv_upload(h_, v.data(), 0, data_size_));
  }
  template <typename list_t> void initialize_from_vector(const list_t &vec) {
    if (data_size_ != vec.size()) throw std::runtime_error("QubitVector::initialize input vector is incorrect length");
    std::vector<std::complex<data_t>> tmp(vec.size());
    for (size_t i = 0; i < vec.size(); i++) tmp[i] = std::complex<data_t>(vec[i]);
    flush(); ck(b200sv_upload(h_, tmp.data(), 0, data_size_));
  }
  void initialize_from_vector(std::vector<std::complex<data_t>> &&vec) { initialize_from_data(vec.data(), vec.size()); }
  void initialize_from_vector(AER::Vector<std::complex<data_t>> &&vec) { initialize_from_data(vec.data(), vec.size()); }
  virtual void move_from_vector(AER::Vector<std::complex<data_t>> &&vec) { initialize_from_data(vec.data(), vec.size()); }
  void initialize_from_data(const std::complex<data_t> *data, const size_t num_states) {
    if (data_size_ != num_states) throw std::runtime_error("QubitVector::initialize input vector is incorrect length");
    flush(); ck(b200sv_upload(h_, data, 0, data_size_));
  }
  virtual void initialize_creg(uint_t, uint_t) {}
  virtual void initialize_creg(uint_t, uint_t, const std::string &, const std::string &) {}
  void initialize_component(const reg_t &qubits, const cvector_t<double> &state) {
    flush(); ck(b200sv_initialize_component(h_, qubits.data(), (int)qubits.size(), (const double *)state.data()));
  }
  cvector_t<data_t> vector() const {
    cvector_t<data_t> ret(data_size_);
    flush(); ck(b200sv_download(h_, ret.data(), 0, data_size_));
    return ret;
  }
  AER::Vector<std::complex<data_t>> copy_to_vector() const {
    AER::Vector<std::complex<data_t>> ret(data_size_, false);
    flush(); ck(b200sv_download(h_, ret.data(), 0, data_size_));
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
    flush(); ck(b200sv_download(h_, &v, pos, 1));
    return v;
  }
  void set_state(uint_t pos, std::complex<data_t> &val) { flush(); ck(b200sv_upload(h_, &val, pos, 1)); }
  void checkpoint() { flush(); ck(b200sv_checkpoint(h_)); }
  void revert(bool keep) { flush(); ck(b200sv_revert(h_, keep ? 1 : 0)); }
  std::complex<double> inner_product() const {
    double re, im;
    flush(); ck(b200sv_inner_product(h_, &re, &im));
    return {re, im};
  }

  //---------------------------------------------------------------- gates (qubitvector.hpp:225-294)
  void apply_matrix(const reg_t &qubits, const cvector_t<double> &mat) {
    if (enqueue(qubits, mat)) return;
    flush();
    ck(b200sv_apply_matrix(h_, qubits.data(), (int)qubits.size(), (const double *)mat.data()));
  }
  void apply_multiplexer(const reg_t &control_qubits, const reg_t &target_qubits, const cvector_t<double> &mat) {
    flush(); ck(b200sv_apply_multiplexer(h_, control_qubits.data(), (int)control_qubits.size(), target_qubits.data(),
                                (int)target_qubits.size(), (const double *)mat.data()));
  }
  void apply_diagonal_matrix(const reg_t &qubits, const cvector_t<double> &mat) {
    flush(); ck(b200sv_apply_diagonal(h_, qubits.data(), (int)qubits.size(), (const double *)mat.data()));
  }
  void apply_permutation_matrix(const reg_t &qubits, const std::vector<std::pair<uint_t, uint_t>> &pairs) {
    std::vector<uint64_t> flat;
    for (auto &p : pairs) { flat.push_back(p.first); flat.push_back(p.second); }
    flush(); ck(b200sv_apply_permutation(h_, qubits.data(), (int)qubits.size(), flat.data(), (int)pairs.size()));
  }
  void apply_mcx(const reg_t &qubits) { flush(); ck(b200sv_apply_mcx(h_, qubits.data(), (int)qubits.size())); }
  void apply_mcy(const reg_t &qubits) { flush(); ck(b200sv_apply_mcy(h_, qubits.data(), (int)qubits.size())); }
  void apply_mcphase(const reg_t &qubits, const std::complex<double> phase) {
    flush(); ck(b200sv_apply_mcphase(h_, qubits.data(), (int)qubits.size(), phase.real(), phase.imag()));
  }
  void apply_mcu(const reg_t &qubits, const cvector_t<double> &mat) {
    // an uncontrolled, non-diagonal 2x2 is a plain 1-qubit matrix (qubitvector.hpp:1676-1680): queue it
    if (qubits.size() == 1 && !(mat[1] == 0.0 && mat[2] == 0.0) && enqueue(qubits, mat)) return;
    flush();
    ck(b200sv_apply_mcu(h_, qubits.data(), (int)qubits.size(), (const double *)mat.data()));
  }
  void apply_mcswap(const reg_t &qubits) { flush(); ck(b200sv_apply_mcswap(h_, qubits.data(), (int)qubits.size())); }
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
    flush(); ck(b200sv_apply_pauli(h_, qubits.data(), (int)qubits.size(), pauli.c_str(), coeff.real(), coeff.imag()));
  }

  //---------------------------------------------------------------- reductions (qubitvector.hpp:302-411)
  virtual double probability(const uint_t outcome) const { return std::norm(std::complex<double>(get_state(outcome))); }
  virtual std::vector<double> probabilities() const {
    reg_t all(num_qubits_);
    for (size_t i = 0; i < num_qubits_; i++) all[i] = i;
    return probabilities(all);
  }
  virtual std::vector<double> probabilities(const reg_t &qubits) const {
    std::vector<double> p(1ull << qubits.size());
    flush(); ck(b200sv_probabilities(h_, qubits.data(), (int)qubits.size(), p.data()));
    return p;
  }
  virtual reg_t sample_measure(const std::vector<double> &rnds) const {
    reg_t s(rnds.size());
    flush(); ck(b200sv_sample_measure(h_, rnds.data(), (int64_t)rnds.size(), s.data()));
    return s;
  }
  double norm() const { double v; flush(); ck(b200sv_norm(h_, &v)); return v; }
  double norm(const uint_t qubit, const cvector_t<double> &mat) const { return norm(reg_t({qubit}), mat); }
  double norm(const reg_t &qubits, const cvector_t<double> &mat) const {
    double v;
    flush(); ck(b200sv_norm_matrix(h_, qubits.data(), (int)qubits.size(), (const double *)mat.data(), &v));
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
    flush(); ck(b200sv_expval_pauli(h_, qubits.data(), (int)qubits.size(), pauli.c_str(), initial_phase.real(),
                           initial_phase.imag(), &v));
    return v;
  }
  double expval_pauli(const reg_t &qubits, const std::string &pauli, const QubitVectorB200<data_t> &pair_chunk,
                      const uint_t z_count, const uint_t z_count_pair, const complex_t initial_phase = 1.0) const {
    double v;
    pair_chunk.sync_const();
    flush(); ck(b200sv_expval_pauli_pair(h_, qubits.data(), (int)qubits.size(), pauli.c_str(), pair_chunk.device_data(), z_count,
                                z_count_pair, initial_phase.real(), initial_phase.imag(), &v));
    return v;
  }

  //---------------------------------------------------------------- batched-shot hooks (not enabled yet: CPU-class stubs)
  virtual void apply_bfunc(const Operations::Op &) {}
  virtual void set_conditional(int_t) {}
  virtual void apply_roerror(const Operations::Op &, std::vector<RngEngine> &) {}
  virtual void apply_batched_measure(const reg_t &, std::vector<RngEngine> &, const reg_t &, const reg_t &) {}
  virtual void apply_batched_reset(const reg_t &, std::vector<RngEngine> &) {}
  template <typename storage_t> void read_measured_data(storage_t &) {}
  virtual int_t set_batched_system_conditional(int_t, reg_t &) { return -1; }
  virtual void apply_batched_pauli_ops(const std::vector<std::vector<Operations::Op>> &) {}
  void apply_batched_kraus(const reg_t &, const std::vector<cmatrix_t> &, std::vector<RngEngine> &) {}
  void apply_batched_matrix(const reg_t &, const cvector_t<double> &, const uint_t, const uint_t) {}
  void apply_batched_diagonal_matrix(const reg_t &, const cvector_t<double> &, const uint_t, const uint_t) {}
  void batched_expval_pauli(std::vector<double> &, const reg_t &, const std::string &, bool, std::complex<double>, bool,
                            const complex_t = 1.0) const {}

  //---------------------------------------------------------------- gate queue (tile-blocked multi-gate passes)
  // Dense 1-/2-qubit gates are queued and flushed through b200sv_apply_gate_sequence, which packs them
  // into as few HBM passes as possible; every other method flushes first, so the observable semantics
  // are those of immediate application (cf. the reference's blocked-gate queue,
  // qubitvector_thrust.hpp:1102-1111,1511-1512).  Disable with B200SV_GATE_QUEUE=0.
  static bool queue_enabled() {
    static const bool on = [] { const char *e = getenv("B200SV_GATE_QUEUE"); return !(e && e[0] == '0'); }();
    return on;
  }
  void flush() const {
    if (q_nq_.empty()) return;
    const int ng = (int)q_nq_.size();
    const int rc = b200sv_apply_gate_sequence(h_, ng, q_nq_.data(), q_qubits_.data(), q_mats_.data(), nullptr);
    q_nq_.clear(); q_qubits_.clear(); q_mats_.clear();
    ck(rc);
  }

protected:
  bool enqueue(const reg_t &qubits, const cvector_t<double> &mat) {
    if (!queue_enabled() || qubits.size() < 1 || qubits.size() > 2 || mat.size() != (1ull << (2 * qubits.size())))
      return false;
    q_nq_.push_back((int)qubits.size());
    q_qubits_.push_back(qubits[0]);
    q_qubits_.push_back(qubits.size() == 2 ? qubits[1] : 0);
    const size_t off = q_mats_.size();
    q_mats_.resize(off + 32, 0.0);
    std::memcpy(&q_mats_[off], mat.data(), mat.size() * sizeof(std::complex<double>));
    if (q_nq_.size() >= 4096) flush();
    return true;
  }
  mutable std::vector<int> q_nq_;
  mutable std::vector<uint64_t> q_qubits_;
  mutable std::vector<double> q_mats_;

  static void ck(int rc) {
    if (rc) throw std::runtime_error(std::string("b200sv: ") + b200sv_last_error());
  }
  void sync_const() const { if (h_) { flush(); ck(b200sv_synchronize(h_)); } }
  void release() {
    q_nq_.clear(); q_qubits_.clear(); q_mats_.clear();
    if (h_) { b200sv_destroy(h_); h_ = nullptr; }
  }
  b200sv_handle h_ = nullptr;
  size_t num_qubits_ = 0;
  size_t data_size_ = 0;
  uint_t chunk_index_ = 0;
  uint_t omp_threads_ = 1, omp_threshold_ = 14;
  int sample_measure_index_size_ = 10;
  double json_chop_threshold_ = 0;
};

}  // namespace QV
}  // namespace AER
#endif
