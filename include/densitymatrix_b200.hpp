// DensityMatrixB200<data_t>: density-matrix state on the B200 engine.
//
// Mirrors AER::QV::DensityMatrix<data_t> (src/simulators/density_matrix/densitymatrix.hpp:35-170; its GPU
// twin densitymatrix_thrust.hpp:302-1306) for `DensityMatrix::State<densmat_t>`
// (src/simulators/density_matrix/densitymatrix_state.hpp): rho is stored as the 2n-qubit vector vec(rho)
// (index = row + col * 2^n), every channel is a vector operation on the doubled register -- U (x) conj(U) as a
// superoperator matrix or as two passes (densitymatrix.hpp:292-330), X / CX / SWAP / Toffoli as permutations
// (:338-450), phases as diagonals -- forwarded to QubitVectorB200 and from there to the C ABI.  Pauli expectation
// values, the trace and marginal probabilities are device reductions over one 2^n-entry line of the matrix
// (b200sv_dm_expval_pauli / b200sv_dm_probabilities); sampling draws from the diagonal (b200sv_download_line) with a
// prefix sum + binary search on the host.
#ifndef _qv_density_matrix_b200_hpp_
#define _qv_density_matrix_b200_hpp_

#include "framework/linalg/matrix_utils.hpp"
#include "framework/matrix.hpp"
#include "qubitvector_b200.hpp"

namespace AER {
namespace QV {

// square matrix stored as a 2n-qubit vector (unitary/unitarymatrix.hpp:38-125)
template <typename data_t = double> class UnitaryMatrixB200 : public QubitVectorB200<data_t> {
public:
  using BaseVector = QubitVectorB200<data_t>;
  UnitaryMatrixB200() = default;
  explicit UnitaryMatrixB200(size_t num_qubits) { set_num_qubits(num_qubits); }
  UnitaryMatrixB200(const UnitaryMatrixB200 &) {}
  UnitaryMatrixB200 &operator=(const UnitaryMatrixB200 &) = delete;
  UnitaryMatrixB200 &operator=(UnitaryMatrixB200 &&o) {
    num_qubits_ = o.num_qubits_; rows_ = o.rows_;
    BaseVector::operator=(std::move(o));
    return *this;
  }
  static std::string name() { return "unitary_b200"; }
  void set_num_qubits(size_t num_qubits) override {
    num_qubits_ = num_qubits;
    rows_ = 1ULL << num_qubits;
    BaseVector::set_num_qubits(2 * num_qubits);
  }
  size_t num_rows() const { return rows_; }
  virtual uint_t num_qubits() const override { return num_qubits_; }
  matrix<std::complex<data_t>> copy_to_matrix() const {
    matrix<std::complex<data_t>> ret(rows_, rows_);
    auto v = BaseVector::vector();
    for (size_t i = 0; i < v.size(); i++) ret[i] = v[i];
    return ret;
  }
  matrix<std::complex<data_t>> move_to_matrix() { return copy_to_matrix(); }
  std::complex<double> trace() const {
    double t = 0;
    BaseVector::flush();
    b200detail::ck(b200sv_dm_expval_pauli(this->Hs(), (int)num_qubits_, nullptr, 0, "", 1.0, 0.0, &t));
    return t;
  }
  template <typename T> void initialize_from_matrix(const matrix<std::complex<T>> &mat) {
    if (mat.GetRows() != rows_ || mat.GetColumns() != rows_) throw std::runtime_error("UnitaryMatrix::initialize input matrix is incorrect shape");
    std::vector<std::complex<data_t>> v(rows_ * rows_);
    for (size_t i = 0; i < v.size(); i++) v[i] = std::complex<data_t>(mat[i]);
    BaseVector::initialize_from_data(v.data(), v.size());
  }
  void initialize_from_matrix(matrix<std::complex<data_t>> &&mat) { initialize_from_matrix<data_t>(mat); }

protected:
  std::vector<std::complex<data_t>> line(uint64_t xor_mask) const {
    std::vector<std::complex<data_t>> out(rows_);
    BaseVector::flush();
    b200detail::ck(b200sv_download_line(this->Hs(), (int)num_qubits_, xor_mask, out.data()));
    return out;
  }
  size_t num_qubits_ = 0;
  size_t rows_ = 1;
};

template <typename data_t = double> class DensityMatrixB200 : public UnitaryMatrixB200<data_t> {
public:
  using BaseVector = QubitVectorB200<data_t>;
  using BaseMatrix = UnitaryMatrixB200<data_t>;
  DensityMatrixB200() = default;
  explicit DensityMatrixB200(size_t num_qubits) : BaseMatrix(num_qubits) {}
  DensityMatrixB200(const DensityMatrixB200 &) {}
  DensityMatrixB200 &operator=(const DensityMatrixB200 &) = delete;
  DensityMatrixB200 &operator=(DensityMatrixB200 &&o) noexcept {
    BaseMatrix::operator=(std::move(o));
    return *this;
  }
  static std::string name() { return "density_matrix_b200"; }
  virtual bool batched_optimization_supported(void) override { return false; }

  void initialize() { BaseVector::initialize(); }  // |0><0|: vec index 0 (densitymatrix.hpp:203-209)
  void initialize(const DensityMatrixB200<data_t> &obj) { BaseVector::initialize(obj); }
  template <typename list_t> void initialize_from_vector(const list_t &vec) {  // densitymatrix.hpp:211-245
    if (this->size() == vec.size()) {
      BaseVector::initialize_from_vector(vec);
    } else if (this->size() == vec.size() * vec.size()) {
      BaseVector::initialize_from_vector(AER::Utils::tensor_product(AER::Utils::conjugate(vec), vec));
    } else {
      throw std::runtime_error("DensityMatrix::initialize input vector is incorrect length. Expected: " +
                               std::to_string(this->size()) + " Received: " + std::to_string(vec.size()));
    }
  }
  void transpose() {  // rho[i, j] <-> rho[j, i]: swap the row and column registers (densitymatrix.hpp:247-263)
    const uint_t nq = this->num_qubits();
    for (uint_t q = 0; q < nq; q++) BaseVector::apply_mcswap({q, q + nq});
  }
  virtual uint_t num_qubits() const override { return BaseMatrix::num_qubits_; }
  virtual reg_t superop_qubits(const reg_t &qubits) const {
    reg_t sq = qubits;
    const auto nq = num_qubits();
    for (const auto &q : qubits) sq.push_back(q + nq);
    return sq;
  }

  //---------------------------------------------------------------- channels (densitymatrix.hpp:265-332)
  void apply_superop_matrix(const reg_t &qubits, const cvector_t<double> &mat) { BaseVector::apply_matrix(superop_qubits(qubits), mat); }
  void apply_diagonal_superop_matrix(const reg_t &qubits, const cvector_t<double> &diag) { BaseVector::apply_diagonal_matrix(superop_qubits(qubits), diag); }
  void apply_unitary_matrix(const reg_t &qubits, const cvector_t<double> &mat) {
    // one 2k-qubit superoperator pass while it stays HBM bound on B200 (2k <= 4), else U then conj(U)
    if (qubits.size() > apply_unitary_threshold_) {
      const auto nq = num_qubits();
      reg_t conj_qubits;
      for (const auto &q : qubits) conj_qubits.push_back(q + nq);
      BaseVector::apply_matrix(qubits, mat);
      BaseVector::apply_matrix(conj_qubits, AER::Utils::conjugate(mat));
    } else {
      apply_superop_matrix(qubits, vmat2vsuperop(mat));
    }
  }
  void apply_diagonal_unitary_matrix(const reg_t &qubits, const cvector_t<double> &diag) {
    apply_diagonal_superop_matrix(qubits, AER::Utils::tensor_product(AER::Utils::conjugate(diag), diag));
  }

  //---------------------------------------------------------------- specialised gates (densitymatrix.hpp:338-450)
  void apply_cnot(const uint_t qctrl, const uint_t qtrgt) {
    const size_t nq = num_qubits();
    BaseVector::apply_mcx({qctrl, qtrgt});            // CX (x) CX: two exact permutation passes
    BaseVector::apply_mcx({qctrl + nq, qtrgt + nq});
  }
  void apply_cy(const uint_t qctrl, const uint_t qtrgt) { apply_unitary_matrix({qctrl, qtrgt}, Linalg::VMatrix::CY); }
  void apply_phase(const uint_t q, const complex_t &phase) {
    const auto nq = num_qubits();
    BaseVector::apply_diagonal_matrix({q, q + nq}, {1.0, phase, std::conj(phase), 1.0});
  }
  void apply_cphase(const uint_t q0, const uint_t q1, const complex_t &phase) {
    const auto nq = num_qubits();
    BaseVector::apply_mcphase({q0, q1}, phase);
    BaseVector::apply_mcphase({q0 + nq, q1 + nq}, std::conj(phase));
  }
  void apply_swap(const uint_t q0, const uint_t q1) {
    const size_t nq = num_qubits();
    BaseVector::apply_mcswap({q0, q1});
    BaseVector::apply_mcswap({q0 + nq, q1 + nq});
  }
  void apply_ecr(const uint_t q0, const uint_t q1) { apply_unitary_matrix({q0, q1}, Linalg::VMatrix::ECR); }
  void apply_x(const uint_t qubit) {
    BaseVector::apply_mcx({qubit});
    BaseVector::apply_mcx({qubit + num_qubits()});
  }
  void apply_y(const uint_t qubit) {
    // Y (x) conj(Y): swap 00<->11 and 01<->10 with a sign on the latter pair (densitymatrix.hpp:409-424)
    BaseVector::apply_mcy({qubit});
    BaseVector::apply_mcy({qubit + num_qubits()});
    BaseVector::apply_diagonal_matrix({qubit, qubit + num_qubits()}, {-1.0, -1.0, -1.0, -1.0});
  }
  void apply_toffoli(const uint_t qctrl0, const uint_t qctrl1, const uint_t qtrgt) {
    const size_t nq = num_qubits();
    BaseVector::apply_mcx({qctrl0, qctrl1, qtrgt});
    BaseVector::apply_mcx({qctrl0 + nq, qctrl1 + nq, qtrgt + nq});
  }
  void apply_reset(const reg_t &qubits) {  // densitymatrix.hpp:595-602
    const auto reset_op = Linalg::SMatrix::reset(1ULL << qubits.size());
    apply_superop_matrix(qubits, Utils::vectorize_matrix(reset_op));
  }

  //---------------------------------------------------------------- measurement statistics from the diagonal
  virtual double probability(const uint_t outcome) const override {
    return std::real(std::complex<double>(BaseVector::get_state(outcome * (BaseMatrix::rows_ + 1))));
  }
  virtual std::vector<double> probabilities() const override {
    auto d = BaseMatrix::line(0);
    std::vector<double> p(d.size());
    for (size_t i = 0; i < d.size(); i++) p[i] = std::real(d[i]);
    return p;
  }
  virtual std::vector<double> probabilities(const reg_t &qubits) const override {  // qubitvector.hpp:2108-2143 on the diagonal
    if (qubits.size() <= 12) {  // one strided pass over the diagonal on the device
      std::vector<double> p(1ull << qubits.size(), 0.0);
      BaseVector::flush();
      b200detail::ck(b200sv_dm_probabilities(this->Hs(), (int)num_qubits(), qubits.data(), (int)qubits.size(), p.data()));
      return p;
    }
    const auto diag = probabilities();
    std::vector<double> p(1ull << qubits.size(), 0.0);
    for (size_t i = 0; i < diag.size(); i++) {
      size_t m = 0;
      for (size_t j = 0; j < qubits.size(); j++) m |= ((i >> qubits[j]) & 1ull) << j;
      p[m] += diag[i];
    }
    return p;
  }
  virtual reg_t sample_measure(const std::vector<double> &rnds) const override {  // qubitvector.hpp:2149-2228
    const auto diag = probabilities();
    const size_t END = diag.size();
    std::vector<double> cum(END);  // inclusive prefix sums in index order, as the reference's running sum
    double p = 0;
    for (size_t i = 0; i < END; i++) { p += diag[i]; cum[i] = p; }
    reg_t samples(rnds.size(), 0);
    for (size_t s = 0; s < rnds.size(); s++) {
      // first index with rnd < cumulative probability, at most END - 1
      const size_t sample = std::upper_bound(cum.begin(), cum.end() - 1, rnds[s]) - cum.begin();
      samples[s] = sample;
    }
    return samples;
  }
  double expval_pauli(const reg_t &qubits, const std::string &pauli, const complex_t initial_phase = 1.0) const {
    uint_t x_mask, z_mask, num_y, x_max;
    std::tie(x_mask, z_mask, num_y, x_max) = pauli_masks_and_phase(qubits, pauli);
    if (x_mask + z_mask == 0) return std::real(BaseMatrix::trace());  // densitymatrix.hpp:463-466
    // sum_i Re(phase * rho[i ^ x, i]) * (-1)^popcount(i & z)   (densitymatrix.hpp:470-520): device reduction
    double val = 0;
    BaseVector::flush();
    b200detail::ck(b200sv_dm_expval_pauli(this->Hs(), (int)num_qubits(), qubits.data(), (int)qubits.size(), pauli.c_str(),
                                          initial_phase.real(), initial_phase.imag(), &val));
    return val;
  }
  double expval_pauli_non_diagonal_chunk(const reg_t &qubits, const std::string &pauli, const complex_t initial_phase = 1.0) const {
    return expval_pauli(qubits, pauli, initial_phase);
  }

protected:
  cvector_t<double> vmat2vsuperop(const cvector_t<double> &vmat) const {  // densitymatrix.hpp:276-288
    size_t dim = size_t(std::sqrt(vmat.size()));
    cvector_t<double> ret(dim * dim * dim * dim, 0.);
    for (size_t i = 0; i < dim; i++)
      for (size_t j = 0; j < dim; j++)
        for (size_t k = 0; k < dim; k++)
          for (size_t l = 0; l < dim; l++)
            ret[dim * i + k + (dim * dim) * (dim * j + l)] = std::conj(vmat[i + dim * j]) * vmat[k + dim * l];
    return ret;
  }
  size_t apply_unitary_threshold_ = 2;
};

}  // namespace QV
}  // namespace AER
#endif
