// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// A C-ABI shim around the *unmodified* reference CPU statevector
// (AER::QV::QubitVector<double/float>, /root/reference/src/simulators/
// statevector/qubitvector.hpp:62-651) so that tests and bench.py's CPU
// baseline can drive the reference's own arithmetic from ctypes.
//
// This file contains no reference code: it only #includes the reference
// headers where they lie and forwards calls.  It is compiled by
// oracle/Makefile into oracle/_ref/libaer_qv_ref.so (git-ignored).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may load the result.
#include <complex>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "simulators/statevector/qubitvector.hpp"

using AER::reg_t;
using AER::uint_t;
typedef std::complex<double> cplx;

namespace {
template <typename T> struct Ref {
  AER::QV::QubitVector<T> qv;
};
inline reg_t to_reg(const uint64_t *q, int k) { return reg_t(q, q + k); }
inline std::vector<cplx> to_cvec(const double *m, size_t n) {
  std::vector<cplx> v(n);
  for (size_t i = 0; i < n; i++)
    v[i] = cplx(m[2 * i], m[2 * i + 1]);
  return v;
}
} // namespace

#define REF_API(T, SFX)                                                        \
  extern "C" void *refqv_create##SFX(int nq, int threads) {                    \
    auto *r = new Ref<T>();                                                    \
    r->qv.set_omp_threads(threads);                                            \
    r->qv.set_omp_threshold(1);                                                \
    r->qv.set_num_qubits(nq);                                                  \
    r->qv.initialize();                                                        \
    return r;                                                                  \
  }                                                                            \
  extern "C" void refqv_destroy##SFX(void *h) { delete (Ref<T> *)h; }          \
  extern "C" void *refqv_data##SFX(void *h) {                                  \
    return (void *)((Ref<T> *)h)->qv.data();                                   \
  }                                                                            \
  extern "C" void refqv_initialize##SFX(void *h) {                             \
    ((Ref<T> *)h)->qv.initialize();                                            \
  }                                                                            \
  extern "C" void refqv_set_sample_index_size##SFX(void *h, int s) {           \
    ((Ref<T> *)h)->qv.set_sample_measure_index_size(s);                        \
  }                                                                            \
  extern "C" void refqv_apply_matrix##SFX(void *h, const uint64_t *q, int k,   \
                                          const double *m) {                   \
    ((Ref<T> *)h)->qv.apply_matrix(to_reg(q, k),                               \
                                   to_cvec(m, (size_t)1 << (2 * k)));          \
  }                                                                            \
  extern "C" void refqv_apply_diagonal##SFX(void *h, const uint64_t *q, int k, \
                                            const double *d) {                 \
    ((Ref<T> *)h)->qv.apply_diagonal_matrix(to_reg(q, k),                      \
                                            to_cvec(d, (size_t)1 << k));       \
  }                                                                            \
  extern "C" void refqv_apply_mcx##SFX(void *h, const uint64_t *q, int k) {    \
    ((Ref<T> *)h)->qv.apply_mcx(to_reg(q, k));                                 \
  }                                                                            \
  extern "C" void refqv_apply_mcy##SFX(void *h, const uint64_t *q, int k) {    \
    ((Ref<T> *)h)->qv.apply_mcy(to_reg(q, k));                                 \
  }                                                                            \
  extern "C" void refqv_apply_mcswap##SFX(void *h, const uint64_t *q, int k) { \
    ((Ref<T> *)h)->qv.apply_mcswap(to_reg(q, k));                              \
  }                                                                            \
  extern "C" void refqv_apply_mcphase##SFX(void *h, const uint64_t *q, int k,  \
                                           double re, double im) {             \
    ((Ref<T> *)h)->qv.apply_mcphase(to_reg(q, k), cplx(re, im));               \
  }                                                                            \
  extern "C" void refqv_apply_mcu##SFX(void *h, const uint64_t *q, int k,      \
                                       const double *m) {                      \
    ((Ref<T> *)h)->qv.apply_mcu(to_reg(q, k), to_cvec(m, 4));                  \
  }                                                                            \
  extern "C" void refqv_apply_multiplexer##SFX(                                \
      void *h, const uint64_t *cq, int nc, const uint64_t *tq, int nt,         \
      const double *m) {                                                       \
    ((Ref<T> *)h)->qv.apply_multiplexer(                                       \
        to_reg(cq, nc), to_reg(tq, nt),                                        \
        to_cvec(m, ((size_t)1 << (nc + nt)) << nt));                           \
  }                                                                            \
  extern "C" void refqv_apply_permutation##SFX(void *h, const uint64_t *q,     \
                                               int k, const uint64_t *pairs,   \
                                               int npairs) {                   \
    std::vector<std::pair<uint_t, uint_t>> p;                                  \
    for (int i = 0; i < npairs; i++)                                           \
      p.push_back({pairs[2 * i], pairs[2 * i + 1]});                           \
    ((Ref<T> *)h)->qv.apply_permutation_matrix(to_reg(q, k), p);               \
  }                                                                            \
  extern "C" void refqv_apply_pauli##SFX(void *h, const uint64_t *q, int k,    \
                                         const char *pauli, double cre,        \
                                         double cim) {                         \
    ((Ref<T> *)h)->qv.apply_pauli(to_reg(q, k), std::string(pauli),            \
                                  cplx(cre, cim));                             \
  }                                                                            \
  extern "C" double refqv_norm##SFX(void *h) {                                 \
    return ((Ref<T> *)h)->qv.norm();                                           \
  }                                                                            \
  extern "C" double refqv_norm_matrix##SFX(void *h, const uint64_t *q, int k,  \
                                           const double *m) {                  \
    return ((Ref<T> *)h)->qv.norm(to_reg(q, k),                                \
                                  to_cvec(m, (size_t)1 << (2 * k)));           \
  }                                                                            \
  extern "C" void refqv_probabilities##SFX(void *h, const uint64_t *q, int k,  \
                                           double *out) {                      \
    auto p = ((Ref<T> *)h)->qv.probabilities(to_reg(q, k));                    \
    std::memcpy(out, p.data(), p.size() * sizeof(double));                     \
  }                                                                            \
  extern "C" void refqv_sample_measure##SFX(void *h, const double *rnds,       \
                                            int64_t shots, uint64_t *out) {    \
    std::vector<double> r(rnds, rnds + shots);                                 \
    auto s = ((Ref<T> *)h)->qv.sample_measure(r);                              \
    std::memcpy(out, s.data(), s.size() * sizeof(uint64_t));                   \
  }                                                                            \
  extern "C" double refqv_expval_pauli##SFX(void *h, const uint64_t *q, int k, \
                                            const char *pauli, double pre,     \
                                            double pim) {                      \
    return ((Ref<T> *)h)->qv.expval_pauli(to_reg(q, k), std::string(pauli),    \
                                          cplx(pre, pim));                     \
  }

REF_API(double, _f64)
REF_API(float, _f32)

extern "C" int refqv_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
