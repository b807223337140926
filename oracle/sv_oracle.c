/* TEST INFRASTRUCTURE ONLY -- "parity pinned" (see tests/test_oracle_pin.py:
 * every function here is checked against the unmodified reference compiled
 * into oracle/_ref/libaer_qv_ref.so and against tests/golden/ fixtures the
 * reference generated).
 *
 * Plain-C restatement of the reference's CPU statevector amplitude-update
 * path (Qiskit Aer 0.17.2, src/simulators/statevector/{qubitvector,indexes,
 * transformer}.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; the product (libb200sv.so) never does.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void svo_sort(const uint64_t *q, int k, uint64_t *out) {
  for (int i = 0; i < k; i++) {
    uint64_t v = q[i];
    int j = i;
    while (j > 0 && out[j - 1] > v) {
      out[j] = out[j - 1];
      j--;
    }
    out[j] = v;
  }
}

/* indexes.hpp:212-222 -- insert a zero bit at every sorted qubit position. */
static inline uint64_t svo_index0(const uint64_t *sorted, int k, uint64_t g) {
  uint64_t ret = g;
  for (int j = 0; j < k; j++) {
    const uint64_t low = ret & ((1ull << sorted[j]) - 1);
    ret >>= sorted[j];
    ret <<= sorted[j] + 1;
    ret |= low;
  }
  return ret;
}

/* indexes.hpp:237-250 -- all 2^k indices of group g; bit i of the matrix
 * index corresponds to qubits[i] (unsorted order). */
static inline void svo_indexes(const uint64_t *qubits, const uint64_t *sorted,
                               int k, uint64_t g, uint64_t *ret) {
  ret[0] = svo_index0(sorted, k, g);
  for (int i = 0; i < k; i++) {
    const uint64_t n = 1ull << i, bit = 1ull << qubits[i];
    for (uint64_t j = 0; j < n; j++)
      ret[n + j] = ret[j] | bit;
  }
}

#define REAL double
#define SFX _f64
#include "sv_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL float
#define SFX _f32
#include "sv_oracle_impl.h"
#undef REAL
#undef SFX
