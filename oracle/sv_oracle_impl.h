/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference statevector
 * amplitude-update path, templated on the real type by textual inclusion
 * (sv_oracle.c includes this twice: REAL=double SFX=_f64, REAL=float SFX=_f32).
 *
 * Every function cites the reference (paths relative to
 * /root/reference/src/simulators/statevector/) it restates.  The state is
 * `REAL psi[2*dim]`, interleaved (re, im), amplitude index little-endian in
 * the qubit number (qubit q <-> bit q), as in the reference.
 * Matrices / diagonals / phases always arrive as complex<double> (interleaved
 * doubles), column-major vectorised (mat[i + DIM*j] = M[i][j]) and are
 * converted to REAL first, like QubitVector::convert (qubitvector.hpp:1262).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* (-i)^num_y folded into the coefficient: qubitvector.hpp:2275-2298. */
static void FN(svo_y_phase)(int num_y, REAL *re, REAL *im) {
  const REAL r = *re, i = *im;
  switch (num_y & 3) {
  case 1: *re = i;  *im = -r; break;
  case 2: *re = -r; *im = -i; break;
  case 3: *re = -i; *im = r;  break;
  default: break;
  }
}

/* ---- dense k-qubit matrix: transformer.hpp:90-164 (apply_matrix_n), the
 * generic branch: cache the 2^k inputs of a group, zero, accumulate
 * data[inds[i]] += mat[i + DIM*j] * cache[j].  qubits[0] is the least
 * significant matrix bit (indexes.hpp:224-250). */
void FN(svo_apply_matrix)(REAL *psi, int nq, const uint64_t *qubits, int k,
                          const double *mat) {
  const uint64_t DIM = 1ull << k, END = 1ull << (nq - k);
  uint64_t sorted[64];
  svo_sort(qubits, k, sorted);
  REAL *m = (REAL *)malloc(sizeof(REAL) * 2 * DIM * DIM);
  for (uint64_t i = 0; i < 2 * DIM * DIM; i++)
    m[i] = (REAL)mat[i];
#pragma omp parallel
  {
    uint64_t *inds = (uint64_t *)malloc(sizeof(uint64_t) * DIM);
    REAL *cache = (REAL *)malloc(sizeof(REAL) * 2 * DIM);
#pragma omp for
    for (int64_t g = 0; g < (int64_t)END; g++) {
      svo_indexes(qubits, sorted, k, (uint64_t)g, inds);
      for (uint64_t i = 0; i < DIM; i++) {
        cache[2 * i] = psi[2 * inds[i]];
        cache[2 * i + 1] = psi[2 * inds[i] + 1];
      }
      for (uint64_t i = 0; i < DIM; i++) {
        REAL re = 0, im = 0;
        for (uint64_t j = 0; j < DIM; j++) {
          const REAL mr = m[2 * (i + DIM * j)], mi = m[2 * (i + DIM * j) + 1];
          re += mr * cache[2 * j] - mi * cache[2 * j + 1];
          im += mr * cache[2 * j + 1] + mi * cache[2 * j];
        }
        psi[2 * inds[i]] = re;
        psi[2 * inds[i] + 1] = im;
      }
    }
    free(inds);
    free(cache);
  }
  free(m);
}

/* ---- diagonal: transformer.hpp:235-259 (general N): psi[k] *= diag[iv],
 * iv = sum_j ((k >> qubits[j]) & 1) << j; entries equal to 1 are skipped
 * (:252).  We follow the mathematics for the 1-qubit special cases
 * (transformer.hpp:261-372) -- including NOT reproducing the index slip at
 * :315-340 that the AVX2 path (qv_avx2.cpp:1103) does not have either. */
void FN(svo_apply_diagonal)(REAL *psi, int nq, const uint64_t *qubits, int k,
                            const double *diag) {
  const int64_t END = 1ll << nq;
#pragma omp parallel for
  for (int64_t idx = 0; idx < END; idx++) {
    uint64_t iv = 0;
    for (int j = 0; j < k; j++)
      iv |= (((uint64_t)idx >> qubits[j]) & 1ull) << j;
    const REAL dr = (REAL)diag[2 * iv], di = (REAL)diag[2 * iv + 1];
    if (dr == (REAL)1 && di == (REAL)0)
      continue;
    const REAL ar = psi[2 * idx], ai = psi[2 * idx + 1];
    psi[2 * idx] = dr * ar - di * ai;
    psi[2 * idx + 1] = dr * ai + di * ar;
  }
}

/* ---- mcx: qubitvector.hpp:1447-1486.  swap inds[pos0] <-> inds[pos1],
 * pos0 = MASKS[N-1] (all controls 1, target 0), pos1 = MASKS[N]. Bit exact. */
void FN(svo_apply_mcx)(REAL *psi, int nq, const uint64_t *qubits, int k) {
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64], cmask = 0;
  svo_sort(qubits, k, sorted);
  for (int j = 0; j < k - 1; j++)
    cmask |= 1ull << qubits[j];
  const uint64_t tbit = 1ull << qubits[k - 1];
#pragma omp parallel for
  for (int64_t g = 0; g < END; g++) {
    const uint64_t i0 = svo_index0(sorted, k, (uint64_t)g) | cmask;
    const uint64_t i1 = i0 | tbit;
    const REAL r = psi[2 * i0], i = psi[2 * i0 + 1];
    psi[2 * i0] = psi[2 * i1];
    psi[2 * i0 + 1] = psi[2 * i1 + 1];
    psi[2 * i1] = r;
    psi[2 * i1 + 1] = i;
  }
}

/* ---- mcy: qubitvector.hpp:1489-1537.  cache = d[pos0];
 * d[pos0] = -i * d[pos1]; d[pos1] = i * cache.  Exact (swap + sign). */
void FN(svo_apply_mcy)(REAL *psi, int nq, const uint64_t *qubits, int k) {
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64], cmask = 0;
  svo_sort(qubits, k, sorted);
  for (int j = 0; j < k - 1; j++)
    cmask |= 1ull << qubits[j];
  const uint64_t tbit = 1ull << qubits[k - 1];
#pragma omp parallel for
  for (int64_t g = 0; g < END; g++) {
    const uint64_t i0 = svo_index0(sorted, k, (uint64_t)g) | cmask;
    const uint64_t i1 = i0 | tbit;
    const REAL r0 = psi[2 * i0], m0 = psi[2 * i0 + 1];
    const REAL r1 = psi[2 * i1], m1 = psi[2 * i1 + 1];
    /* -i*(r1 + i m1) = m1 - i r1 ;  i*(r0 + i m0) = -m0 + i r0 */
    psi[2 * i0] = m1;
    psi[2 * i0 + 1] = -r1;
    psi[2 * i1] = -m0;
    psi[2 * i1 + 1] = r0;
  }
}

/* ---- mcswap: qubitvector.hpp:1540-1573.  pos0 = MASKS[N-1] (controls 1,
 * q[N-2]=1, q[N-1]=0), pos1 = pos0 + BITS[N-2] (q[N-2]=0, q[N-1]=1). */
void FN(svo_apply_mcswap)(REAL *psi, int nq, const uint64_t *qubits, int k) {
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64], cmask = 0;
  svo_sort(qubits, k, sorted);
  for (int j = 0; j < k - 2; j++)
    cmask |= 1ull << qubits[j];
  const uint64_t b0 = 1ull << qubits[k - 2], b1 = 1ull << qubits[k - 1];
#pragma omp parallel for
  for (int64_t g = 0; g < END; g++) {
    const uint64_t base = svo_index0(sorted, k, (uint64_t)g) | cmask;
    const uint64_t i0 = base | b0, i1 = base | b1;
    const REAL r = psi[2 * i0], i = psi[2 * i0 + 1];
    psi[2 * i0] = psi[2 * i1];
    psi[2 * i0 + 1] = psi[2 * i1 + 1];
    psi[2 * i1] = r;
    psi[2 * i1 + 1] = i;
  }
}

/* ---- mcphase: qubitvector.hpp:1576-1612.  d[inds[MASKS[N]]] *= phase. */
void FN(svo_apply_mcphase)(REAL *psi, int nq, const uint64_t *qubits, int k,
                           double pre, double pim) {
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64], mask = 0;
  svo_sort(qubits, k, sorted);
  for (int j = 0; j < k; j++)
    mask |= 1ull << qubits[j];
  const REAL pr = (REAL)pre, pi_ = (REAL)pim;
#pragma omp parallel for
  for (int64_t g = 0; g < END; g++) {
    const uint64_t i1 = svo_index0(sorted, k, (uint64_t)g) | mask;
    const REAL ar = psi[2 * i1], ai = psi[2 * i1 + 1];
    psi[2 * i1] = ar * pr - ai * pi_;
    psi[2 * i1 + 1] = ar * pi_ + ai * pr;
  }
}

/* ---- mcu: qubitvector.hpp:1615-1720.  Exact-== routing: off-diagonals 0 ->
 * (m[0]==1 ? mcphase(m[3]) : diagonal on pos0/pos1); N==1 -> apply_matrix;
 * else d[pos0] = m0 d0 + m2 d1 ; d[pos1] = m1 d0 + m3 d1. */
void FN(svo_apply_mcu)(REAL *psi, int nq, const uint64_t *qubits, int k,
                       const double *mat) {
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64], cmask = 0;
  svo_sort(qubits, k, sorted);
  for (int j = 0; j < k - 1; j++)
    cmask |= 1ull << qubits[j];
  const uint64_t tbit = 1ull << qubits[k - 1];
  const int diag = (mat[2] == 0.0 && mat[3] == 0.0 && mat[4] == 0.0 &&
                    mat[5] == 0.0);
  if (diag && mat[0] == 1.0 && mat[1] == 0.0) {
    FN(svo_apply_mcphase)(psi, nq, qubits, k, mat[6], mat[7]);
    return;
  }
  REAL m[8];
  for (int i = 0; i < 8; i++)
    m[i] = (REAL)mat[i];
#pragma omp parallel for
  for (int64_t g = 0; g < END; g++) {
    const uint64_t i0 = svo_index0(sorted, k, (uint64_t)g) | cmask;
    const uint64_t i1 = i0 | tbit;
    const REAL r0 = psi[2 * i0], m0 = psi[2 * i0 + 1];
    const REAL r1 = psi[2 * i1], m1 = psi[2 * i1 + 1];
    if (diag) {
      psi[2 * i0] = m[0] * r0 - m[1] * m0;
      psi[2 * i0 + 1] = m[0] * m0 + m[1] * r0;
      psi[2 * i1] = m[6] * r1 - m[7] * m1;
      psi[2 * i1 + 1] = m[6] * m1 + m[7] * r1;
    } else {
      psi[2 * i0] = (m[0] * r0 - m[1] * m0) + (m[4] * r1 - m[5] * m1);
      psi[2 * i0 + 1] = (m[0] * m0 + m[1] * r0) + (m[4] * m1 + m[5] * r1);
      psi[2 * i1] = (m[2] * r0 - m[3] * m0) + (m[6] * r1 - m[7] * m1);
      psi[2 * i1 + 1] = (m[2] * m0 + m[3] * r0) + (m[6] * m1 + m[7] * r1);
    }
  }
}

/* ---- multiplexer: qubitvector.hpp:1305-1340.  qubits = targets ++ controls;
 * DIM = 2^(nt+nc), columns = 2^nt, blocks = 2^nc;
 * d[inds[i + b*columns]] = sum_j mat[i + b*columns + DIM*j] * cache[b*columns + j]. */
void FN(svo_apply_multiplexer)(REAL *psi, int nq, const uint64_t *cq, int nc,
                               const uint64_t *tq, int nt, const double *mat) {
  const int k = nc + nt;
  uint64_t qubits[64], sorted[64];
  for (int j = 0; j < nt; j++)
    qubits[j] = tq[j];
  for (int j = 0; j < nc; j++)
    qubits[nt + j] = cq[j];
  svo_sort(qubits, k, sorted);
  const uint64_t DIM = 1ull << k, columns = 1ull << nt, blocks = 1ull << nc;
  const int64_t END = 1ll << (nq - k);
#pragma omp parallel
  {
    uint64_t *inds = (uint64_t *)malloc(sizeof(uint64_t) * DIM);
    REAL *cache = (REAL *)malloc(sizeof(REAL) * 2 * DIM);
#pragma omp for
    for (int64_t g = 0; g < END; g++) {
      svo_indexes(qubits, sorted, k, (uint64_t)g, inds);
      for (uint64_t i = 0; i < DIM; i++) {
        cache[2 * i] = psi[2 * inds[i]];
        cache[2 * i + 1] = psi[2 * inds[i] + 1];
      }
      for (uint64_t b = 0; b < blocks; b++)
        for (uint64_t i = 0; i < columns; i++) {
          REAL re = 0, im = 0;
          for (uint64_t j = 0; j < columns; j++) {
            const uint64_t mi_ = i + b * columns + DIM * j;
            const REAL mr = (REAL)mat[2 * mi_], mi = (REAL)mat[2 * mi_ + 1];
            const REAL cr = cache[2 * (b * columns + j)],
                       ci = cache[2 * (b * columns + j) + 1];
            re += mr * cr - mi * ci;
            im += mr * ci + mi * cr;
          }
          psi[2 * inds[i + b * columns]] = re;
          psi[2 * inds[i + b * columns] + 1] = im;
        }
    }
    free(inds);
    free(cache);
  }
}

/* ---- permutation: qubitvector.hpp:1354-1437.  Sequential swaps of
 * inds[p.first] <-> inds[p.second] inside each group. Bit exact. */
void FN(svo_apply_permutation)(REAL *psi, int nq, const uint64_t *qubits, int k,
                               const uint64_t *pairs, int npairs) {
  const uint64_t DIM = 1ull << k;
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64];
  svo_sort(qubits, k, sorted);
#pragma omp parallel
  {
    uint64_t *inds = (uint64_t *)malloc(sizeof(uint64_t) * DIM);
#pragma omp for
    for (int64_t g = 0; g < END; g++) {
      svo_indexes(qubits, sorted, k, (uint64_t)g, inds);
      for (int p = 0; p < npairs; p++) {
        const uint64_t a = inds[pairs[2 * p]], b = inds[pairs[2 * p + 1]];
        const REAL r = psi[2 * a], i = psi[2 * a + 1];
        psi[2 * a] = psi[2 * b];
        psi[2 * a + 1] = psi[2 * b + 1];
        psi[2 * b] = r;
        psi[2 * b + 1] = i;
      }
    }
    free(inds);
  }
}

/* ---- Pauli string apply: qubitvector.hpp:2393-2437 (masks :2236-2298).
 * x_mask/z_mask/num_y from pauli[N-1-i] <-> qubits[i]; phase = coeff*(-i)^num_y;
 * pairs (i0, i0^x_mask) swapped then each multiplied by (-1)^popc(idx&z)*phase. */
void FN(svo_apply_pauli)(REAL *psi, int nq, uint64_t x_mask, uint64_t z_mask,
                         int num_y, int x_max, double cre, double cim) {
  if ((x_mask | z_mask) == 0)
    return;
  REAL pr = (REAL)cre, pi_ = (REAL)cim;
  FN(svo_y_phase)(num_y, &pr, &pi_);
  if (!x_mask) {
    const int64_t END = 1ll << nq;
#pragma omp parallel for
    for (int64_t i = 0; i < END; i++) {
      REAL ar = psi[2 * i], ai = psi[2 * i + 1];
      if (__builtin_popcountll((uint64_t)i & z_mask) & 1) {
        ar = -ar;
        ai = -ai;
      }
      psi[2 * i] = ar * pr - ai * pi_;
      psi[2 * i + 1] = ar * pi_ + ai * pr;
    }
    return;
  }
  const uint64_t mask_u = ~((1ull << (x_max + 1)) - 1);
  const uint64_t mask_l = (1ull << x_max) - 1;
  const int64_t END = 1ll << (nq - 1);
#pragma omp parallel for
  for (int64_t i = 0; i < END; i++) {
    uint64_t idx[2];
    idx[0] = (((uint64_t)i << 1) & mask_u) | ((uint64_t)i & mask_l);
    idx[1] = idx[0] ^ x_mask;
    REAL v[2][2] = {{psi[2 * idx[1]], psi[2 * idx[1] + 1]},
                    {psi[2 * idx[0]], psi[2 * idx[0] + 1]}};
    for (int j = 0; j < 2; j++) {
      REAL ar = v[j][0], ai = v[j][1];
      if (__builtin_popcountll(idx[j] & z_mask) & 1) {
        ar = -ar;
        ai = -ai;
      }
      psi[2 * idx[j]] = ar * pr - ai * pi_;
      psi[2 * idx[j] + 1] = ar * pi_ + ai * pr;
    }
  }
}

/* ---- norm: qubitvector.hpp:1879-1886.  sum |psi|^2, FP64 accumulate. */
double FN(svo_norm)(const REAL *psi, int nq) {
  const int64_t END = 1ll << nq;
  double s = 0;
#pragma omp parallel for reduction(+ : s)
  for (int64_t i = 0; i < END; i++)
    s += (double)(psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1]);
  return s;
}

/* ---- norm(qubits, mat): qubitvector.hpp:1889-1963.  sum_groups sum_i
 * | sum_j mat[i + DIM*j] psi[inds[j]] |^2  (Kraus probability). */
double FN(svo_norm_matrix)(const REAL *psi, int nq, const uint64_t *qubits,
                           int k, const double *mat) {
  const uint64_t DIM = 1ull << k;
  const int64_t END = 1ll << (nq - k);
  uint64_t sorted[64];
  svo_sort(qubits, k, sorted);
  double s = 0;
#pragma omp parallel reduction(+ : s)
  {
    uint64_t *inds = (uint64_t *)malloc(sizeof(uint64_t) * DIM);
#pragma omp for
    for (int64_t g = 0; g < END; g++) {
      svo_indexes(qubits, sorted, k, (uint64_t)g, inds);
      for (uint64_t i = 0; i < DIM; i++) {
        REAL re = 0, im = 0;
        for (uint64_t j = 0; j < DIM; j++) {
          const REAL mr = (REAL)mat[2 * (i + DIM * j)],
                     mi = (REAL)mat[2 * (i + DIM * j) + 1];
          const REAL ar = psi[2 * inds[j]], ai = psi[2 * inds[j] + 1];
          re += mr * ar - mi * ai;
          im += mr * ai + mi * ar;
        }
        s += (double)(re * re + im * im);
      }
    }
    free(inds);
  }
  return s;
}

/* ---- probabilities(qubits): qubitvector.hpp:2108-2143.  probs[m] =
 * sum over groups of |psi[inds[m]]|^2 ; bit j of m <-> qubits[j]. */
void FN(svo_probabilities)(const REAL *psi, int nq, const uint64_t *qubits,
                           int k, double *out) {
  const uint64_t DIM = 1ull << k;
  const int64_t END = 1ll << nq;
  for (uint64_t m = 0; m < DIM; m++)
    out[m] = 0;
  for (int64_t idx = 0; idx < END; idx++) {
    uint64_t m = 0;
    for (int j = 0; j < k; j++)
      m |= (((uint64_t)idx >> qubits[j]) & 1ull) << j;
    out[m] += (double)(psi[2 * idx] * psi[2 * idx] +
                       psi[2 * idx + 1] * psi[2 * idx + 1]);
  }
}

/* ---- sample_measure: qubitvector.hpp:2149-2228.  Strict `rnd < p`
 * cumulative search; when 2^nq >= 2^index_size the state is first summed in
 * 2^index_size contiguous blocks of `loop` amplitudes (:2188-2203), the block
 * is located with `rnd < p + idxs[j]` (:2212-2218), then the scan continues
 * amplitude by amplitude up to END-1 (:2220-2225). */
void FN(svo_sample_measure)(const REAL *psi, int nq, int index_size,
                            const double *rnds, int64_t shots, uint64_t *out) {
  const int64_t END = 1ll << nq;
  const int64_t INDEX_END = 1ll << index_size;
#define PROB(i)                                                                \
  ((double)(psi[2 * (i)] * psi[2 * (i)] + psi[2 * (i) + 1] * psi[2 * (i) + 1]))
  if (END < INDEX_END) {
    for (int64_t s = 0; s < shots; s++) {
      const double rnd = rnds[s];
      double p = 0;
      int64_t sample;
      for (sample = 0; sample < END - 1; ++sample) {
        p += PROB(sample);
        if (rnd < p)
          break;
      }
      out[s] = (uint64_t)sample;
    }
    return;
  }
  double *idxs = (double *)calloc((size_t)INDEX_END, sizeof(double));
  const uint64_t loop = (uint64_t)(END >> index_size);
#pragma omp parallel for
  for (int64_t i = 0; i < INDEX_END; i++) {
    const uint64_t base = loop * (uint64_t)i;
    double total = 0;
    for (uint64_t j = 0; j < loop; j++)
      total += PROB(base | j);
    idxs[i] = total;
  }
#pragma omp parallel for
  for (int64_t s = 0; s < shots; s++) {
    const double rnd = rnds[s];
    double p = 0;
    int64_t sample = 0;
    for (int64_t j = 0; j < INDEX_END; j++) {
      if (rnd < (p + idxs[j]))
        break;
      p += idxs[j];
      sample += (int64_t)loop;
    }
    for (; sample < END - 1; ++sample) {
      p += PROB(sample);
      if (rnd < p)
        break;
    }
    out[s] = (uint64_t)sample;
  }
  free(idxs);
#undef PROB
}

/* ---- expval_pauli: qubitvector.hpp:2300-2348.  phase = initial*(-i)^num_y;
 * Z-only: sum Re(phase |psi_i|^2) * (-1)^popc(i&z); else over pairs
 * (i0, i0^x): Re(phase psi[i1] conj(psi[i0])) and its mirror, each signed by
 * the parity of idx&z.  All-identity returns norm() (:2309-2311). */
double FN(svo_expval_pauli)(const REAL *psi, int nq, uint64_t x_mask,
                            uint64_t z_mask, int num_y, int x_max, double pre,
                            double pim) {
  if ((x_mask | z_mask) == 0)
    return FN(svo_norm)(psi, nq);
  REAL pr = (REAL)pre, pi_ = (REAL)pim;
  FN(svo_y_phase)(num_y, &pr, &pi_);
  double s = 0;
  if (!x_mask) {
    const int64_t END = 1ll << nq;
#pragma omp parallel for reduction(+ : s)
    for (int64_t i = 0; i < END; i++) {
      /* Re(phase * d * conj(d)) evaluated left to right like the reference */
      const REAL ar = psi[2 * i], ai = psi[2 * i + 1];
      const REAL tr = pr * ar - pi_ * ai, ti = pr * ai + pi_ * ar;
      double v = (double)(tr * ar + ti * ai);
      if (__builtin_popcountll((uint64_t)i & z_mask) & 1)
        v = -v;
      s += v;
    }
    return s;
  }
  const uint64_t mask_u = ~((1ull << (x_max + 1)) - 1);
  const uint64_t mask_l = (1ull << x_max) - 1;
  const int64_t END = 1ll << (nq - 1);
#pragma omp parallel for reduction(+ : s)
  for (int64_t i = 0; i < END; i++) {
    uint64_t idx[2];
    idx[0] = (((uint64_t)i << 1) & mask_u) | ((uint64_t)i & mask_l);
    idx[1] = idx[0] ^ x_mask;
    const REAL a0r = psi[2 * idx[0]], a0i = psi[2 * idx[0] + 1];
    const REAL a1r = psi[2 * idx[1]], a1i = psi[2 * idx[1] + 1];
    /* vals[0] = Re(phase * d1 * conj(d0)); vals[1] = Re(phase * d0 * conj(d1)) */
    REAL tr = pr * a1r - pi_ * a1i, ti = pr * a1i + pi_ * a1r;
    double v0 = (double)(tr * a0r + ti * a0i);
    tr = pr * a0r - pi_ * a0i;
    ti = pr * a0i + pi_ * a0r;
    double v1 = (double)(tr * a1r + ti * a1i);
    if (__builtin_popcountll(idx[0] & z_mask) & 1)
      v0 = -v0;
    if (__builtin_popcountll(idx[1] & z_mask) & 1)
      v1 = -v1;
    s += v0 + v1;
  }
  return s;
}

#undef FN
#undef CAT
#undef CAT_
