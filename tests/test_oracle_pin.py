"""Pin the C restatement (oracle/sv_oracle.c) to the UNMODIFIED reference
QubitVector (oracle/_ref/libaer_qv_ref.so, built from /root/reference by
oracle/Makefile) on seeded random op streams, for double and single precision.

Permutation-type gates (mcx, mcy, mcswap, permutation) must be bit exact; the
floating-point gates must agree to a few ulp per op (the reference runs AVX2
FMA kernels, qv_avx2.cpp, so rounding differs in the last bits).
"""
import numpy as np
import pytest

import opgen
from oracle.oracle import OracleQV, RefQV

needs_ref = pytest.mark.skipif(not RefQV.available(), reason="oracle/_ref not built")

EXACT = {"apply_mcx", "apply_mcy", "apply_mcswap", "apply_permutation_matrix"}


@needs_ref
@pytest.mark.parametrize("n", [3, 5, 9, 13])
@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-13), (np.complex64, 2e-5)])
def test_random_stream_matches_reference(n, dtype, tol):
    kinds = None
    if dtype == np.complex64:
        # Reference bug (Aer 0.17.2): the float AVX2 diagonal kernel iterates `END = data_size >>
        # (batch+1)` blocks of `1 << (batch+2)` amplitudes (qv_avx2.cpp:1197-1203), i.e. it walks 2x
        # past the end of the state and corrupts the heap.  We follow the mathematics (as the double
        # path does) and therefore cannot pin float diagonals against the reference; every other
        # float op is pinned.  A k-qubit matrix on a k-qubit float state also overruns (n=k=3).
        if n < 5:
            pytest.skip("reference float AVX2 path overruns tiny states")
        kinds = ["matrix", "mcx", "mcy", "mcswap", "mcphase", "mcu", "pauli", "multiplexer", "permutation"]
    rng = np.random.default_rng(100 + n)
    psi0 = opgen.random_state(rng, n, dtype)
    ora, ref = OracleQV(n, dtype), RefQV(n, dtype)
    ora.set_state(psi0)
    ref.set_state(psi0)
    for op in opgen.random_ops(7 * n, n, 60, kinds=kinds):
        before = ref.vector()
        opgen.apply(ora, op)
        opgen.apply(ref, op)
        a, b = ora.vector(), ref.vector()
        if op[0] in EXACT:
            # exactness is a per-op property: replay the op on identical input
            o2 = OracleQV(n, dtype)
            o2.set_state(before)
            opgen.apply(o2, op)
            assert np.array_equal(o2.vector().view(np.uint8), b.view(np.uint8)), op[0]
        assert np.max(np.abs(a - b)) < tol * 10, op[0]
        ora.set_state(b)  # resync so per-op error does not accumulate
    assert opgen.fidelity_gap(ora.vector(), ref.vector()) < tol


@needs_ref
@pytest.mark.parametrize("n", [4, 9, 12, 14])
def test_reductions_match_reference(n):
    rng = np.random.default_rng(n)
    psi = opgen.random_state(rng, n)
    ora, ref = OracleQV(n), RefQV(n)
    ora.set_state(psi)
    ref.set_state(psi)
    assert abs(ora.norm() - ref.norm()) < 1e-13
    for k in (1, 2, 3):
        qs = opgen.pick(rng, n, k)
        K = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
        m = opgen.colmajor(K)
        assert abs(ora.norm(qs, m) - ref.norm(qs, m)) < 1e-11 * max(1.0, ref.norm(qs, m))
    for k in (1, 2, 3, min(n, 6)):
        qs = opgen.pick(rng, n, k)
        np.testing.assert_allclose(ora.probabilities(qs), ref.probabilities(qs), atol=1e-14)
    np.testing.assert_allclose(ora.probabilities(), ref.probabilities(), atol=1e-15)
    for qs, p in opgen.random_paulis(n, n, 30):
        ph = complex(np.exp(2j * np.pi * rng.random()))
        assert abs(ora.expval_pauli(qs, p, ph) - ref.expval_pauli(qs, p, ph)) < 1e-13, (qs, p)
    rnds = rng.random(500)
    assert np.array_equal(ora.sample_measure(rnds), ref.sample_measure(rnds))
    # edge draws: 0, just below 1
    edge = np.array([0.0, np.nextafter(1.0, 0.0), 0.5])
    assert np.array_equal(ora.sample_measure(edge), ref.sample_measure(edge))


@needs_ref
def test_sample_measure_small_and_indexed_paths():
    # END < 2^index_size (loop over shots) and END >= 2^index_size (block index) -- qubitvector.hpp:2159-2226
    rng = np.random.default_rng(5)
    for n, idx in ((6, 10), (12, 10), (12, 4)):
        psi = opgen.random_state(rng, n)
        ora, ref = OracleQV(n, index_size=idx), RefQV(n, index_size=idx)
        ora.set_state(psi)
        ref.set_state(psi)
        r = rng.random(300)
        assert np.array_equal(ora.sample_measure(r), ref.sample_measure(r))
