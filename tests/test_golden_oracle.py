"""The C restatement must reproduce the committed golden vectors that the
unmodified reference produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_io
import opgen
from oracle.oracle import OracleQV


@pytest.mark.parametrize("name,dtype,tol", [
    ("kernel_n6_f64.npz", np.complex128, 1e-12),
    ("kernel_n10_f64.npz", np.complex128, 1e-12),
    ("kernel_n12_f64.npz", np.complex128, 1e-12),
    ("kernel_n10_f32.npz", np.complex64, 2e-5),
])
def test_kernel_golden(name, dtype, tol):
    z, meta, ops, paulis = golden_io.load_kernel(name)
    n = meta["n"]
    qv = OracleQV(n, dtype)
    qv.set_state(z["psi0"])
    for op in ops:
        opgen.apply(qv, op)
    assert opgen.fidelity_gap(qv.vector(), z["final"]) < tol
    assert np.max(np.abs(qv.vector() - z["final"])) < tol * 10
    qv.set_state(z["final"])
    assert abs(qv.norm() - float(z["norm"])) < tol
    for (q, p), want in zip(paulis, z["expval"]):
        assert abs(qv.expval_pauli(q, p) - want) < tol
    for i, q in enumerate(meta["prob_qubits"]):
        np.testing.assert_allclose(qv.probabilities(q), z["probs%d" % i], atol=tol)
    assert np.array_equal(qv.sample_measure(z["rnds"]), z["samples"])
    assert abs(qv.norm(meta["kraus_qubits"], z["kraus"]) - float(z["kraus_norm"])) < tol * 100


@pytest.mark.parametrize("name", ["circuit_qv10.npz", "circuit_qv12.npz"])
def test_circuit_golden_statevector(name):
    z, meta, ops, paulis = golden_io.load_circuit(name)
    n = meta["n"]
    qv = OracleQV(n)
    for _, q, u in ops:
        qv.apply_matrix(q, opgen.colmajor(u))
    for tag in ("fused", "plain"):
        assert opgen.fidelity_gap(qv.vector(), z["sv_" + tag]) < 1e-12
        for (q, p), want in zip(paulis, z["ev_" + tag]):
            assert abs(qv.expval_pauli(q, p) - want) < 1e-12
