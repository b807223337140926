"""CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.

Tolerances (north star): |1 - <psi_ref|psi>| <= 1e-10 double, 1e-5 single;
expectation values likewise; permutation gates bit exact; sampled indices equal.
"""
import numpy as np
import pytest

import golden_io
import opgen
from oracle.oracle import OracleQV

pytestmark = pytest.mark.gpu

EXACT = {"apply_mcx", "apply_mcy", "apply_mcswap", "apply_permutation_matrix"}
TOL = {np.dtype(np.complex128): 1e-10, np.dtype(np.complex64): 1e-5}
AMP_TOL = {np.dtype(np.complex128): 1e-12, np.dtype(np.complex64): 5e-6}


def gpu_qv(n, dtype=np.complex128, **kw):
    import qiskit_aer_b200 as q
    return q.QubitVectorB200(n, dtype, **kw)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11, 14, 17])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_random_op_stream(n, dtype):
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(1000 + n)
    psi0 = opgen.random_state(rng, n, dtype)
    ora, gpu = OracleQV(n, dtype), gpu_qv(n, dtype)
    ora.set_state(psi0)
    gpu.set_state(psi0)
    kinds = None
    if n < 3:
        kinds = ["matrix", "diagonal", "mcx", "mcy", "mcphase", "mcu", "mcu_diag", "pauli"]
    for op in opgen.random_ops(31 * n, n, 50, kinds=kinds, max_k=min(5, n)):
        opgen.apply(ora, op)
        opgen.apply(gpu, op)
        a, b = ora.vector(), gpu.vector()
        if op[0] in EXACT:
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), op[0]
        assert np.max(np.abs(a - b)) < AMP_TOL[dtype], (op[0], op[1][0])
        gpu.set_state(a)  # resync: per-op comparison
    assert opgen.fidelity_gap(ora.vector(), gpu.vector()) < TOL[dtype]


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_unsynced_stream_fidelity(dtype):
    """No resync: 200 ops back to back, final-state fidelity at the north-star bar."""
    dtype = np.dtype(dtype)
    n = 14
    rng = np.random.default_rng(7)
    psi0 = opgen.random_state(rng, n, dtype)
    ora, gpu = OracleQV(n, dtype), gpu_qv(n, dtype)
    ora.set_state(psi0)
    gpu.set_state(psi0)
    for op in opgen.random_ops(99, n, 200):
        opgen.apply(ora, op)
        opgen.apply(gpu, op)
    assert opgen.fidelity_gap(ora.vector(), gpu.vector()) < TOL[dtype]


def _placements(n, k):
    """qubit placement classes: all-low, all-high, straddling warp/block boundaries, reversed order."""
    out = [list(range(k)), list(range(n - k, n)), list(range(k))[::-1]]
    mid = [0, 4, 5, 9, n - 1, 2, 7][:k]
    out.append(mid)
    out.append(sorted(mid)[::-1])
    out.append([3, 8, 1, n - 2, 6][:k])
    return out


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_dense_matrix_qubit_placements(k):
    n = 16
    rng = np.random.default_rng(k)
    psi0 = opgen.random_state(rng, n)
    for qs in _placements(n, k):
        U = opgen.colmajor(opgen.haar_unitary(rng, 1 << k))
        ora, gpu = OracleQV(n), gpu_qv(n)
        ora.set_state(psi0)
        gpu.set_state(psi0)
        ora.apply_matrix(qs, U)
        gpu.apply_matrix(qs, U)
        assert np.max(np.abs(ora.vector() - gpu.vector())) < 1e-13, qs


@pytest.mark.parametrize("k", [6, 7, 8])
def test_dense_matrix_generic_large_k(k):
    n = 12
    rng = np.random.default_rng(60 + k)
    psi0 = opgen.random_state(rng, n)
    qs = opgen.pick(rng, n, k)
    U = opgen.colmajor(opgen.haar_unitary(rng, 1 << k))
    ora, gpu = OracleQV(n), gpu_qv(n)
    ora.set_state(psi0)
    gpu.set_state(psi0)
    ora.apply_matrix(qs, U)
    gpu.apply_matrix(qs, U)
    assert np.max(np.abs(ora.vector() - gpu.vector())) < 1e-12


@pytest.mark.parametrize("k", [1, 3, 5, 7, 10, 12, 13])
def test_diagonal_sizes(k):
    n = 13
    rng = np.random.default_rng(70 + k)
    psi0 = opgen.random_state(rng, n)
    for qs in (opgen.pick(rng, n, k), list(range(k)), list(range(n - k, n))):
        d = np.exp(2j * np.pi * rng.random(1 << k))
        ora, gpu = OracleQV(n), gpu_qv(n)
        ora.set_state(psi0)
        gpu.set_state(psi0)
        ora.apply_diagonal_matrix(qs, d)
        gpu.apply_diagonal_matrix(qs, d)
        assert np.max(np.abs(ora.vector() - gpu.vector())) < 1e-14


def test_multi_controlled_gates_many_controls():
    n = 12
    rng = np.random.default_rng(5)
    psi0 = opgen.random_state(rng, n)
    for nc in (0, 1, 3, 6, n - 1):
        qs = opgen.pick(rng, n, nc + 1)
        for name, args in (("apply_mcx", ()), ("apply_mcy", ()), ("apply_mcphase", (np.exp(0.3j),)),
                           ("apply_mcu", (opgen.colmajor(opgen.haar_unitary(rng, 2)),))):
            ora, gpu = OracleQV(n), gpu_qv(n)
            ora.set_state(psi0)
            gpu.set_state(psi0)
            getattr(ora, name)(qs, *args)
            getattr(gpu, name)(qs, *args)
            a, b = ora.vector(), gpu.vector()
            if name in EXACT:
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
            assert np.max(np.abs(a - b)) < 1e-14, (name, qs)
        if nc + 2 <= n:
            qs = opgen.pick(rng, n, nc + 2)
            ora, gpu = OracleQV(n), gpu_qv(n)
            ora.set_state(psi0)
            gpu.set_state(psi0)
            ora.apply_mcswap(qs)
            gpu.apply_mcswap(qs)
            assert np.array_equal(ora.vector().view(np.uint8), gpu.vector().view(np.uint8))


@pytest.mark.parametrize("name,dtype", [
    ("kernel_n6_f64.npz", np.complex128), ("kernel_n10_f64.npz", np.complex128),
    ("kernel_n12_f64.npz", np.complex128), ("kernel_n10_f32.npz", np.complex64)])
def test_golden_kernel_fixture(name, dtype):
    """Committed vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    dtype = np.dtype(dtype)
    z, meta, ops, paulis = golden_io.load_kernel(name)
    n = meta["n"]
    gpu = gpu_qv(n, dtype)
    gpu.set_state(z["psi0"])
    for op in ops:
        opgen.apply(gpu, op)
    assert opgen.fidelity_gap(gpu.vector(), z["final"]) < TOL[dtype]
    gpu.set_state(z["final"])
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    assert abs(gpu.norm() - float(z["norm"])) < tol
    for (q, p), want in zip(paulis, z["expval"]):
        assert abs(gpu.expval_pauli(q, p) - want) < tol
    for i, q in enumerate(meta["prob_qubits"]):
        np.testing.assert_allclose(gpu.probabilities(q), z["probs%d" % i], atol=tol)
    assert np.array_equal(gpu.sample_measure(z["rnds"]), z["samples"])
    assert abs(gpu.norm(meta["kraus_qubits"], z["kraus"]) - float(z["kraus_norm"])) < tol * 100


@pytest.mark.parametrize("name", ["circuit_qv10.npz", "circuit_qv12.npz"])
def test_golden_circuit_counts_and_expvals(name):
    """QV circuit: statevector, Pauli expvals and fixed-seed counts equal the reference Controller's."""
    import qiskit_aer_b200 as q
    z, meta, ops, paulis = golden_io.load_circuit(name)
    n = meta["n"]
    gpu = gpu_qv(n)
    for _, qs, u in ops:
        gpu.apply_matrix(qs, opgen.colmajor(u))
    assert opgen.fidelity_gap(gpu.vector(), z["sv_plain"]) < 1e-10
    assert opgen.fidelity_gap(gpu.vector(), z["sv_fused"]) < 1e-10
    for (qs, p), want in zip(paulis, z["ev_plain"]):
        assert abs(gpu.expval_pauli(qs, p) - want) < 1e-10
    samples = gpu.sample_measure(q.rng_uniform(meta["seed"], meta["shots"]))
    counts = np.bincount(samples.astype(np.int64), minlength=1 << n)
    assert np.array_equal(counts, z["counts_plain"])


@pytest.mark.parametrize("n", [1, 4, 9, 13, 18])
def test_reductions(n):
    rng = np.random.default_rng(n)
    psi = opgen.random_state(rng, n)
    ora, gpu = OracleQV(n), gpu_qv(n)
    ora.set_state(psi)
    gpu.set_state(psi)
    assert abs(ora.norm() - gpu.norm()) < 1e-12
    for k in range(1, min(n, 5) + 1):
        qs = opgen.pick(rng, n, k)
        K = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
        m = opgen.colmajor(K)
        want = ora.norm(qs, m)
        assert abs(want - gpu.norm(qs, m)) < 1e-11 * max(1.0, want)
    for k in sorted({1, min(2, n), min(3, n), min(n, 6), min(n, 11), n if n <= 14 else 12}):
        qs = opgen.pick(rng, n, k)
        np.testing.assert_allclose(gpu.probabilities(qs), ora.probabilities(qs), atol=1e-13)
    for qs, p in opgen.random_paulis(n, n, 30):
        ph = complex(np.exp(2j * np.pi * rng.random()))
        assert abs(ora.expval_pauli(qs, p, ph) - gpu.expval_pauli(qs, p, ph)) < 1e-12, (qs, p)
    rnds = np.concatenate([rng.random(2000), [0.0, 1.0 - 1e-9]])
    assert np.array_equal(ora.sample_measure(rnds), gpu.sample_measure(rnds))
    # rnd >= the (rounded) total probability: the reference's block-index path walks off the end and
    # returns END (qubitvector.hpp:2212-2225 -- `sample += loop` for every block, then the `< END-1`
    # loop never runs); we clamp to the last valid index instead.
    last = gpu.sample_measure(np.array([np.nextafter(1.0, 0.0)]))[0]
    assert last <= (1 << n) - 1


def test_sampler_is_non_destructive_and_handles_concentrated_states():
    n = 15
    gpu, ora = gpu_qv(n), OracleQV(n)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[12345] = np.sqrt(0.25)
    psi[(1 << n) - 1] = np.sqrt(0.75)
    gpu.set_state(psi)
    ora.set_state(psi)
    r = np.random.default_rng(3).random(500)
    assert np.array_equal(gpu.sample_measure(r), ora.sample_measure(r))
    assert np.array_equal(gpu.vector(), psi)


def test_checkpoint_revert_inner_product():
    n = 10
    rng = np.random.default_rng(2)
    a, b = opgen.random_state(rng, n), opgen.random_state(rng, n)
    gpu = gpu_qv(n)
    gpu.set_state(a)
    gpu.checkpoint()
    gpu.set_state(b)
    z = gpu.inner_product()  # sum data * conj(checkpoint)  (qubitvector.hpp:1030-1041)
    assert abs(z - np.sum(b * np.conj(a))) < 1e-12
    gpu.revert(True)
    assert np.array_equal(gpu.vector(), a)


def test_initialize_and_component():
    n = 7
    gpu = gpu_qv(n)
    v = gpu.vector()
    assert v[0] == 1 and np.count_nonzero(v) == 1
    rng = np.random.default_rng(0)
    comp = opgen.random_state(rng, 2)
    gpu.initialize_component([1, 4], comp)  # qubitvector.hpp:879-900
    want = np.zeros(1 << n, dtype=np.complex128)
    for e in range(4):
        idx = ((e & 1) << 1) | (((e >> 1) & 1) << 4)
        want[idx] = comp[e]
    assert np.max(np.abs(gpu.vector() - want)) < 1e-15
    gpu.zero()
    assert np.count_nonzero(gpu.vector()) == 0


def test_batched_states_and_per_shot_pauli():
    """num_states > 1: one launch covers every state; per-state Pauli masks (batched_pauli_func)."""
    n, S = 8, 13
    rng = np.random.default_rng(4)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    gpu = gpu_qv(n, num_states=S)
    gpu.set_state(np.concatenate(states))
    oras = []
    for s in states:
        o = OracleQV(n)
        o.set_state(s)
        oras.append(o)
    for op in opgen.random_ops(8, n, 25):
        opgen.apply(gpu, op)
        for o in oras:
            opgen.apply(o, op)
    masks = np.zeros((S, 4), dtype=np.uint64)
    paulis = opgen.random_paulis(9, n, S, max_weight=4)
    for i, (qs, p) in enumerate(paulis):
        from oracle.oracle import pauli_masks
        x, zm, ny, _ = pauli_masks(qs, p)
        apply_it = int(i % 3 != 0)
        masks[i] = (x, zm, ny, apply_it)
        if apply_it:
            oras[i].apply_pauli(qs, p)
    gpu.apply_batched_pauli_ops(masks)
    got = gpu.vector().reshape(S, -1)
    for i, o in enumerate(oras):
        assert np.max(np.abs(got[i] - o.vector())) < 1e-12, i
    np.testing.assert_allclose(gpu.norm(), [o.norm() for o in oras], atol=1e-12)
    qs = [0, 5, 3]
    np.testing.assert_allclose(gpu.probabilities(qs), [o.probabilities(qs) for o in oras], atol=1e-13)
    np.testing.assert_allclose(gpu.expval_pauli([1, 6], "XZ"), [o.expval_pauli([1, 6], "XZ") for o in oras],
                               atol=1e-12)
    r = rng.random((S, 40))
    got_s = gpu.sample_measure(r)
    for i, o in enumerate(oras):
        assert np.array_equal(got_s[i], o.sample_measure(r[i]))


def test_global_qubits_resolved_from_chunk_index():
    """Sharded chunk: diagonal / control qubits >= n_local need no data movement (base_index_ rule)."""
    n_local, n_global = 9, 11
    rng = np.random.default_rng(6)
    full = opgen.random_state(rng, n_global)
    ora = OracleQV(n_global)
    ora.set_state(full)
    chunks = []
    for c in range(1 << (n_global - n_local)):
        g = gpu_qv(n_local)
        g.chunk_setup(n_global, c)
        g.set_state(full[c << n_local:(c + 1) << n_local])
        chunks.append(g)
    ops = [("apply_diagonal_matrix", ([2, 10, 9, 5], np.exp(2j * np.pi * rng.random(16)))),
           ("apply_diagonal_matrix", ([10, 9], np.exp(2j * np.pi * rng.random(4)))),
           ("apply_mcx", ([10, 3, 1],)), ("apply_mcy", ([9, 10, 0],)), ("apply_mcswap", ([10, 4, 7],)),
           ("apply_mcphase", ([9, 2, 10], np.exp(0.7j))), ("apply_mcphase", ([9, 10], np.exp(0.2j))),
           ("apply_mcu", ([10, 9, 6], opgen.colmajor(opgen.haar_unitary(rng, 2)))),
           ("apply_mcu", ([3, 10], opgen.colmajor(np.diag(np.exp(1j * rng.random(2))))))]
    for op in ops:
        opgen.apply(ora, op)
        for g in chunks:
            opgen.apply(g, op)
    got = np.concatenate([g.vector() for g in chunks])
    assert np.max(np.abs(got - ora.vector())) < 1e-13


def test_error_behaviour_matches_reference_exceptions():
    import qiskit_aer_b200 as q
    gpu = gpu_qv(4)
    with pytest.raises(q.B200Error):
        gpu.apply_mcx([0, 7])  # out-of-range qubit
    with pytest.raises(q.B200Error):
        gpu.apply_pauli([0, 1], "XQ")  # invalid Pauli (qubitvector.hpp:2267)
    with pytest.raises(q.B200Error):
        gpu.apply_mcx([1, 1])


@pytest.mark.parametrize("n", [26, 29])
def test_large_state_size_independent_properties(n):
    """Sizes the oracle cannot hold: unitarity, involution and linearity properties on the GPU alone."""
    rng = np.random.default_rng(n)
    gpu = gpu_qv(n)
    # spread the state with a layer of random 2-qubit unitaries
    perm = rng.permutation(n)
    for i in range(n // 2):
        gpu.apply_matrix([int(perm[2 * i]), int(perm[2 * i + 1])], opgen.colmajor(opgen.haar_unitary(rng, 4)))
    assert abs(gpu.norm() - 1.0) < 1e-12
    gpu.checkpoint()
    # U then U^dagger restores the state (k = 5, high / low / mixed qubits)
    for qs in ([0, 1, 2, 3, 4], [n - 1, n - 2, n - 3, n - 4, n - 5], [0, 7, n - 1, 13, 2]):
        U = opgen.haar_unitary(rng, 32)
        gpu.apply_matrix(qs, opgen.colmajor(U))
        assert abs(gpu.norm() - 1.0) < 1e-12
        gpu.apply_matrix(qs, opgen.colmajor(U.conj().T))
    z = gpu.inner_product()
    assert abs(1.0 - z) < 1e-10
    # permutation gates are involutions, bit exact: checksum of probabilities over 12 qubits unchanged
    before = gpu.vector(offset=12345, count=4096)
    for qs in ([3, n - 1], [n - 2, 0, 9]):
        gpu.apply_mcx(qs)
        gpu.apply_mcx(qs)
        gpu.apply_mcswap(qs + [5])
        gpu.apply_mcswap(qs + [5])
    assert np.array_equal(before, gpu.vector(offset=12345, count=4096))
    # probabilities over any qubit subset sum to the norm; Z expval equals p0 - p1
    p = gpu.probabilities([n - 1, 0, 10])
    assert abs(p.sum() - gpu.norm()) < 1e-12
    p1 = gpu.probabilities([n - 1])
    assert abs(gpu.expval_pauli([n - 1], "Z") - (p1[0] - p1[1])) < 1e-12
    # sampled indices follow the marginal of the top qubit
    s = gpu.sample_measure(np.random.default_rng(1).random(20000))
    frac = np.mean((s >> np.uint64(n - 1)) & np.uint64(1))
    assert abs(frac - p1[1]) < 0.02


@pytest.mark.parametrize("n", [12, 13, 16, 19])
def test_gate_sequence_tile_passes(n):
    """b200sv_apply_gate_sequence (tile-blocked multi-gate passes) == the same gates applied one by one."""
    rng = np.random.default_rng(200 + n)
    psi0 = opgen.random_state(rng, n)
    for trial in range(4):
        gates = []
        for _ in range(int(rng.integers(1, 40))):
            k = int(rng.integers(1, 3))
            if trial == 1:   # stress the low (lane / bank) bits
                qs = [int(q) for q in rng.choice(min(n, 6), size=k, replace=False)]
            elif trial == 2:  # chains on a few qubits: many dependent rounds
                qs = [int(q) for q in rng.choice(4, size=k, replace=False) + (n - 4)]
            else:
                qs = opgen.pick(rng, n, k)
            if rng.random() < 0.3:  # diagonal gates (cz / cp / rzz / phase) take the one-multiply tile forms
                gates.append((qs, opgen.colmajor(np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << k))))))
            else:
                gates.append((qs, opgen.colmajor(opgen.haar_unitary(rng, 1 << k))))
        ora, gpu = OracleQV(n), gpu_qv(n)
        ora.set_state(psi0)
        gpu.set_state(psi0)
        for qs, m in gates:
            ora.apply_matrix(qs, m)
        passes = gpu.apply_gate_sequence(gates)
        assert 1 <= passes <= len(gates)
        assert np.max(np.abs(ora.vector() - gpu.vector())) < 1e-12, (trial, len(gates))
        assert opgen.fidelity_gap(ora.vector(), gpu.vector()) < 1e-10


def test_gate_sequence_quantum_volume_uses_few_passes():
    from qiskit_aer_b200 import circuits, executor
    n = 20
    ops = circuits.quantum_volume(n, 10, seed=3)
    ora, gpu = OracleQV(n), gpu_qv(n)
    executor.apply_ops(ora, ops)
    stats = {}
    executor.apply_ops_queued(gpu, ops, stats)
    assert opgen.fidelity_gap(ora.vector(), gpu.vector()) < 1e-10
    assert stats["passes"] < len(ops) // 3, stats  # 100 gates must not need 100 passes


def test_gate_sequence_small_and_batched_states():
    rng = np.random.default_rng(9)
    for n, S in ((5, 1), (12, 3)):
        states = [opgen.random_state(rng, n) for _ in range(S)]
        gpu = gpu_qv(n, num_states=S)
        gpu.set_state(np.concatenate(states))
        gates = [(opgen.pick(rng, n, 2), opgen.colmajor(opgen.haar_unitary(rng, 4))) for _ in range(12)]
        gpu.apply_gate_sequence(gates)
        got = gpu.vector().reshape(S, -1)
        for i, st in enumerate(states):
            o = OracleQV(n)
            o.set_state(st)
            for qs, m in gates:
                o.apply_matrix(qs, m)
            assert np.max(np.abs(got[i] - o.vector())) < 1e-12


@pytest.mark.parametrize("n", [13, 14, 17, 20])
def test_gate_sequence_single_precision_tile_passes(n):
    """complex64 states >= 13 qubits flush the gate queue through the float tile passes (16-byte slots = amplitude
    pairs over qubit 0, rounds A(1,2)|A(0,1) [+ B(3,4)], 1-qubit gates promoted to 4x4): same result as the gates
    applied one by one by the double-precision oracle, to single-precision tolerance."""
    rng = np.random.default_rng(300 + n)
    psi0 = opgen.random_state(rng, n)
    for trial in range(4):
        gates = []
        for _ in range(int(rng.integers(1, 48))):
            k = int(rng.integers(1, 3))
            if trial == 1:    # stress qubit 0 (inside the slot) and the low coalescing bits
                qs = [int(q) for q in rng.choice(min(n, 5), size=k, replace=False)]
            elif trial == 2:  # chains on a few high qubits
                qs = [int(q) for q in rng.choice(4, size=k, replace=False) + (n - 4)]
            else:
                qs = opgen.pick(rng, n, k)
            if rng.random() < 0.25:
                gates.append((qs, opgen.colmajor(np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << k))))))
            else:
                gates.append((qs, opgen.colmajor(opgen.haar_unitary(rng, 1 << k))))
        ora, gpu = OracleQV(n), gpu_qv(n, np.complex64)
        ora.set_state(psi0)
        gpu.set_state(psi0.astype(np.complex64))
        for qs, m in gates:
            ora.apply_matrix(qs, m)
        passes = gpu.apply_gate_sequence(gates)
        assert 1 <= passes <= len(gates)
        got = gpu.vector().astype(np.complex128)
        assert np.max(np.abs(ora.vector() - got)) < 1e-3 / np.sqrt(1 << n), (trial, len(gates))  # amplitudes ~ 2^(-n/2)
        assert opgen.fidelity_gap(ora.vector(), got) < 1e-5


def test_gate_sequence_single_precision_quantum_volume_and_batch():
    from qiskit_aer_b200 import circuits, executor
    n = 18
    ops = circuits.quantum_volume(n, 8, seed=5)
    ora, gpu = OracleQV(n), gpu_qv(n, np.complex64)
    executor.apply_ops(ora, ops)
    stats = {}
    executor.apply_ops_queued(gpu, ops, stats)
    assert opgen.fidelity_gap(ora.vector(), gpu.vector().astype(np.complex128)) < 1e-5
    assert stats["passes"] < len(ops) // 3, stats
    # several states in one container: tiles never straddle two states
    rng = np.random.default_rng(11)
    S, n = 3, 13
    states = [opgen.random_state(rng, n) for _ in range(S)]
    gpu = gpu_qv(n, np.complex64, num_states=S)
    gpu.set_state(np.concatenate(states).astype(np.complex64))
    gates = [(opgen.pick(rng, n, 2), opgen.colmajor(opgen.haar_unitary(rng, 4))) for _ in range(12)]
    gpu.apply_gate_sequence(gates)
    got = gpu.vector().reshape(S, -1).astype(np.complex128)
    for i, st in enumerate(states):
        o = OracleQV(n)
        o.set_state(st)
        for qs, m in gates:
            o.apply_matrix(qs, m)
        assert opgen.fidelity_gap(o.vector(), got[i]) < 1e-5


def test_concurrent_host_threads_on_different_handles():
    """Aer calls the vector from OpenMP threads (one State per shot / chunk group,
    circuit_executor.hpp:958, parallel_state_executor.hpp:831-840): the ABI must be re-entrant."""
    import threading
    n, nthreads = 12, 8
    results, errors = [None] * nthreads, []

    def work(t):
        try:
            rng = np.random.default_rng(500 + t)
            ora, gpu = OracleQV(n), gpu_qv(n)
            for op in opgen.random_ops(700 + t, n, 60):
                opgen.apply(ora, op)
                opgen.apply(gpu, op)
            gates = [(opgen.pick(rng, n, 2), opgen.colmajor(opgen.haar_unitary(rng, 4))) for _ in range(10)]
            for qs, m in gates:
                ora.apply_matrix(qs, m)
            gpu.apply_gate_sequence(gates)
            results[t] = float(np.max(np.abs(ora.vector() - gpu.vector())))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert max(results) < 1e-12, results


def test_full_size_qv33_round_trip_across_engines():
    """BASELINE size (33 qubits, 128 GiB): forward through the tile-blocked gate queue, backward (inverse gates,
    reverse order) through the dense per-gate kernels -> the register must return to |0...0>.  Size-independent
    property, no second copy of the state needed; also exercises 64-bit indexing at full scale."""
    import torch
    free, total = torch.cuda.mem_get_info()
    n = 33
    if free < (16 << n) + (4 << 30):
        pytest.skip("needs a 128 GiB state")
    from qiskit_aer_b200 import circuits
    ops = circuits.quantum_volume(n, 2, seed=5)
    gpu = gpu_qv(n)
    gates = [(op[1], opgen.colmajor(op[2])) for op in ops]
    passes = gpu.apply_gate_sequence(gates)
    assert passes < len(gates)
    assert abs(gpu.norm() - 1.0) < 1e-12
    p_top = gpu.probabilities([n - 1, 0])
    assert abs(p_top.sum() - 1.0) < 1e-12
    for op in reversed(ops):
        gpu.apply_matrix(op[1], opgen.colmajor(np.asarray(op[2]).conj().T))
    amp0 = gpu.vector(offset=0, count=1)[0]
    assert abs(1.0 - amp0) < 1e-10
    assert abs(gpu.norm() - 1.0) < 1e-12
    # bit-exact permutation at the top of the index range
    gpu.apply_mcx([n - 1])
    assert abs(gpu.vector(offset=1 << (n - 1), count=1)[0] - amp0) == 0.0
    s = gpu.sample_measure(np.array([0.5]))
    assert int(s[0]) == 1 << (n - 1)
    gpu.close()


@pytest.mark.parametrize("n,dtype", [(4, np.complex128), (9, np.complex128), (13, np.complex128), (17, np.complex128),
                                     (15, np.complex64)])
def test_wide_diagonal_layer(n, dtype):
    """b200sv_apply_diagonal_layer: many commuting diagonal 1-/2-qubit gates over all qubits in one pass == the gates
    applied one by one (apply_diagonal_matrix, qubitvector.hpp:1343)."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(n)
    psi0 = opgen.random_state(rng, n, dtype)
    ora, gpu = OracleQV(n, dtype), gpu_qv(n, dtype)
    ora.set_state(psi0)
    gpu.set_state(psi0)
    gates = []
    for _ in range(6 * n):
        k = int(rng.integers(1, 3))
        qs = opgen.pick(rng, n, k)
        d = np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << k)) * rng.uniform(0.9, 1.1, 1 << k)
        gates.append((qs, d))
    gpu.apply_diagonal_layer(gates)
    for qs, d in gates:
        ora.apply_diagonal_matrix(qs, d)
    a, b = ora.vector(), gpu.vector()
    scale = np.max(np.abs(a))
    assert np.max(np.abs(a - b)) < (1e-11 if dtype == np.complex128 else 2e-5) * scale


def test_wide_diagonal_layer_on_a_batch_and_qft_front_end():
    import qiskit_aer_b200 as q
    from qiskit_aer_b200 import circuits, executor, fusion
    # batched container: every state gets the same layer
    n, S = 13, 5
    rng = np.random.default_rng(3)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    gpu = q.QubitVectorB200(n, np.complex128, num_states=S)
    gpu.set_state(np.concatenate(states))
    gates = [(opgen.pick(rng, n, 2), np.exp(1j * rng.uniform(0, 6.28, 4))) for _ in range(40)]
    gpu.apply_diagonal_layer(gates)
    got = gpu.vector().reshape(S, -1)
    for s in range(S):
        ora = OracleQV(n)
        ora.set_state(states[s])
        for qs, d in gates:
            ora.apply_diagonal_matrix(qs, d)
        assert np.max(np.abs(got[s] - ora.vector())) < 1e-12
    # QFT through the fusion front end with wide diagonal layers
    n = 16
    ops = circuits.qft(n)
    fused = fusion.fuse(ops, max_qubit=4, max_diag_qubit=40, max_table_qubit=6)
    assert any(op[0] == "diag_layer" for op in fused)
    g, o = q.QubitVectorB200(n), OracleQV(n)
    psi0 = opgen.random_state(rng, n)
    g.set_state(psi0)
    o.set_state(psi0)
    executor.apply_ops(g, fused)
    executor.apply_ops(o, ops)
    assert opgen.fidelity_gap(o.vector(), g.vector()) < 1e-10


def test_gate_queue_splits_wide_diagonal_layers_out():
    """A queue rich in commuting controlled phases (QFT): the flush regroups it into tile passes for the dense gates and
    one-pass diagonal layers (b200sv_apply_gate_sequence -> fuse_assign -> diag_layer_kernel); result == gate by gate."""
    from qiskit_aer_b200 import circuits, executor, fusion
    n = 17
    rng = np.random.default_rng(5)
    psi0 = opgen.random_state(rng, n)
    ops = circuits.qft(n) + circuits.random_noisy_circuit(n, 2, seed=3)
    seq = []
    for op in ops:
        qs, U = fusion.op_matrix(op)
        seq.append((list(qs), executor.colmajor(U)))
    ora, gpu = OracleQV(n), gpu_qv(n)
    ora.set_state(psi0)
    gpu.set_state(psi0)
    passes = gpu.apply_gate_sequence(seq)
    for qs, m in seq:
        ora.apply_matrix(qs, m)
    assert opgen.fidelity_gap(ora.vector(), gpu.vector()) < 1e-10
    assert np.max(np.abs(ora.vector() - gpu.vector())) < 1e-11
    assert passes < len(seq) // 4
