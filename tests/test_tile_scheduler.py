"""CPU coverage of the tile-pass scheduler (csrc/tile.cu host code): b200sv_selftest_op_sequence partitions an op
list into passes / rounds / warp-local segments exactly as the product path does and INTERPRETS every parameter
block on a host array with the kernels' addressing.  The result must equal the ops applied one by one by the
oracle.  No GPU, no handle: this tests host logic (and the staging / swizzle / round-block address algebra the
kernels share with the interpreter), not the kernels -- those are covered by tests/test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

import opgen
from oracle.oracle import OracleQV

import qiskit_aer_b200  # noqa: F401
from qiskit_aer_b200 import capi, circuits


def selftest(n, state, ops, num_states=1, codes=None):
    """ops: (kind, qubits, matrix-or-slot); kind 1/2 dense column-major, 3 per-state Pauli (slot)."""
    lib = capi.lib()
    nops = len(ops)
    kind = np.zeros(nops, dtype=np.int32)
    slot = np.zeros(nops, dtype=np.int32)
    qs = np.zeros(2 * nops, dtype=np.uint64)
    mats = np.zeros((nops, 16), dtype=np.complex128)
    for i, (k, q, m) in enumerate(ops):
        kind[i] = k
        qs[2 * i:2 * i + len(q)] = q
        if k == 3:
            slot[i] = m
        else:
            m = np.asarray(m, dtype=np.complex128).reshape(-1)
            mats[i, :m.size] = m
    passes = C.c_int(0)
    prec = 64 if state.dtype == np.complex128 else 32  # B200SV_F64 / B200SV_F32
    nslots = 0 if codes is None else codes.shape[0]
    capi.check(lib.b200sv_selftest_op_sequence(
        n, num_states, prec, C.c_void_p(state.ctypes.data), nops, kind.ctypes.data_as(C.POINTER(C.c_int)),
        qs.ctypes.data_as(C.POINTER(C.c_uint64)), mats.ctypes.data_as(C.POINTER(C.c_double)),
        slot.ctypes.data_as(C.POINTER(C.c_int)),
        None if codes is None else codes.ctypes.data_as(C.POINTER(C.c_uint8)), nslots, C.byref(passes)))
    return passes.value


def random_ops(rng, n, count, mode):
    ops = []
    for _ in range(count):
        k = int(rng.integers(1, 3))
        if mode == "low":
            qs = [int(q) for q in rng.choice(min(n, 6), size=k, replace=False)]
        elif mode == "chain":
            qs = [int(q) for q in rng.choice(4, size=k, replace=False) + (n - 4)]
        else:
            qs = opgen.pick(rng, n, k)
        if rng.random() < 0.3:
            m = np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << k)))
        else:
            m = opgen.haar_unitary(rng, 1 << k)
        ops.append((k, qs, opgen.colmajor(m)))
    return ops


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [13, 15])
def test_scheduler_random_gate_lists(n, dtype):
    rng = np.random.default_rng(400 + n)
    psi0 = opgen.random_state(rng, n)
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    for mode in ("any", "low", "chain", "any"):
        ops = random_ops(rng, n, int(rng.integers(1, 40)), mode)
        ora = OracleQV(n)
        ora.set_state(psi0)
        for _, qs, m in ops:
            ora.apply_matrix(qs, m)
        state = psi0.astype(dtype)
        passes = selftest(n, state, ops)
        assert 1 <= passes <= len(ops)
        assert opgen.fidelity_gap(ora.vector(), state.astype(np.complex128)) < tol, (mode, len(ops))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_scheduler_quantum_volume_fast_rounds(dtype):
    """QV layers are all dense 2-qubit gates: every round is a fast round, passes carry several gates."""
    n = 14
    gate_ops = circuits.quantum_volume(n, 6, seed=4)
    ops = [(2, list(op[1]), opgen.colmajor(np.asarray(op[2]))) for op in gate_ops]
    rng = np.random.default_rng(1)
    psi0 = opgen.random_state(rng, n)
    ora = OracleQV(n)
    ora.set_state(psi0)
    for _, qs, m in ops:
        ora.apply_matrix(qs, m)
    state = psi0.astype(dtype)
    passes = selftest(n, state, ops)
    assert passes < len(ops) // 3
    assert opgen.fidelity_gap(ora.vector(), state.astype(np.complex128)) < (1e-12 if dtype == np.complex128 else 1e-5)


def test_scheduler_batched_states_with_per_state_paulis():
    """kind 3 ops (sampled Pauli noise, one code per state and slot) ride on the passes of a multi-state container."""
    n, S = 12, 3
    rng = np.random.default_rng(7)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    ops, nslots = [], 0
    for i in range(24):
        if i % 3 == 2:
            ops.append((3, [int(rng.integers(0, n))], nslots))
            nslots += 1
        else:
            k = int(rng.integers(1, 3))
            ops.append((k, opgen.pick(rng, n, k), opgen.colmajor(opgen.haar_unitary(rng, 1 << k))))
    codes = rng.integers(0, 4, size=(nslots, S)).astype(np.uint8)
    state = np.concatenate(states).astype(np.complex128)
    selftest(n, state, ops, num_states=S, codes=codes)
    got = state.reshape(S, -1)
    P = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1, -1])]
    for s_i, st in enumerate(states):
        o = OracleQV(n)
        o.set_state(st)
        for k, qs, m in ops:
            if k == 3:
                o.apply_matrix(qs, opgen.colmajor(P[int(codes[m, s_i])].astype(np.complex128)))
            else:
                o.apply_matrix(qs, m)
        assert np.max(np.abs(got[s_i] - o.vector())) < 1e-12


def test_selftest_rejects_sizes_outside_the_tile_path():
    state = np.zeros(1 << 10, dtype=np.complex128)
    with pytest.raises(capi.B200Error):
        selftest(10, state, [(1, [0], opgen.colmajor(np.eye(2)))])


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_scheduler_absorbs_one_qubit_gates_into_their_two_qubit_neighbours(dtype):
    """rz / sx / u around every cx (a transpiled circuit): the 1-qubit gates are multiplied into the neighbouring dense
    2-qubit gate before the passes are planned; consecutive 1-qubit gates collapse; diagonal 2-qubit gates and
    per-state Paulis are left alone / act as barriers.  Result unchanged."""
    n = 13
    rng = np.random.default_rng(21)
    CX = opgen.colmajor(np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128))
    CZ = opgen.colmajor(np.diag([1, 1, 1, -1]).astype(np.complex128))
    ops = []
    for layer in range(6):
        perm = rng.permutation(n)
        for i in range(n // 2):
            a, b = int(perm[2 * i]), int(perm[2 * i + 1])
            for q in (a, b, a):
                ops.append((1, [q], opgen.colmajor(opgen.haar_unitary(rng, 2))))
            ops.append((2, [a, b], CZ if (layer + i) % 3 == 0 else CX))
            ops.append((1, [b], opgen.colmajor(np.diag(np.exp(1j * rng.uniform(0, 6.28, 2))))))
            if i % 2 == 0:  # cx rz cx (a zz rotation) and a gate on the same pair in the other qubit order: one gate each
                ops.append((2, [a, b], CX))
                ops.append((2, [b, a], opgen.colmajor(opgen.haar_unitary(rng, 4))))
    psi0 = opgen.random_state(rng, n)
    ora = OracleQV(n)
    ora.set_state(psi0)
    for _, qs, m in ops:
        ora.apply_matrix(qs, m)
    state = psi0.astype(dtype)
    passes = selftest(n, state, ops)
    assert passes <= len(ops) // 8, (passes, len(ops))
    assert opgen.fidelity_gap(ora.vector(), state.astype(np.complex128)) < (1e-12 if dtype == np.complex128 else 1e-5)


def test_scheduler_folds_noise_paulis_into_the_next_gate_round():
    """A noisy-circuit pattern: every gate is followed by a sampled Pauli on each of its qubits.  Paulis whose next op on
    that qubit is a gate of the same pass are applied as pre-Paulis of that gate's round, the others get rounds of
    their own; both paths must agree with the oracle, for several states with different draws."""
    n, S = 13, 4
    rng = np.random.default_rng(77)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    ops, nslots = [], 0
    for layer in range(5):
        perm = rng.permutation(n)
        for i in range(n // 2):
            a, b = int(perm[2 * i]), int(perm[2 * i + 1])
            if (layer + i) % 3 == 0:
                ops.append((1, [a], opgen.colmajor(opgen.haar_unitary(rng, 2))))
                ops.append((3, [a], nslots)); nslots += 1
            ops.append((2, [a, b], opgen.colmajor(opgen.haar_unitary(rng, 4))))
            for q in (a, b):
                ops.append((3, [q], nslots)); nslots += 1
    codes = rng.choice(4, size=(nslots, S), p=[0.7, 0.1, 0.1, 0.1]).astype(np.uint8)  # far noisier than real models
    state = np.concatenate(states).astype(np.complex128)
    passes = selftest(n, state, ops, num_states=S, codes=codes)
    assert passes < len(ops) // 6
    got = state.reshape(S, -1)
    P = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1, -1])]
    for s_i, st in enumerate(states):
        o = OracleQV(n)
        o.set_state(st)
        for k, qs, m in ops:
            o.apply_matrix(qs, opgen.colmajor(P[int(codes[m, s_i])].astype(np.complex128)) if k == 3 else m)
        assert np.max(np.abs(got[s_i] - o.vector())) < 1e-12


_PAULI = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1, -1])]
_CX_FIRST_CONTROLS = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)   # index = bit(q0) + 2 bit(q1)
_CX_SECOND_CONTROLS = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)


def test_noisy_layers_against_the_per_pass_budgets():
    """Twelve qubits = one tile: everything could ride on one pass, so the per-pass budgets (matrix slots, ops, Pauli
    table entries -- a round with two folded cx gates adds up to eight) are what cuts the circuit into passes.  Every
    gate is followed by Paulis with a high hit rate; must equal the oracle."""
    n, S = 12, 2
    rng = np.random.default_rng(4242)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    ops, nslots = [], 0
    for layer in range(14):
        for q in range(n):
            ops.append((1, [q], opgen.colmajor(opgen.haar_unitary(rng, 2))))
            ops.append((3, [q], nslots)); nslots += 1
        perm = rng.permutation(n)
        for i in range(n // 2):
            a, b = int(perm[2 * i]), int(perm[2 * i + 1])
            ops.append((2, [a, b], opgen.colmajor(_CX_FIRST_CONTROLS if (layer + i) % 2 else _CX_SECOND_CONTROLS)))
            for q in (a, b):
                ops.append((3, [q], nslots)); nslots += 1
    codes = rng.choice(4, size=(nslots, S), p=[0.4, 0.2, 0.2, 0.2]).astype(np.uint8)
    state = np.concatenate(states).astype(np.complex128)
    passes = selftest(n, state, ops, num_states=S, codes=codes)
    assert passes >= 8  # 14 layers x (12 + 12 + 6 + 12) ops against 64 ops per pass
    got = state.reshape(S, -1)
    for s_i, st in enumerate(states):
        o = OracleQV(n)
        o.set_state(st)
        for k, qs, m in ops:
            o.apply_matrix(qs, opgen.colmajor(_PAULI[int(codes[m, s_i])].astype(np.complex128)) if k == 3 else m)
        assert np.max(np.abs(got[s_i] - o.vector())) < 1e-12


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bare_cx_gates_of_noisy_passes_become_load_permutations(seed):
    """The config-5 pattern: layers of 1-qubit gates and cx gates (either qubit as the control), a sampled Pauli after
    every gate on each of its qubits.  In passes that carry Paulis a bare cx takes no matrix slot and no arithmetic: it is
    folded into the load offsets of its round (TileRound::eoff_ld); 1-qubit gates that are diagonal or real take the
    cheaper per-kind forms of a pair slot.  Must equal the oracle; and because matrix slots are
    what limits a noisy pass, the circuit must fit in fewer passes than one per 16 gates."""
    n, S = 14, 3
    rng = np.random.default_rng(900 + seed)
    states = [opgen.random_state(rng, n) for _ in range(S)]
    ops, nslots, ngates = [], 0, 0
    for layer in range(6):
        for q in range(n):
            kind = int(rng.integers(4))  # dense, diagonal (rz), real (h / ry), dense again: slots pair them by kind
            if kind == 1:
                u = np.diag(np.exp(1j * rng.uniform(0, 6.28, 2)))
            elif kind == 2:
                th = rng.uniform(0, 6.28)
                u = np.array([[1, 1], [1, -1]]) / np.sqrt(2) if rng.random() < 0.5 else np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            else:
                u = opgen.haar_unitary(rng, 2)
            if layer == 4 and q % 3 == 0:
                continue  # unpaired 1-qubit gates (an absent partner)
            ops.append((1, [q], opgen.colmajor(np.asarray(u, dtype=np.complex128))))
            ops.append((3, [q], nslots)); nslots += 1
            ngates += 1
        perm = rng.permutation(n)
        for i in range(n // 2):
            a, b = int(perm[2 * i]), int(perm[2 * i + 1])
            ops.append((2, [a, b], opgen.colmajor(_CX_FIRST_CONTROLS if rng.random() < 0.5 else _CX_SECOND_CONTROLS)))
            ngates += 1
            for q in (a, b):
                ops.append((3, [q], nslots)); nslots += 1
    codes = rng.choice(4, size=(nslots, S), p=[0.85, 0.05, 0.05, 0.05]).astype(np.uint8)
    state = np.concatenate(states).astype(np.complex128)
    passes = selftest(n, state, ops, num_states=S, codes=codes)
    assert passes <= ngates // 16
    got = state.reshape(S, -1)
    for s_i, st in enumerate(states):
        o = OracleQV(n)
        o.set_state(st)
        for k, qs, m in ops:
            o.apply_matrix(qs, opgen.colmajor(_PAULI[int(codes[m, s_i])].astype(np.complex128)) if k == 3 else m)
        assert np.max(np.abs(got[s_i] - o.vector())) < 1e-12


def _plan_only_passes(n, ops):
    lib = capi.lib()
    nops = len(ops)
    kind = np.zeros(nops, dtype=np.int32)
    slot = np.zeros(nops, dtype=np.int32)
    qs = np.zeros(2 * nops, dtype=np.uint64)
    mats = np.zeros((nops, 16), dtype=np.complex128)
    for i, (q, m) in enumerate(ops):
        kind[i] = len(q)
        qs[2 * i:2 * i + len(q)] = q
        m = np.asarray(m, dtype=np.complex128).reshape(-1)
        mats[i, :m.size] = m
    passes = C.c_int(0)
    capi.check(lib.b200sv_selftest_op_sequence(
        n, 1, 64, None, nops, kind.ctypes.data_as(C.POINTER(C.c_int)), qs.ctypes.data_as(C.POINTER(C.c_uint64)),
        mats.ctypes.data_as(C.POINTER(C.c_double)), slot.ctypes.data_as(C.POINTER(C.c_int)), None, 0, C.byref(passes)))
    return passes.value


def test_pass_packing_quality_on_the_headline_circuit():
    """Plan-only mode (no state): the 33-qubit, depth-10 Quantum Volume circuit of the bench (160 SU(4) gates) must
    keep fitting in <= 18 HBM passes (first-fit in program order needs 21-24; the ready-set packer 16-18)."""
    for seed in (1234, 1, 2):
        ops = circuits.quantum_volume(33, 10, seed)
        passes = _plan_only_passes(33, [(list(op[1]), opgen.colmajor(np.asarray(op[2]))) for op in ops])
        assert 14 <= passes <= 18, (seed, passes)


@pytest.mark.parametrize("block", range(4))
def test_scheduler_randomized_op_mixes(block):
    """Random mixes of dense / diagonal 1- and 2-qubit gates and per-state Paulis (densities from none to 80 %, optionally
    concentrated on five qubits, 1-3 states, double and single precision): absorption, pass packing, slot / generic
    rounds, folded and stand-alone Paulis, leftovers.  1000 seeds of this generator were run when the scheduler was
    written; 60 stay in the suite."""
    for seed in range(15 * block, 15 * block + 15):
        rng = np.random.default_rng(seed)
        n, S = int(rng.integers(12, 15)), int(rng.integers(1, 4))
        nops = int(rng.integers(1, 140))
        ppauli, pdiag, p1q = rng.choice([0.0, 0.2, 0.5, 0.8]), rng.choice([0.0, 0.2, 0.6]), rng.choice([0.1, 0.5, 0.9])
        pool = min(n, 5) if rng.random() < 0.3 else n
        ops, nslots = [], 0
        for _ in range(nops):
            if rng.random() < ppauli:
                ops.append((3, [int(rng.integers(0, pool))], nslots))
                nslots += 1
                continue
            k = 1 if rng.random() < p1q else 2
            qs = [int(q) for q in rng.choice(pool, size=k, replace=False)]
            m = np.diag(np.exp(1j * rng.uniform(0, 6.28, 1 << k))) if rng.random() < pdiag else opgen.haar_unitary(rng, 1 << k)
            if k == 2 and seed % 2 and rng.random() < 0.4:  # odd seeds: bare cx gates (load permutations in noisy passes)
                m = _CX_FIRST_CONTROLS if rng.random() < 0.5 else _CX_SECOND_CONTROLS
            ops.append((k, qs, opgen.colmajor(m)))
        if all(o[0] == 3 for o in ops):
            ops.append((1, [0], opgen.colmajor(opgen.haar_unitary(rng, 2))))
        codes = rng.choice(4, size=(max(nslots, 1), S), p=[0.5, 0.2, 0.15, 0.15]).astype(np.uint8)
        dtype = np.complex128 if (nslots or rng.random() < 0.6) else np.complex64
        if dtype == np.complex64:
            n = max(n, 13)
        states = [opgen.random_state(rng, n) for _ in range(S)]
        state = np.concatenate(states).astype(dtype)
        selftest(n, state, ops, num_states=S, codes=codes if nslots else None)
        got = state.reshape(S, -1).astype(np.complex128)
        for si, st in enumerate(states):
            o = OracleQV(n)
            o.set_state(st)
            for k, qs, m in ops:
                o.apply_matrix(qs, opgen.colmajor(_PAULI[int(codes[m, si])].astype(np.complex128)) if k == 3 else m)
            gap = opgen.fidelity_gap(o.vector(), got[si])
            assert gap < (1e-11 if dtype == np.complex128 else 1e-5), (seed, n, S, nops, gap)


@pytest.mark.parametrize("world,max_exchanges,max_passes", [(2, 2, 21), (4, 2, 20), (8, 3, 24)])
def test_sharded_plan_quality(world, max_exchanges, max_passes):
    """Epoch planner + pass packer on the bench's sharded workloads (33 local qubits per GPU), without any state:
    number of global-qubit exchange steps and of HBM passes per rank."""
    from qiskit_aer_b200 import sharded
    g = int(np.log2(world))
    n = 33 + g
    ops = circuits.quantum_volume(n, 10, 1234)
    r = object.__new__(sharded.ShardedRunner)  # the planner only reads these fields
    r.nl, r.gbits, r.n, r.phys, r.multi_swap = 33, g, n, list(range(n)), True
    r.world, r.rank, r.min_run_bits, r.exchange = world, 0, 20, "p2p"
    plan = r.plan(ops)
    exchanges = sum(1 for op in plan if op[0] in ("swap", "mswap"))
    if world == 2:
        exchanges = sum(1 for op in plan if op[0] == "swap")
    passes, seg = 0, []
    for op in plan + [("swap",)]:
        if op[0] in ("swap", "mswap"):
            if seg:
                passes += _plan_only_passes(33, [(list(o[1]), opgen.colmajor(np.asarray(o[2]))) for o in seg])
            seg = []
        else:
            seg.append(op)
    assert exchanges <= max_exchanges, exchanges
    assert passes <= max_passes, passes
    assert sorted(r.phys) == list(range(n))  # still a permutation of the physical positions
