"""CPU tests of the host-side passes: the engine's fusion (role of src/transpile/fusion.hpp), the circuit
generators, and the sharded swap planner's invariants (role of src/transpile/cacheblocking.hpp)."""
import numpy as np
import pytest

import opgen
from oracle.oracle import OracleQV


def _run(ops, n, psi):
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import executor
    o = OracleQV(n)
    o.set_state(psi)
    executor.apply_ops(o, ops)
    return o.vector()


@pytest.mark.parametrize("max_qubit", [2, 3, 4, 5])
@pytest.mark.parametrize("max_diag", [5, 10, 12])
def test_fusion_is_equivalent(max_qubit, max_diag):
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import circuits, fusion
    rng = np.random.default_rng(max_qubit * 31 + max_diag)
    for n, ops in ((9, circuits.quantum_volume(9, 5, 1)), (11, circuits.qft(11)),
                   (8, circuits.random_noisy_circuit(8, 4, 2) + circuits.qft(8))):
        psi = opgen.random_state(rng, n)
        fused = fusion.fuse(ops, max_qubit=max_qubit, max_diag_qubit=max_diag)
        assert len(fused) <= len(ops)
        assert all(len(op[1]) <= (max(max_diag, max_qubit) if op[0] == "diagonal" else max_qubit) for op in fused)
        assert opgen.fidelity_gap(_run(ops, n, psi), _run(fused, n, psi)) < 1e-12


def test_fusion_respects_non_commuting_order():
    """H on a qubit between two controlled phases must not be commuted past them."""
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import fusion
    ops = [("gate", "cp", [0, 1], [0.3]), ("gate", "h", [1], []), ("gate", "cp", [1, 2], [0.9]),
           ("gate", "h", [0], []), ("gate", "cp", [0, 2], [1.7]), ("gate", "h", [2], []), ("gate", "cp", [0, 1], [0.2])]
    rng = np.random.default_rng(0)
    psi = opgen.random_state(rng, 3)
    for mq in (1, 2, 3):
        fused = fusion.fuse(ops, max_qubit=mq, max_diag_qubit=3)
        assert opgen.fidelity_gap(_run(ops, 3, psi), _run(fused, 3, psi)) < 1e-13


def test_amplitude_update_accounting():
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import circuits
    n = 10
    assert circuits.amplitudes_written(circuits.quantum_volume(n, 3, 0), n) == 3 * (n // 2) << n
    q = circuits.qft(4)
    # 4 H (2^n each) + 6 cp (2^(n-2) each) + 2 swaps (2^(n-1) each)   -- BASELINE.md section 3
    assert circuits.amplitudes_written(q, 4) == 4 * 16 + 6 * 4 + 2 * 8


@pytest.mark.parametrize("world,n", [(2, 12), (4, 13), (8, 15)])
def test_epoch_planner_invariants(world, n):
    """Every gate is emitted exactly once, in an order consistent with its qubit dependencies, with all of its
    non-diagonal qubits on local physical positions; swaps are (local, global-bit) pairs."""
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import circuits, sharded
    r = sharded.ShardedRunner.__new__(sharded.ShardedRunner)
    r.n, r.world, r.gbits, r.rank, r.min_run_bits = n, world, int(np.log2(world)), 0, 6
    r.nl = n - r.gbits
    r.phys = list(range(n))
    r.multi_swap = True
    ops = circuits.quantum_volume(n, 8, seed=world) + circuits.qft(n)
    plan = r.plan(ops)
    phys = list(range(n))          # replay the map
    emitted = []
    for p in plan:
        if p[0] in ("swap", "mswap"):
            pairs = [(p[1], p[2])] if p[0] == "swap" else list(zip(p[1], p[2]))
            assert len({a for a, _ in pairs}) == len(pairs) and len({b for _, b in pairs}) == len(pairs)
            for lpos, gb in pairs:
                gpos = gb + r.nl
                assert 0 <= lpos < r.nl and r.nl <= gpos < n
                inv = {pp: q for q, pp in enumerate(phys)}
                a, b = inv[lpos], inv[gpos]
                phys[a], phys[b] = gpos, lpos
            continue
        qs = p[1] if p[0] in ("unitary", "diagonal") else p[2]
        inv = {pp: q for q, pp in enumerate(phys)}
        logical = [inv[x] for x in qs]
        if p[0] == "unitary":
            assert all(x < r.nl for x in qs)
        elif p[0] == "gate" and p[1] in ("h", "swap"):
            assert all(x < r.nl for x in (qs if p[1] == "swap" else qs[-1:]))
        emitted.append((p[0], p[1] if p[0] == "gate" else None, tuple(logical)))
    assert phys == r.phys
    want = [(op[0], op[1] if op[0] == "gate" else None, tuple(op[1] if op[0] != "gate" else op[2])) for op in ops]
    assert sorted(map(str, emitted)) == sorted(map(str, want))
    # dependency order: for every qubit, the subsequence of gates touching it is unchanged
    for q in range(n):
        assert [e for e in emitted if q in e[2]] == [w for w in want if q in w[2]]
    nsw = sum(1 if p[0] == "swap" else len(p[1]) if p[0] == "mswap" else 0 for p in plan)
    assert 0 < nsw <= 4 * r.gbits * 9  # far fewer than one exchange per layer and global qubit


@pytest.mark.parametrize("seed", range(12))
def test_cpp_epoch_planner_equals_python_reference(seed):
    """b200sv_plan_epochs (csrc/planner.cu, what ShardedRunner.plan calls) against the same algorithm in Python
    (ShardedRunner._plan_py): identical plans and final qubit maps on random circuits mixing dense / diagonal ops and
    named gates with controls, random initial maps, pairwise and all-to-all swap modes."""
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import circuits, sharded
    rng = np.random.default_rng(seed)
    n = int(rng.integers(8, 15))
    gbits = int(rng.integers(1, 4))
    ops = []
    for _ in range(int(rng.integers(5, 120))):
        kind = rng.integers(0, 6)
        if kind <= 1:
            k = int(rng.integers(1, 4))
            ops.append(("unitary", [int(q) for q in rng.choice(n, size=k, replace=False)], np.eye(1 << k)))
        elif kind == 2:
            k = int(rng.integers(1, 4))
            ops.append(("diagonal", [int(q) for q in rng.choice(n, size=k, replace=False)], np.ones(1 << k)))
        else:
            name = ["h", "cx", "cp", "swap", "ccx"][int(rng.integers(0, 5))]
            k = {"h": 1, "cx": 2, "cp": 2, "swap": 2, "ccx": 3}[name]
            ops.append(("gate", name, [int(q) for q in rng.choice(n, size=k, replace=False)], [0.3] if name == "cp" else []))
    ops += circuits.quantum_volume(n, 2, seed)
    phys0 = [int(x) for x in rng.permutation(n)]
    plans = []
    for fn in ("plan", "_plan_py"):
        r = sharded.ShardedRunner.__new__(sharded.ShardedRunner)
        r.n, r.gbits, r.nl, r.world, r.rank = n, gbits, n - gbits, 1 << gbits, 0
        r.min_run_bits = int(rng.integers(0, n)) if fn == "plan" else plans[0][2]
        r.multi_swap = bool(seed % 2)
        r.phys = list(phys0)
        mrb = r.min_run_bits
        plan = getattr(r, fn)(ops)
        plans.append(([(p[0], [int(x) for x in p[1]], None) if p[0] in ("unitary", "diagonal") else
                       (p[0], p[1], [int(x) for x in p[2]]) if p[0] == "gate" else
                       (p[0], p[1], p[2]) for p in plan], list(r.phys), mrb))
    assert plans[0][0] == plans[1][0]
    assert plans[0][1] == plans[1][1]


@pytest.mark.parametrize("case", ["qft", "qv", "mixed0", "mixed1", "mixed2"])
@pytest.mark.parametrize("mq,md", [(4, 16), (5, 10), (3, 12)])
def test_cpp_fusion_equals_python_reference(case, mq, md):
    """b200sv_fuse_assign / b200sv_fuse_block_matrix (what fusion.fuse calls) against fusion._fuse_py: same blocks in the
    same order, matrices / diagonals equal to rounding."""
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import circuits, fusion
    n = 11
    if case == "qft":
        ops = circuits.qft(n)
    elif case == "qv":
        ops = circuits.quantum_volume(n, 6, 3)
    else:
        rng = np.random.default_rng(int(case[-1]))
        ops = []
        for _ in range(150):
            r = rng.integers(0, 6)
            if r == 0:
                ops.append(("gate", "h", [int(rng.integers(0, n))], []))
            elif r == 1:
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                ops.append(("gate", "cp", [a, b], [float(rng.uniform(0, 3))]))
            elif r == 2:
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                ops.append(("gate", "cx", [a, b], []))
            elif r == 3:
                k = int(rng.integers(1, 4))
                ops.append(("diagonal", [int(x) for x in rng.choice(n, size=k, replace=False)],
                            np.exp(1j * rng.uniform(0, 6.28, 1 << k))))
            else:
                k = int(rng.integers(1, 3))
                ops.append(("unitary", [int(x) for x in rng.choice(n, size=k, replace=False)], opgen.haar_unitary(rng, 1 << k)))
    a = fusion.fuse(ops, max_qubit=mq, max_diag_qubit=md)
    b = fusion._fuse_py(ops, max_qubit=mq, max_diag_qubit=md)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[0] == y[0] and list(x[1]) == list(y[1])
        assert np.max(np.abs(np.asarray(x[2]) - np.asarray(y[2]))) < 1e-13


def test_sharded_plan_only_counts():
    """Host-only view of the C++ sharded executor: exchanges, staged / in-place decisions, slab counts."""
    from qiskit_aer_b200 import circuits, sharded
    ops = circuits.quantum_volume(36, 10, seed=1234)
    big = sharded.ShardedState.plan_only(36, 8, ops, 34 << 30)
    assert big["exchanges"] == 3 and big["qubit_swaps"] == 9 and big["staged"] == 3 and big["inplace"] == 0
    assert big["passes"] <= 23 and big["passes_overlapped"] >= 6 and big["slabs_min"] >= 4
    # 112 GiB leave a shard per exchange: a 20 GiB staging area still pipelines (8 slabs, one buffer) ...
    mid = sharded.ShardedState.plan_only(36, 8, ops, 20 << 30)
    assert mid["staged"] == 3 and mid["slabs_min"] >= 8
    # ... a 1 GiB one does not: in-place peer swaps
    small = sharded.ShardedState.plan_only(36, 8, ops, 1 << 30)
    assert small["staged"] == 0 and small["inplace"] == 3
    one = sharded.ShardedState.plan_only(33, 1, circuits.quantum_volume(33, 10, seed=1234), 0)
    assert one["exchanges"] == 0 and one["passes"] == 17
