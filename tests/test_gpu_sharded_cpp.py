"""The C++ sharded executor (csrc/sharded.cu, b200sv_sharded_*) against the CPU oracle.

All shards of a register are placed on ONE device here (device ids repeat), so the 1-GPU test tier covers exactly the
code the multi-GPU bench times: the epoch planner, per-shard gate queues -> tile passes, the staged / slab-pipelined
exchange (copy-engine pushes into the partners' staging areas + unstage kernel, with the neighbouring passes run slab
by slab), the in-place peer-swap fallback, and the cross-shard reductions.  On >= 2 GPUs the same tests also spread the
shards over the devices (peer access over NVLink).  Reference behaviour being matched: chunked == unchunked,
/root/reference/test/terra/backends/aer_simulator/test_chunk.py:31-168.
"""
import os

import numpy as np
import pytest

import opgen
from oracle.oracle import OracleQV

pytestmark = pytest.mark.gpu


def _ref(n, ops):
    from qiskit_aer_b200 import executor
    ref = OracleQV(n)
    executor.apply_ops(ref, ops)
    return ref


def _devices(world):
    import torch
    nd = max(1, torch.cuda.device_count())
    return [(r * nd) // world if nd >= world else r % nd for r in range(world)]


def _make(n, world, staging, env=None, dtype=np.complex128):
    from qiskit_aer_b200 import sharded
    keys = ("B200SV_SHARD_MIN_RUN_BITS", "B200SV_SHARD_SLAB_BITS", "B200SV_SHARD_STAGED", "B200SV_SHARD_UNSTAGE",
            "B200SV_SHARD_PUSH_STREAMS", "B200SV_SHARD_MAX_RIDE")
    saved = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env or {})
    try:
        return sharded.ShardedState(n, devices=_devices(world), staging_bytes=staging, dtype=dtype)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _check(st, n, ops, tol=1e-10, amp_tol=1e-12, seed=5):
    import qiskit_aer_b200 as q
    ref = _ref(n, ops)
    st.initialize()
    st.apply_ops(ops)
    stats = st.stats()
    # expectation values BEFORE the qubit order is restored: Z / X / Y factors land on global positions
    for qs, pl in opgen.random_paulis(seed, n, 8, max_weight=4):
        assert abs(st.expval_pauli(qs, pl) - ref.expval_pauli(qs, pl)) < tol, (qs, pl, st.phys)
    assert abs(st.norm() - 1.0) < tol
    rn = q.rng_uniform(99 + seed, 400)
    assert np.array_equal(st.sample_measure(rn), ref.sample_measure(rn))
    got, want = st.vector(), ref.vector()
    assert np.max(np.abs(got - want)) < amp_tol
    assert opgen.fidelity_gap(want, got.astype(np.complex128)) < tol
    return stats


@pytest.mark.parametrize("unstage,streams", [("kernel", 4), ("dma", 1), ("dma", 3)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_qv_staged_pipelined_exchange(world, unstage, streams):
    """Quantum Volume: every exchange staged, slabs cut along free index bits, neighbouring passes ride along.
    Received slabs are moved into place by a kernel or by the copy engines; pushes fan out over several streams."""
    from qiskit_aer_b200 import circuits
    n = 17 + int(np.log2(world))
    ops = circuits.quantum_volume(n, 6, seed=world)
    st = _make(n, world, staging=1 << 22, env={"B200SV_SHARD_MIN_RUN_BITS": "5", "B200SV_SHARD_SLAB_BITS": "2",
                                               "B200SV_SHARD_UNSTAGE": unstage, "B200SV_SHARD_PUSH_STREAMS": str(streams)})
    stats = _check(st, n, ops)
    assert stats["exchanges"] > 0 and stats["staged"] == stats["exchanges"] and stats["inplace"] == 0
    assert stats["overlapped_passes"] > 0 and stats["copies"] > 0
    st.close()


@pytest.mark.parametrize("ride", [0, 1, 3])
def test_overlap_window_sizes(ride):
    """0, 1 or up to 3 passes on either side of an exchange run slab-wise under it."""
    from qiskit_aer_b200 import circuits
    n, world = 19, 4
    ops = circuits.quantum_volume(n, 8, seed=50 + ride)
    st = _make(n, world, staging=1 << 22, env={"B200SV_SHARD_MIN_RUN_BITS": "5", "B200SV_SHARD_SLAB_BITS": "2",
                                               "B200SV_SHARD_MAX_RIDE": str(ride)})
    stats = _check(st, n, ops, seed=21)
    assert stats["staged"] == stats["exchanges"] > 0
    assert (stats["overlapped_passes"] == 0) == (ride == 0)
    st.close()


@pytest.mark.parametrize("world", [2, 8])
def test_qv_small_staging_area_forces_fine_slabs(world):
    """A staging area far smaller than the outgoing data: more slab bits, one buffer."""
    from qiskit_aer_b200 import circuits
    n = 17 + int(np.log2(world))
    ops = circuits.quantum_volume(n, 5, seed=10 + world)
    out_bytes = (1 << 17) * 16
    st = _make(n, world, staging=out_bytes // 4, env={"B200SV_SHARD_MIN_RUN_BITS": "4", "B200SV_SHARD_SLAB_BITS": "3"})
    stats = _check(st, n, ops, seed=7)
    assert stats["staged"] > 0
    st.close()


@pytest.mark.parametrize("world", [2, 4])
def test_qv_inplace_exchange(world):
    """No staging area: the in-place peer swap kernel between two rendezvous."""
    from qiskit_aer_b200 import circuits
    n = 16 + int(np.log2(world))
    ops = circuits.quantum_volume(n, 5, seed=20 + world)
    st = _make(n, world, staging=0)
    stats = _check(st, n, ops, seed=9)
    assert stats["exchanges"] > 0 and stats["inplace"] == stats["exchanges"]
    st.close()


@pytest.mark.parametrize("world", [2, 4])
def test_qft_and_named_gates(world):
    """QFT (h, cp with global controls, swaps) + cx / rz / sx layers: diagonal and controlled ops on global qubits are
    resolved from the shard index, no data moves for them."""
    from qiskit_aer_b200 import circuits
    n = 15 + int(np.log2(world))
    ops = circuits.qft(n) + circuits.random_noisy_circuit(n, 3, seed=4) + circuits.quantum_volume(n, 2, seed=3)
    st = _make(n, world, staging=1 << 21, env={"B200SV_SHARD_MIN_RUN_BITS": "4", "B200SV_SHARD_SLAB_BITS": "2"})
    _check(st, n, ops, seed=11)
    st.close()


def test_wide_ops_take_their_own_kernels():
    """3-qubit unitaries and wide diagonals between the queued gates (direct steps between tile steps)."""
    from qiskit_aer_b200 import circuits
    n, world = 16, 4
    rng = np.random.default_rng(8)
    ops = circuits.quantum_volume(n, 2, seed=1)
    ops.append(("unitary", [1, 9, 4], circuits.haar_unitary(rng, 8)))
    ops.append(("diagonal", [0, 3, 15, 14, 7], np.exp(1j * rng.uniform(0, 6.28, 32))))
    ops += circuits.quantum_volume(n, 2, seed=2)
    ops.append(("gate", "ccx", [15, 2, 6], []))
    ops.append(("gate", "cswap", [14, 3, 8], []))
    ops += circuits.quantum_volume(n, 1, seed=5)
    st = _make(n, world, staging=1 << 20, env={"B200SV_SHARD_MIN_RUN_BITS": "4"})
    _check(st, n, ops, seed=13)
    st.close()


def test_single_precision():
    from qiskit_aer_b200 import circuits
    n, world = 17, 2
    ops = circuits.quantum_volume(n, 4, seed=6)
    st = _make(n, world, staging=1 << 20, env={"B200SV_SHARD_MIN_RUN_BITS": "5", "B200SV_SHARD_SLAB_BITS": "2"}, dtype=np.complex64)
    ref = _ref(n, ops)
    st.initialize()
    st.apply_ops(ops)
    got = st.vector().astype(np.complex128)
    assert opgen.fidelity_gap(ref.vector(), got) < 1e-5
    st.close()


def test_repeated_runs_reuse_the_handle():
    """Sequence counters, event ring and staging buffers survive many exchanges (the bench re-runs one circuit)."""
    from qiskit_aer_b200 import circuits
    n, world = 17, 4
    ops = circuits.quantum_volume(n, 8, seed=33)
    st = _make(n, world, staging=1 << 21, env={"B200SV_SHARD_MIN_RUN_BITS": "5", "B200SV_SHARD_SLAB_BITS": "3"})
    ref = _ref(n, ops).vector()
    for _ in range(12):
        st.initialize()
        st.apply_ops(ops)
    assert np.max(np.abs(st.vector() - ref)) < 1e-12
    assert st.elapsed_ms() >= 0.0
    st.close()
