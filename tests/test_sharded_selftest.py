"""CPU tier: the C++ sharded executor (csrc/sharded.cu) interpreted on host arrays against the oracle -- epoch plan,
per-shard tile plans (whole and slab-restricted passes through the tile engine's host interpreter), the staged exchange
pipeline (pushes into staging slots, unstage, riding passes, buffer reuse in issue order), the in-place fallback and
restore_order.  The GPU twins are tests/test_gpu_sharded_cpp.py and tools/check_sharded.py."""
import os

import numpy as np
import pytest

import opgen
from oracle.oracle import OracleQV

ENV = ("B200SV_SHARD_MIN_RUN_BITS", "B200SV_SHARD_SLAB_BITS", "B200SV_SHARD_STAGED", "B200SV_SHARD_MAX_RIDE")


def _run(n, world, ops, staging, env, dtype=np.complex128):
    from qiskit_aer_b200 import executor, sharded
    saved = {k: os.environ.get(k) for k in ENV}
    for k in ENV:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        psi = np.zeros(1 << n, dtype=dtype)
        psi[0] = 1.0
        stats = sharded.ShardedState.selftest(n, world, ops, staging, psi)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ref = OracleQV(n)
    executor.apply_ops(ref, ops)
    return psi, ref.vector(), stats


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("ride", [0, 2])
def test_staged_pipeline_on_host_arrays(world, ride):
    from qiskit_aer_b200 import circuits
    n = 14 + int(np.log2(world))            # 14 qubits per shard: 12-bit tiles, 2 free bits for slabs
    ops = circuits.quantum_volume(n, 5, seed=7 * world + ride)
    got, want, stats = _run(n, world, ops, staging=((16 << 14) * 3) // 8,
                            env={"B200SV_SHARD_MIN_RUN_BITS": "4", "B200SV_SHARD_SLAB_BITS": "2", "B200SV_SHARD_MAX_RIDE": str(ride)})
    assert stats["exchanges"] > 0 and stats["staged"] == stats["exchanges"] and stats["copies"] > 0
    assert (stats["overlapped_passes"] > 0) == (ride > 0)
    assert np.max(np.abs(got - want)) < 1e-12
    assert opgen.fidelity_gap(want, got) < 1e-10


def test_inplace_fallback_and_small_shards_on_host_arrays():
    from qiskit_aer_b200 import circuits
    # no staging area: in-place exchanges; 13 qubits per shard
    n, world = 15, 4
    ops = circuits.quantum_volume(n, 4, seed=2) + [("gate", "cx", [0, n - 1], []), ("gate", "cp", [n - 1, 3], [0.4]),
                                                   ("gate", "h", [n - 2], []), ("gate", "swap", [1, n - 1], [])]
    got, want, stats = _run(n, world, ops, staging=0, env={})
    assert stats["inplace"] == stats["exchanges"] > 0
    assert np.max(np.abs(got - want)) < 1e-12
    # shards below the tile size (8 qubits each): every gate is its own small pass
    n, world = 10, 4
    ops = circuits.quantum_volume(n, 3, seed=5)
    got, want, stats = _run(n, world, ops, staging=1 << 12, env={"B200SV_SHARD_MIN_RUN_BITS": "2"})
    assert np.max(np.abs(got - want)) < 1e-12


def test_single_precision_on_host_arrays():
    from qiskit_aer_b200 import circuits
    n, world = 16, 2
    ops = circuits.quantum_volume(n, 4, seed=11)
    got, want, stats = _run(n, world, ops, staging=1 << 17, env={"B200SV_SHARD_MIN_RUN_BITS": "4"}, dtype=np.complex64)
    assert stats["exchanges"] > 0
    assert opgen.fidelity_gap(want, got.astype(np.complex128)) < 1e-5
