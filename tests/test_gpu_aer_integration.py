"""The reference's own Controller / Fusion / executors / RngEngine, compiled with
QubitVectorB200 as the device="GPU" statevector (qiskit-aer_b200/aer/), against the same module's
device="CPU" path (the unmodified reference QubitVector): identical circuits, seeds and options."""
import numpy as np
import pytest

import opgen

pytestmark = pytest.mark.gpu


def _backend():
    import qiskit_aer_b200  # noqa: F401
    from qiskit_aer_b200 import aer_backend
    if not aer_backend.available():
        pytest.skip("Aer integration module not built (needs the reference tree at build time)")
    return aer_backend


def _counts(res, n):
    c = np.zeros(1 << n, dtype=np.int64)
    for k, v in res["data"]["counts"].items():
        c[int(k, 16)] = v
    return c


@pytest.mark.parametrize("n,fusion,fmax", [(10, False, 5), (12, True, 5), (14, True, 4), (15, True, 2)])
def test_qv_circuit_gpu_equals_cpu_through_the_reference_controller(n, fusion, fmax):
    from qiskit_aer_b200 import circuits
    be = _backend()
    assert "GPU" in be.load().aer_controller_execute().available_devices()
    ops = circuits.quantum_volume(n, 6, seed=n)
    paulis = opgen.random_paulis(n, n, 6, max_weight=4)
    kw = dict(shots=2000, seed=99, fusion=fusion, fusion_max_qubit=fmax, fusion_threshold=1, expvals=paulis,
              save_statevector=True)
    gpu = be.run_circuit(n, ops, device="GPU", **kw)
    cpu = be.run_circuit(n, ops, device="CPU", **kw)
    assert gpu["metadata"]["device"] == "GPU" and cpu["metadata"]["device"] == "CPU"
    assert opgen.fidelity_gap(np.asarray(cpu["data"]["sv"]), np.asarray(gpu["data"]["sv"])) < 1e-10
    for i in range(len(paulis)):
        assert abs(gpu["data"]["ev%d" % i] - cpu["data"]["ev%d" % i]) < 1e-10
    assert np.array_equal(_counts(gpu, n), _counts(cpu, n))  # same seed -> identical sampled counts


def test_named_gates_qft_and_measure_reset():
    """Gate table of Statevector::State (statevector_state.hpp:314-384) on the B200 vector, plus
    mid-circuit measure / reset (apply_measure :936-1014 -> probabilities + diagonal/permutation)."""
    from qiskit_aer_b200 import circuits
    be = _backend()
    n = 9
    ops = circuits.qft(n)
    extra = [("gate", "ccx", [0, 1, 2], []), ("gate", "cswap", [3, 4, 5], []), ("gate", "y", [6], []),
             ("gate", "cy", [6, 7], []), ("gate", "cz", [1, 8], []), ("gate", "t", [2], []),
             ("gate", "rx", [3], [0.3]), ("gate", "ry", [4], [1.1]), ("gate", "rz", [5], [0.7]),
             ("gate", "rxx", [0, 8], [0.4]), ("gate", "rzz", [1, 7], [0.9]), ("gate", "u", [2], [0.1, 0.2, 0.3]),
             ("gate", "sdg", [3], []), ("gate", "sx", [4], []), ("gate", "ecr", [5, 6], []),
             ("gate", "mcp", [0, 3, 6], [0.5]), ("gate", "cu", [7, 8], [0.3, 0.2, 0.1, 0.4])]
    ops = [("gate", "h", [q], []) for q in range(n)] + ops + extra
    kw = dict(shots=500, seed=5, fusion=False, save_statevector=True)
    gpu = be.run_circuit(n, ops, device="GPU", **kw)
    cpu = be.run_circuit(n, ops, device="CPU", **kw)
    assert opgen.fidelity_gap(np.asarray(cpu["data"]["sv"]), np.asarray(gpu["data"]["sv"])) < 1e-10
    assert np.array_equal(_counts(gpu, n), _counts(cpu, n))
    # mid-circuit measure + reset forces the per-shot path (no sampling optimisation): same seeds -> same counts
    ops2 = ops[:20] + [("measure", [0, 3], [0, 3]), ("reset", [1]), ("gate", "h", [1], [])] + ops[20:40]
    kw2 = dict(shots=60, seed=11, fusion=False)
    g2 = be.run_circuit(n, ops2, device="GPU", **kw2)
    c2 = be.run_circuit(n, ops2, device="CPU", **kw2)
    assert np.array_equal(_counts(g2, n), _counts(c2, n))


@pytest.mark.parametrize("blocking_qubits", [8, 10])
def test_cache_blocking_multi_chunk_path(blocking_qubits):
    """blocking_enable/blocking_qubits: ParallelStateExecutor + CacheBlocking + swap_chunk
    (parallel_state_executor.hpp:772-1336) drive apply_chunk_swap on B200 chunks; chunked == unchunked
    exactly, as test/terra/backends/aer_simulator/test_chunk.py:31-168 demands."""
    from qiskit_aer_b200 import circuits
    be = _backend()
    n = 12
    ops = circuits.quantum_volume(n, 5, seed=3) + circuits.qft(n)
    paulis = opgen.random_paulis(1, n, 4, max_weight=3)
    kw = dict(shots=1000, seed=21, fusion=True, fusion_max_qubit=3, fusion_threshold=1, expvals=paulis,
              save_statevector=True)
    chunked = be.run_circuit(n, ops, device="GPU", blocking_qubits=blocking_qubits, **kw)
    plain = be.run_circuit(n, ops, device="CPU", **kw)
    assert chunked["metadata"]["cacheblocking"]["enabled"]
    assert opgen.fidelity_gap(np.asarray(plain["data"]["sv"]), np.asarray(chunked["data"]["sv"])) < 1e-10
    for i in range(len(paulis)):
        assert abs(chunked["data"]["ev%d" % i] - plain["data"]["ev%d" % i]) < 1e-10
    assert np.array_equal(_counts(chunked, n), _counts(plain, n))


def test_single_precision_through_controller():
    from qiskit_aer_b200 import circuits
    be = _backend()
    n = 12
    ops = circuits.quantum_volume(n, 5, seed=8)
    kw = dict(shots=0, seed=3, fusion=True, fusion_threshold=1, save_statevector=True, precision="single",
              measure=False)
    gpu = be.run_circuit(n, ops, device="GPU", **kw)
    cpu = be.run_circuit(n, ops, device="CPU", **kw)
    assert opgen.fidelity_gap(np.asarray(cpu["data"]["sv"]), np.asarray(gpu["data"]["sv"])) < 1e-5
