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


@pytest.mark.parametrize("n", [9, 13])
def test_named_gates_qft_and_measure_reset(n):
    """Gate table of Statevector::State (statevector_state.hpp:314-384) on the B200 vector, plus
    mid-circuit measure / reset (apply_measure :936-1014 -> probabilities + diagonal/permutation).
    n = 13 reaches the tile engine: x / cx / cy / cz / cp / swap / cu are rewritten as (diagonal) 4x4 gates
    and ride on the adapter's gate queue."""
    from qiskit_aer_b200 import circuits
    be = _backend()
    ops = circuits.qft(n)
    extra = [("gate", "ccx", [0, 1, 2], []), ("gate", "cswap", [3, 4, 5], []), ("gate", "y", [6], []),
             ("gate", "cy", [6, 7], []), ("gate", "cz", [1, 8], []), ("gate", "t", [2], []),
             ("gate", "rx", [3], [0.3]), ("gate", "ry", [4], [1.1]), ("gate", "rz", [5], [0.7]),
             ("gate", "rxx", [0, 8], [0.4]), ("gate", "rzz", [1, 7], [0.9]), ("gate", "u", [2], [0.1, 0.2, 0.3]),
             ("gate", "sdg", [3], []), ("gate", "sx", [4], []), ("gate", "ecr", [5, 6], []),
             ("gate", "mcp", [0, 3, 6], [0.5]), ("gate", "cu", [7, 8], [0.3, 0.2, 0.1, 0.4]),
             ("gate", "cx", [n - 1, 0], []), ("gate", "swap", [n - 2, 1], []), ("gate", "cy", [2, n - 1], []),
             ("gate", "x", [n - 1], []), ("gate", "cp", [n - 1, 3], [0.8]), ("gate", "cu", [n - 2, n - 1], [0.7, 0.1, 0.9, 0.2]),
             ("gate", "p", [n - 1], [0.33]), ("gate", "z", [n - 2], []), ("gate", "s", [0], [])]
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


@pytest.mark.parametrize("virtual_gpus", [1, 2, 4])
@pytest.mark.parametrize("blocking_qubits,buffer_qubits", [(10, 2), (17, None)])
def test_cache_blocking_multi_chunk_swap_and_multi_device(virtual_gpus, blocking_qubits, buffer_qubits, monkeypatch):
    """apply_multi_chunk_swap (parallel_state_executor.hpp:1339-1552) -> apply_chunk_swap(chunk, dest_offset, src_offset,
    size) range swaps, and the chunks of the register spread over several GPUs (chunk i of C on target_gpus[i*G/C],
    chunk_manager.hpp:330-353): with B200SV_VIRTUAL_GPUS the placement / grouping / peer-swap code runs on one GPU,
    on a multi-GPU box the real devices are used.  Chunked == unchunked (test_chunk.py:31-168)."""
    import torch
    from qiskit_aer_b200 import circuits
    be = _backend()
    n = 13 if blocking_qubits == 10 else 20
    ndev = torch.cuda.device_count()
    if ndev >= 2 and virtual_gpus > 1:
        monkeypatch.delenv("B200SV_VIRTUAL_GPUS", raising=False)
        targets = list(range(min(ndev, virtual_gpus)))
    else:
        monkeypatch.setenv("B200SV_VIRTUAL_GPUS", str(virtual_gpus))
        targets = None
    ops = circuits.quantum_volume(n, 4, seed=3) + circuits.qft(n)[:3 * n]
    paulis = opgen.random_paulis(1, n, 4, max_weight=3)
    kw = dict(shots=1000, seed=21, fusion=True, fusion_max_qubit=3, fusion_threshold=1, expvals=paulis,
              save_statevector=n <= 16)
    chunked = be.run_circuit(n, ops, device="GPU", blocking_qubits=blocking_qubits, target_gpus=targets,
                             chunk_swap_buffer_qubits=buffer_qubits, **kw)
    plain = be.run_circuit(n, ops, device="CPU", **kw)
    assert chunked["metadata"]["cacheblocking"]["enabled"]
    if n <= 16:
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
    # >= 13 qubits, fusion off: 1-/2-qubit gates (named gates included) ride the adapter's queue into the float tile passes.
    # No diagonal gates here: the reference's float AVX2 diagonal kernel writes past its buffer (qv_avx2.cpp:1197-1203),
    # which would corrupt the heap of the CPU leg of this comparison.
    n = 14
    ops = circuits.quantum_volume(n, 4, seed=9) + [("gate", "h", [0], []), ("gate", "cx", [0, n - 1], []),
                                                   ("gate", "ry", [n - 1], [0.4]), ("gate", "swap", [2, n - 2], []),
                                                   ("gate", "x", [0], []), ("gate", "cy", [n - 1, 0], [])]
    kw = dict(shots=0, seed=3, fusion=False, save_statevector=True, precision="single", measure=False)
    gpu = be.run_circuit(n, ops, device="GPU", **kw)
    cpu = be.run_circuit(n, ops, device="CPU", **kw)
    assert opgen.fidelity_gap(np.asarray(cpu["data"]["sv"]), np.asarray(gpu["data"]["sv"])) < 1e-5


def _freq(res, n, shots):
    return _counts(res, n) / float(shots)


def test_batched_shots_gpu_option_noisy_pauli_circuit():
    """`batched_shots_gpu=True`: BatchShotsExecutor (batch_shots_executor.hpp:317-603) drives ONE container of
    `shots` states: gates in one launch for all shots, sampled Pauli noise as per-state codes on the tile
    passes, batched measure, host-side cregs.  The batched path draws differently from the per-shot path
    (rng.rand() thresholds, qubitvector_thrust.hpp:2296-2300), so parity is statistical."""
    from qiskit_aer_b200 import circuits, noise
    be = _backend()
    n, shots = 12, 3000
    # GHZ-like circuit + a few rotations: peaked output distribution
    ops = [("gate", "h", [0], [])] + [("gate", "cx", [q, q + 1], []) for q in range(n - 1)]
    ops += [("gate", "rz", [q], [0.3 * q]) for q in range(n)] + [("gate", "sx", [n - 1], []), ("gate", "sx", [n - 1], [])]
    nm = noise.noise_model_dict(0.01, 0.03)
    obs = [([0, 1], "ZZ"), ([0, n - 2], "ZZ"), ([3], "Z")]
    kw = dict(shots=shots, seed=5, fusion=False, noise_model=nm, expvals=obs)
    bat = be.run_circuit(n, ops, device="GPU", batched_shots_gpu=True, **kw)
    ref = be.run_circuit(n, ops, device="CPU", **kw)
    assert bat["metadata"].get("batched_shots_optimization") is True
    fb, fr = _freq(bat, n, shots), _freq(ref, n, shots)
    # total variation distance between two 3000-shot samples of the same distribution stays small
    assert 0.5 * np.abs(fb - fr).sum() < 0.12
    for i in range(len(obs)):
        assert abs(bat["data"]["ev%d" % i] - ref["data"]["ev%d" % i]) < 0.08
    ideal = be.run_circuit(n, ops, device="CPU", shots=shots, seed=5, fusion=False, expvals=obs)
    assert abs(ideal["data"]["ev0"] - bat["data"]["ev0"]) > 0.02  # the noise is really applied


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("virtual_gpus", [1, 3])
def test_batched_shots_mid_circuit_measure_reset_and_kraus(precision, virtual_gpus, monkeypatch):
    """Also: single-precision batches (aer_controller.hpp:642-646 batches QubitVectorThrust<float> too) and shot
    containers on several GPUs (one group per GPU, batch_shots_executor.hpp:421-435)."""
    from qiskit_aer_b200 import circuits
    be = _backend()
    monkeypatch.setenv("B200SV_VIRTUAL_GPUS", str(virtual_gpus))
    n, shots = 6, 4000
    g = 0.3  # amplitude damping Kraus pair on qubit 2
    K0 = np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex)
    K1 = np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)
    ops = [("gate", "h", [q], []) for q in range(n)] + [("gate", "cx", [0, 1], []), ("gate", "rx", [2], [1.1])]
    ops += [("measure", [0], [0]), ("reset", [1]), ("gate", "h", [1], []), ("kraus", [2], [K0, K1]),
            ("gate", "cx", [1, 3], []), ("gate", "ry", [4], [0.7])]
    kw = dict(shots=shots, seed=9, fusion=False, precision=precision)
    bat = be.run_circuit(n, ops, device="GPU", batched_shots_gpu=True, **kw)
    ref = be.run_circuit(n, ops, device="CPU", **dict(kw, precision="double"))
    assert bat["metadata"].get("batched_shots_optimization") is True
    fb, fr = _freq(bat, n, shots), _freq(ref, n, shots)
    assert 0.5 * np.abs(fb - fr).sum() < 0.1
    # marginal of the damped qubit: P(q2 = 1) agrees
    idx = np.arange(1 << n)
    p1b, p1r = fb[(idx >> 2) & 1 == 1].sum(), fr[(idx >> 2) & 1 == 1].sum()
    assert abs(p1b - p1r) < 0.04


def test_density_matrix_method_on_b200():
    """method="density_matrix", device="GPU": DensityMatrix::State<DensityMatrixB200<double>>
    (include/densitymatrix_b200.hpp) against the reference CPU DensityMatrix on the same noisy circuit:
    rho itself, Pauli expectation values, probabilities and fixed-seed counts."""
    from qiskit_aer_b200 import circuits, noise
    be = _backend()
    n, shots = 6, 2000
    ops = circuits.random_noisy_circuit(n, 3, seed=4)
    ops += [("gate", "ccx", [0, 1, 2], []), ("gate", "swap", [3, 4], []), ("gate", "y", [5], []),
            ("gate", "cz", [1, 4], []), ("gate", "cp", [2, 5], [0.7]), ("gate", "cy", [0, 3], []),
            ("reset", [2]), ("gate", "h", [2], []), ("gate", "t", [2], []), ("gate", "ecr", [1, 2], [])]
    nm = noise.noise_model_dict(0.02, 0.05)
    obs = [([0, 1], "ZZ"), ([2], "X"), ([3, 5], "XY"), ([4], "Z"), ([0, 2, 4], "YZX")]
    kw = dict(shots=shots, seed=13, fusion=True, fusion_threshold=1, noise_model=nm, expvals=obs,
              method="density_matrix", save_density_matrix=True)
    gpu = be.run_circuit(n, ops, device="GPU", **kw)
    cpu = be.run_circuit(n, ops, device="CPU", **kw)
    assert gpu["metadata"]["device"] == "GPU" and gpu["metadata"]["method"] == "density_matrix"
    rg, rc = np.asarray(gpu["data"]["dm"]), np.asarray(cpu["data"]["dm"])
    assert np.max(np.abs(rg - rc)) < 1e-12
    assert abs(np.trace(rg) - 1.0) < 1e-12
    for i in range(len(obs)):
        assert abs(gpu["data"]["ev%d" % i] - cpu["data"]["ev%d" % i]) < 1e-12
    assert np.array_equal(_counts(gpu, n), _counts(cpu, n))
    # fusion off: named-gate paths (apply_cnot / apply_x / apply_phase / apply_toffoli / apply_swap ...)
    kw["fusion"] = False
    gpu2 = be.run_circuit(n, ops, device="GPU", **kw)
    assert np.max(np.abs(np.asarray(gpu2["data"]["dm"]) - rc)) < 1e-12
