"""Loader + thin driver for the Aer integration build: the reference's own pybind module
(qiskit_aer/backends/wrappers/bindings.cc, unmodified) compiled with
`Statevector::State<QV::QubitVectorB200<...>>` as its device="GPU" statevector
(qiskit-aer_b200/aer/Makefile, prelude.hpp).  `AerSimulator`'s C++ stack -- Controller, Fusion,
CacheBlocking, executors, noise sampling, RngEngine -- runs unchanged on top of libb200sv.so.

    from qiskit_aer_b200 import aer_backend
    res = aer_backend.run_circuit(n, ops, device="GPU", shots=1024, seed=7, fusion_max_qubit=4)

`ops` use the reference's lowering conventions (qiskit_aer/backends/aer_compiler.py:875-1050):
("unitary", qubits, U) | ("diagonal", qubits, d) | ("gate", name, qubits, params).
"""
import ctypes
import glob
import os
import sys
import sysconfig

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_MOD_DIR = os.path.join(_HERE, "aer", "controller_wrappers_b200")
_cw = None


def available():
    return os.path.exists(os.path.join(_MOD_DIR, "controller_wrappers.so"))


def load():
    global _cw
    if _cw is None:
        if not available():
            raise RuntimeError("Aer integration module not built: run `make -C qiskit-aer_b200/aer` "
                               "(needs the reference tree)")
        libs = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
        for pat in ("libquadmath*", "libgfortran*", "libopenblas*"):
            for p in sorted(glob.glob(os.path.join(libs, pat))):
                ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        ctypes.CDLL(os.path.join(_HERE, "libb200sv.so"), mode=ctypes.RTLD_GLOBAL)
        try:
            import qiskit_aer  # noqa: F401  (a real installation wins)
        except ImportError:  # plain-dict noise models need two class names importable (pybind_json.hpp:224-229)
            sys.path.insert(0, os.path.join(_HERE, "aer", "pystub"))
        import importlib.util
        spec = importlib.util.spec_from_file_location("controller_wrappers",
                                                      os.path.join(_MOD_DIR, "controller_wrappers.so"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cw = mod
    return _cw


def build_circuit(cw, n, ops, shots, seed, expvals=(), save_statevector=False, measure=True, save_probs=None,
                  save_density_matrix=False, expval_subtype="average", num_memory=None):
    c = cw.AerCircuit()
    c.num_qubits = n
    c.num_memory = int(num_memory) if num_memory is not None else (n if (shots and measure) else 0)
    c.shots = max(shots, 1)
    c.seed = seed
    for op in ops:
        if op[0] == "unitary":
            c.unitary([int(q) for q in op[1]], np.ascontiguousarray(op[2], dtype=np.complex128), -1, None, "u")
        elif op[0] == "diagonal":
            c.diagonal([int(q) for q in op[1]], [complex(x) for x in op[2]], -1, "d")
        elif op[0] == "gate":
            c.gate(op[1], [int(q) for q in op[2]], [float(p) for p in op[3]], [], -1, None, op[1])
        elif op[0] == "measure":
            c.measure([int(q) for q in op[1]], [int(q) for q in op[2]], [])
        elif op[0] == "kraus":
            c.kraus([int(q) for q in op[1]], [np.ascontiguousarray(k, dtype=np.complex128) for k in op[2]], -1, None)
        elif op[0] == "reset":
            c.reset([int(q) for q in op[1]], -1)
        else:
            raise ValueError(op[0])
    for i, (qs, p) in enumerate(expvals):
        c.save_expval([int(q) for q in qs], "save_expval", [p], [1.0], [0.0], expval_subtype, "ev%d" % i)
    if save_probs is not None:
        c.save_state([int(q) for q in save_probs], "save_probabilities", "average", "probs")
    if save_statevector:
        c.save_state(list(range(n)), "save_statevector", "single", "sv")
    if save_density_matrix:
        c.save_state(list(range(n)), "save_density_matrix", "average", "dm")
    if shots and measure:
        c.measure(list(range(n)), list(range(n)), [])
    return c


def run_circuit(n, ops, device="GPU", shots=0, seed=1234, threads=0, fusion=True, fusion_max_qubit=5,
                fusion_threshold=14, precision="double", blocking_qubits=None, noise_model=None,
                batched_shots_gpu=False, batched_shots_gpu_max_qubits=16, method="statevector", target_gpus=None,
                chunk_swap_buffer_qubits=None, **circ_kw):
    """Runs through Controller::execute (src/controllers/aer_controller.hpp:458); returns experiment 0's dict."""
    cw = load()
    c = build_circuit(cw, n, ops, shots, seed, **circ_kw)
    cfg = cw.AerConfig()
    cfg.method = method
    cfg.device = device
    cfg.precision = precision
    cfg.n_qubits = n
    cfg.memory_slots = n
    cfg.seed_simulator = seed
    cfg.shots = max(shots, 1)
    cfg.fusion_enable = bool(fusion)
    cfg.fusion_max_qubit = fusion_max_qubit
    cfg.fusion_threshold = fusion_threshold
    if threads:
        cfg.max_parallel_threads = int(threads)
    if blocking_qubits is not None:
        cfg.blocking_enable = True
        cfg.blocking_qubits = int(blocking_qubits)
    if target_gpus is not None:
        cfg.target_gpus = [int(g) for g in target_gpus]
    if chunk_swap_buffer_qubits is not None:
        cfg.chunk_swap_buffer_qubits = int(chunk_swap_buffer_qubits)
    if batched_shots_gpu:
        cfg.batched_shots_gpu = True
        cfg.batched_shots_gpu_max_qubits = int(batched_shots_gpu_max_qubits)
    out = cw.aer_controller_execute().execute([c], noise_model, cfg)
    if not out.get("success", False):
        raise RuntimeError("Aer controller failed: %s" % out.get("status"))
    return out["results"][0]
