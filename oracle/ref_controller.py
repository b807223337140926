"""TEST INFRASTRUCTURE ONLY -- loader + thin driver for the reference's OWN
pybind module (qiskit_aer/backends/wrappers/bindings.cc compiled unmodified by
oracle/Makefile into oracle/_ref/controller_wrappers.so, CPU only).

This is the closest thing to ``AerSimulator(method="statevector",
device="CPU")`` that runs without qiskit: Controller + Fusion + the OpenMP/AVX2
QubitVector.  Used to (a) generate tests/golden fixtures, (b) time the CPU
baseline in ``bench.py --impl reference``.
"""
import ctypes
import glob
import os
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_cw = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "controller_wrappers.so"))


def load():
    global _cw
    if _cw is None:
        libs = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
        # the bundled OpenBLAS needs its libgfortran/libquadmath siblings resolved first
        for pat in ("libquadmath*", "libgfortran*", "libopenblas*"):
            for p in sorted(glob.glob(os.path.join(libs, pat))):
                ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        sys.path.insert(0, os.path.join(_HERE, "_ref"))
        try:
            import qiskit_aer  # noqa: F401
        except ImportError:  # two-class stub so that dict noise models parse (pybind_json.hpp:224-229)
            sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "qiskit-aer_b200", "aer", "pystub"))
        import controller_wrappers  # noqa
        _cw = controller_wrappers
    return _cw


def run_circuit(n, ops, shots=0, seed=1234, threads=0, fusion=True, fusion_max_qubit=5, fusion_threshold=14,
                expvals=(), save_statevector=False, measure=True, precision="double",
                blocking_qubits=None, noise_model=None):
    """ops: list of ("unitary", qubits, U) | ("diagonal", qubits, d) | ("gate", name, qubits, params).

    Mirrors the reference's own lowering conventions (qiskit_aer/backends/aer_compiler.py:875-1050).
    Returns the result dict of experiment 0.
    """
    import numpy as np
    cw = load()
    c = cw.AerCircuit()
    c.num_qubits = n
    c.num_memory = n if (shots and measure) else 0
    c.shots = max(shots, 1)
    c.seed = seed
    for op in ops:
        if op[0] == "unitary":
            c.unitary([int(q) for q in op[1]], np.ascontiguousarray(op[2], dtype=np.complex128), -1, None, "u")
        elif op[0] == "diagonal":
            c.diagonal([int(q) for q in op[1]], [complex(x) for x in op[2]], -1, "d")
        elif op[0] == "gate":
            c.gate(op[1], [int(q) for q in op[2]], [float(p) for p in op[3]], [], -1, None, op[1])
        else:
            raise ValueError(op[0])
    for i, (qs, p) in enumerate(expvals):
        c.save_expval([int(q) for q in qs], "save_expval", [p], [1.0], [0.0], "average", "ev%d" % i)
    if save_statevector:
        c.save_state(list(range(n)), "save_statevector", "single", "sv")
    if shots and measure:
        c.measure(list(range(n)), list(range(n)), [])
    cfg = cw.AerConfig()
    cfg.method = "statevector"
    cfg.device = "CPU"
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
    out = cw.aer_controller_execute().execute([c], noise_model, cfg)
    if not out.get("success", False):
        raise RuntimeError("reference controller failed: %s" % out.get("status"))
    return out["results"][0]
