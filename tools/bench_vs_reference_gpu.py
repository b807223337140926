"""Head-to-head on one B200: the reference's OWN GPU path (QubitVectorThrust, compiled unmodified for sm_100 by
`make -C oracle ref-gpu` -> oracle/_ref/gpu/controller_wrappers.so) vs the same reference Controller running on
the B200 engine (qiskit-aer_b200/aer/...).  Same circuits, seeds, options; each side in its own process (both are
builds of the same pybind module name).

    python tools/bench_vs_reference_gpu.py --qubits 30 31 32
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys, time, os, ctypes, glob, sysconfig
sys.path.insert(0, %(root)r)
import numpy as np
import qiskit_aer_b200
from qiskit_aer_b200 import circuits
side, n, workload, fusion, fmax = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4] == "1", int(sys.argv[5])
if side == "reference_gpu":
    libs = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
    for pat in ("libquadmath*", "libgfortran*", "libopenblas*"):
        for p in sorted(glob.glob(os.path.join(libs, pat))):
            ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    sys.path.insert(0, os.path.join(%(root)r, "oracle", "_ref", "gpu"))
    import controller_wrappers as cw
    from qiskit_aer_b200 import aer_backend
    aer_backend._cw = cw          # reuse the circuit/config driver with the reference-GPU module
from qiskit_aer_b200 import aer_backend
ops = circuits.qft(n) if workload == "qft" else circuits.quantum_volume(n, 10, 1234)
kw = dict(device="GPU", shots=1024, seed=1234, fusion=fusion, fusion_max_qubit=fmax, expvals=[([0, 1, n - 1], "ZXY")])
aer_backend.run_circuit(20, circuits.quantum_volume(20, 2, 1), device="GPU", shots=16, seed=1, fusion=fusion,
                        fusion_max_qubit=fmax)  # warm-up (context, modules)
best = None
for _ in range(2):
    t0 = time.perf_counter()
    r = aer_backend.run_circuit(n, ops, **kw)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print("RESULT " + json.dumps({"side": side, "n": n, "workload": workload, "fusion": fusion, "fusion_max_qubit": fmax,
                              "wall_s": best, "time_taken_s": float(r["time_taken"]), "ev": float(r["data"]["ev0"]),
                              "amp_updates_per_s": circuits.amplitudes_written(ops, n) / best}))
'''


def run(side, n, workload, fusion, fmax):
    code = CHILD % {"root": ROOT}
    p = subprocess.run([sys.executable, "-c", code, side, str(n), workload, "1" if fusion else "0", str(fmax)],
                       capture_output=True, text=True, timeout=1800)
    for line in p.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    return {"side": side, "n": n, "error": (p.stderr or p.stdout)[-400:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, nargs="+", default=[30])
    ap.add_argument("--workload", default="qv")
    a = ap.parse_args()
    out = []
    for n in a.qubits:
        out.append(run("reference_gpu", n, a.workload, True, 5))      # Aer defaults on its own Thrust kernels
        out.append(run("b200_engine", n, a.workload, True, 5))        # same options on the B200 engine
        out.append(run("b200_engine", n, a.workload, True, 4))        # fusion at the engine's FP64 ridge
        out.append(run("b200_engine", n, a.workload, False, 5))       # fusion off: adapter gate queue -> tile passes
        for r in out[-4:]:
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
