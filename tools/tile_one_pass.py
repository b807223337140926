"""Launch a few single tile passes carrying g disjoint 2-qubit gates (ncu target).  python tools/tile_one_pass.py n g [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qiskit_aer_b200 as q
from qiskit_aer_b200 import circuits
n, g = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
layers = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # > 1: brickwork of `layers` layers on the same 2g qubits
qv = q.QubitVectorB200(n); qv.initialize()
rng = np.random.default_rng(0)
qs = [n - 1 - i for i in range(2 * g)]
gates = []
for l in range(layers):
    for i in range(g):
        a, b = qs[(2 * i + l) % (2 * g)], qs[(2 * i + 1 + l) % (2 * g)]
        gates.append(([a, b], circuits.haar_unitary(rng, 4).reshape(-1, order="F")))
import time
for _ in range(reps):
    qv.synchronize(); t0 = time.perf_counter()
    passes = qv.apply_gate_sequence(gates)
    qv.synchronize(); print("passes", passes, "ms", (time.perf_counter() - t0) * 1e3)
qv.synchronize()
print("norm", qv.norm())
