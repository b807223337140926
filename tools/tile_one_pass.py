"""Launch a few single tile passes carrying g disjoint 2-qubit gates (ncu target).  python tools/tile_one_pass.py n g [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qiskit_aer_b200 as q
from qiskit_aer_b200 import circuits
n, g = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
qv = q.QubitVectorB200(n); qv.initialize()
rng = np.random.default_rng(0)
gates = [([n - 1 - 2 * i, n - 2 - 2 * i], circuits.haar_unitary(rng, 4).reshape(-1, order="F")) for i in range(g)]
for _ in range(reps):
    qv.apply_gate_sequence(gates)
qv.synchronize()
print("norm", qv.norm())
