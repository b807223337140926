import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
import qiskit_aer_b200
from qiskit_aer_b200 import aer_backend, circuits
n = int(sys.argv[1]); fusion = sys.argv[2] == "1"
ops = circuits.quantum_volume(n, 10, 1234)
kw = dict(device="GPU", shots=1024, seed=1234, fusion=fusion, fusion_max_qubit=5, expvals=[([0, 1, n - 1], "ZXY")])
aer_backend.run_circuit(20, circuits.quantum_volume(20, 2, 1), device="GPU", shots=16, seed=1, fusion=fusion, fusion_max_qubit=5)
for i in range(3):
    t0 = time.perf_counter()
    r = aer_backend.run_circuit(n, ops, **kw)
    print("run", i, "wall", time.perf_counter() - t0, "time_taken", r["time_taken"], flush=True)
