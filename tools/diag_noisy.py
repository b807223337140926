import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qiskit_aer_b200
from qiskit_aer_b200 import aer_backend, circuits, noise
n, depth, shots = 20, 20, int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ops = circuits.random_noisy_circuit(n, depth, seed=1)
obs = [([0, 1], "ZZ"), ([3], "X"), ([5, 9, 11], "XYZ")]
nm = noise.noise_model_dict(1e-3, 1e-2)
kw = dict(seed=3, fusion=False, noise_model=nm, expvals=obs, batched_shots_gpu=True, batched_shots_gpu_max_qubits=20)
aer_backend.run_circuit(n, ops, device="GPU", shots=64, **kw)
t0 = time.perf_counter()
r = aer_backend.run_circuit(n, ops, device="GPU", shots=shots, **kw)
dt = time.perf_counter() - t0
print("VIRTUAL_GPUS", os.environ.get("B200SV_VIRTUAL_GPUS"), "shots/s %.0f" % (shots / dt), "time_taken %.2f" % r["time_taken"],
      {k: v for k, v in r["metadata"].items() if "batch" in k or "parallel" in k or "gpu" in k.lower()})
