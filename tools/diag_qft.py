import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qiskit_aer_b200
from qiskit_aer_b200 import aer_backend, circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ops = circuits.qft(n)
for fusion in (False, True, False):
    kw = dict(device="GPU", shots=1024, seed=1234, fusion=fusion, fusion_max_qubit=5, fusion_threshold=14, expvals=[([0, 1, n - 1], "ZXY")])
    for i in range(5):
        t0 = time.perf_counter()
        r = aer_backend.run_circuit(n, ops, **kw)
        md = r["metadata"]
        print("fusion", fusion, "run", i, "wall %.3f" % (time.perf_counter() - t0), "time_taken %.3f" % r["time_taken"],
              {k: (round(v, 3) if isinstance(v, float) else v) for k, v in md.items() if "time" in k or k in ("parallel_state_update",)}, flush=True)
