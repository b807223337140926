"""Single-precision gate queue: QV-n depth 10 through the float tile passes vs one streaming pass per gate
(B200SV_TILE_F32=0).  python tools/f32_tile_check.py [n]"""
import json, os, subprocess, sys
CHILD = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
import qiskit_aer_b200 as q
from qiskit_aer_b200 import circuits, executor
n = int(sys.argv[1])
qv = q.QubitVectorB200(n, np.complex64); qv.initialize()
ops = circuits.quantum_volume(n, 10, 1234)
st = {}
executor.apply_ops_queued(qv, ops, st); qv.synchronize()
t0 = time.perf_counter()
for _ in range(2): executor.apply_ops_queued(qv, ops)
qv.synchronize()
print(json.dumps({"qv_ms": round((time.perf_counter() - t0) / 2 * 1e3, 2), "passes": st.get("passes"), "norm": float(qv.norm()),
                  "ev": float(qv.expval_pauli([0, 1, n - 1], "ZXY"))}))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "30"
for env in ({}, {"B200SV_TILE_F32": "0"}):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD, n], env=e, capture_output=True, text=True)
    print(json.dumps({"env": env, "n": int(n)}), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
