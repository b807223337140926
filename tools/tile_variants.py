"""A/B of tile_pass_kernel variants selected by environment knobs (one child process per setting, because the
library reads its knobs once).  Usage: python tools/tile_variants.py [n] -- prints ms per pass for single passes
carrying g disjoint 2-qubit gates and the time of a whole QV-n depth-10 circuit through the gate queue."""
import json, os, subprocess, sys, time

CHILD = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
import qiskit_aer_b200 as q
from qiskit_aer_b200 import circuits, executor
n = int(sys.argv[1])
qv = q.QubitVectorB200(n); qv.initialize()
rng = np.random.default_rng(0)
out = {}
for g in (2, 4, 6, 8):
    gates = [([n - 1 - 2 * i, n - 2 - 2 * i], circuits.haar_unitary(rng, 4).reshape(-1, order="F")) for i in range(g)]
    for _ in range(2): qv.apply_gate_sequence(gates)
    qv.synchronize(); t0 = time.perf_counter()
    reps = 6
    for _ in range(reps): passes = qv.apply_gate_sequence(gates)
    qv.synchronize(); out["g%%d" %% g] = round((time.perf_counter() - t0) / reps / passes * 1e3, 3)
ops = circuits.quantum_volume(n, 10, 1234)
st = {}
executor.apply_ops_queued(qv, ops, st); qv.synchronize()
t0 = time.perf_counter()
for _ in range(2): executor.apply_ops_queued(qv, ops)
qv.synchronize(); out["qv_ms"] = round((time.perf_counter() - t0) / 2 * 1e3, 2); out["passes"] = st.get("passes")
out["norm"] = qv.norm()
print(json.dumps(out))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    settings = [{}] + [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[2:]]
    for env in settings:
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", CHILD, str(n)], env=e, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
        print(json.dumps({"env": env, "n": n}), line, flush=True)

if __name__ == "__main__":
    main()
