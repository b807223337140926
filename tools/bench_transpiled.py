"""A transpiled-looking circuit (three 1-qubit gates + cx or cz + rz per pair, brickwork of random pairs) through the
gate queue, with and without the queue-level absorption of 1-qubit gates.  python tools/bench_transpiled.py [n] [depth]"""
import json, os, subprocess, sys
CHILD = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
import opgen
import qiskit_aer_b200 as q
from qiskit_aer_b200 import executor
n, depth = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5)
CX = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)
CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)
ops = []
for layer in range(depth):
    perm = rng.permutation(n)
    for i in range(n // 2):
        a, b = int(perm[2 * i]), int(perm[2 * i + 1])
        for x in (a, b, a):
            ops.append(("unitary", [x], opgen.haar_unitary(rng, 2)))
        ops.append(("unitary", [a, b], CZ if (layer + i) %% 4 == 0 else CX))
        ops.append(("unitary", [b], np.diag(np.exp(1j * rng.uniform(0, 6.28, 2)))))
qv = q.QubitVectorB200(n); qv.initialize()
st = {}
executor.apply_ops_queued(qv, ops, st); qv.synchronize()
t0 = time.perf_counter()
for _ in range(2): executor.apply_ops_queued(qv, ops)
qv.synchronize()
print(json.dumps({"ops": len(ops), "ms": round((time.perf_counter() - t0) / 2 * 1e3, 2), "passes": st.get("passes"),
                  "norm": float(qv.norm()), "ev": float(qv.expval_pauli([0, 1, n - 1], "ZXY"))}))
""" % ((os.path.dirname(os.path.dirname(os.path.abspath(__file__))),) * 2)
n = sys.argv[1] if len(sys.argv) > 1 else "30"
depth = sys.argv[2] if len(sys.argv) > 2 else "10"
for env in ({}, {"B200SV_QUEUE_ABSORB": "0"}):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD, n, depth], env=e, capture_output=True, text=True)
    print(json.dumps({"env": env, "n": int(n)}), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
