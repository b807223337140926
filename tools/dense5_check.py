"""k = 5 dense blocks: DMMA kernel vs the register kernel (B200SV_DENSE5_DMMA=0).  python tools/dense5_check.py [n]"""
import json, os, subprocess, sys
CHILD = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
import opgen
import qiskit_aer_b200 as q
n = int(sys.argv[1])
rng = np.random.default_rng(3)
qv = q.QubitVectorB200(n); qv.initialize()
out = {}
for name, qs in (("high", [n - 1, n - 3, n - 5, n - 7, n - 9]), ("mixed", [3, 11, n - 2, 17, 8]), ("low", [0, 1, 2, 3, 4]), ("low2", [1, 4, 6, 2, n - 1])):
    U = opgen.colmajor(opgen.haar_unitary(rng, 32))
    for _ in range(2): qv.apply_matrix(qs, U)
    qv.synchronize(); t0 = time.perf_counter()
    for _ in range(5): qv.apply_matrix(qs, U)
    qv.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    out[name] = {"ms": round(ms, 3), "TBps": round(2 * 16 * 2.0 ** n / (ms / 1e3) / 1e12, 2), "fp64_TF": round(256 * 2.0 ** n / (ms / 1e3) / 1e12, 1)}
out["norm"] = float(qv.norm())
out["ev"] = float(qv.expval_pauli([0, 5, n - 1], "XZY"))
print(json.dumps(out))
""" % ((os.path.dirname(os.path.dirname(os.path.abspath(__file__))),) * 2)
n = sys.argv[1] if len(sys.argv) > 1 else "30"
for env in ({}, {"B200SV_DENSE5_DMMA": "0"}):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD, n], env=e, capture_output=True, text=True)
    print(json.dumps({"env": env, "n": int(n)}), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
