"""QFT-n through the gate queue (b200sv_apply_gate_sequence: diagonal-layer split + tile passes) for several sizes of
the dense blocks between the wide diagonal layers (B200SV_LAYER_DENSE_QUBITS; read once per process -> one child each)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r"""
import os, sys, time
sys.path.insert(0, %r)
import numpy as np
import qiskit_aer_b200 as q
from qiskit_aer_b200 import circuits, executor, fusion
n = int(sys.argv[1])
ops = [("unitary", list(o[2]), fusion.gate_matrix(o[1], o[3])) if o[0] == "gate" and len(o[2]) <= 2 else o for o in circuits.qft(n)]
qv = q.QubitVectorB200(n)
for rep in range(4):
    qv.initialize(); qv.synchronize()
    st = {}
    t0 = time.perf_counter()
    executor.apply_ops_queued(qv, ops, stats=st)
    qv.synchronize()
    dt = time.perf_counter() - t0
print("dense<=%%s  %%.1f ms  stats %%s" %% (os.environ.get("B200SV_LAYER_DENSE_QUBITS"), dt * 1e3, st), flush=True)
if n <= 24:
    from oracle.oracle import OracleQV
    ref = OracleQV(n); executor.apply_ops(ref, ops)
    print("   max |diff| vs oracle %%.2e" %% np.max(np.abs(qv.vector() - ref.vector())))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "30"
for k in sys.argv[2:] or ["4", "5", "6", "7", "8"]:
    env = dict(os.environ, B200SV_LAYER_DENSE_QUBITS=k)
    subprocess.run([sys.executable, "-c", CHILD, n], env=env, timeout=300)
