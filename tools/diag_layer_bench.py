"""QFT-30 through the fusion front end with and without wide diagonal layers + the layer kernel alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qiskit_aer_b200 as q
from qiskit_aer_b200 import circuits, executor, fusion
n = 30
qv = q.QubitVectorB200(n)
ops = circuits.qft(n)
rng = np.random.default_rng(0)
def run(fused, label):
    for rep in range(3):
        qv.initialize(); qv.synchronize()
        t0 = time.perf_counter()
        for op in fused: executor.apply_op(qv, op)
        qv.synchronize()
        dt = time.perf_counter() - t0
    print(label, len(fused), "passes", "%.1f ms" % (dt * 1e3), flush=True)
run(fusion.fuse(ops, max_qubit=4, max_diag_qubit=16), "tables<=16   ")
for mt in (16, 12, 10):
    run(fusion.fuse(ops, max_qubit=4, max_diag_qubit=40, max_table_qubit=mt), "layers (table<=%d)" % mt)
gates = [([int(a), int(b)], np.exp(1j * rng.uniform(0, 6.28, 4))) for a in range(n) for b in range(a + 1, n)]
for rep in range(3):
    qv.synchronize(); t0 = time.perf_counter(); qv.apply_diagonal_layer(gates); qv.synchronize(); dt = time.perf_counter() - t0
print("layer kernel alone, %d gates: %.2f ms = %.0f GB/s" % (len(gates), dt * 1e3, 2 * 16 * 2.0 ** n / dt / 1e9))
