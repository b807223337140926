import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import qiskit_aer_b200, opgen
from qiskit_aer_b200 import aer_backend as be, circuits
def cmp(n, ops, **kw):
    g = be.run_circuit(n, ops, device="GPU", shots=0, save_statevector=True, measure=False, **kw)
    c = be.run_circuit(n, ops, device="CPU", shots=0, save_statevector=True, measure=False, **{k:v for k,v in kw.items() if k!='blocking_qubits'})
    a,b=np.asarray(g["data"]["sv"]),np.asarray(c["data"]["sv"])
    return float(np.max(np.abs(a-b)))
n=12
print("qft blk10 nofusion", cmp(n,circuits.qft(n),fusion=False,blocking_qubits=10))
print("qft noswap blk10 nofusion", cmp(n,circuits.qft(n,do_swaps=False),fusion=False,blocking_qubits=10))
H=[("gate","h",[q],[]) for q in range(n)]+[("gate","rx",[q],[0.1*q+0.2]) for q in range(n)]
print("swap 0,11", cmp(n,H+[("gate","swap",[0,11],[])],fusion=False,blocking_qubits=10))
print("swap 10,11", cmp(n,H+[("gate","swap",[10,11],[])],fusion=False,blocking_qubits=10))
print("swap 3,4", cmp(n,H+[("gate","swap",[3,4],[])],fusion=False,blocking_qubits=10))
print("qv blk10 nofusion", cmp(n,circuits.quantum_volume(n,5,3),fusion=False,blocking_qubits=10))
print("qv blk10 fusion3", cmp(n,circuits.quantum_volume(n,5,3),fusion=True,fusion_max_qubit=3,fusion_threshold=1,blocking_qubits=10))
print("qv blk8 fusion3", cmp(n,circuits.quantum_volume(n,5,3),fusion=True,fusion_max_qubit=3,fusion_threshold=1,blocking_qubits=8))
print("qft blk10 fusion3", cmp(n,circuits.qft(n),fusion=True,fusion_max_qubit=3,fusion_threshold=1,blocking_qubits=10))
# measure/reset
def counts(res,n):
    c=np.zeros(1<<n,dtype=np.int64)
    for k,v in res["data"]["counts"].items(): c[int(k,16)]=v
    return c
n=5
ops=[("gate","h",[q],[]) for q in range(n)]+[("gate","rx",[q],[0.1*q+0.2]) for q in range(n)]
for tag,extra in (("measure",[("measure",[0,3],[0,3])]),("reset",[("reset",[1])]),("both",[("measure",[0,3],[0,3]),("reset",[1]),("gate","h",[1],[])])):
    g=be.run_circuit(n,ops+extra+ops,device="GPU",shots=40,seed=11,fusion=False)
    c=be.run_circuit(n,ops+extra+ops,device="CPU",shots=40,seed=11,fusion=False)
    print(tag, np.array_equal(counts(g,n),counts(c,n)), np.abs(counts(g,n)-counts(c,n)).sum())
