"""Phase breakdown of tile_pass_kernel (needs qiskit-aer_b200/libb200sv_prof.so built with -DB200SV_TILE_PROFILE)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qiskit_aer_b200 as q
from qiskit_aer_b200 import capi, circuits
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libb200sv_prof.so")
capi._lib = None
lib = capi.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
qv = q.QubitVectorB200(n)
rng = np.random.default_rng(0)
for g in (1, 2, 4, 6, 8):
    # g disjoint 2-qubit gates on high qubits -> one pass
    gates = [([n - 1 - 2 * i, n - 2 - 2 * i], circuits.haar_unitary(rng, 4).reshape(-1, order="F")) for i in range(g)]
    qv.apply_gate_sequence(gates)
    qv.synchronize()
    out = (ctypes.c_ulonglong * 4)()
    lib.b200sv_tile_profile(out, 1)
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    passes = qv.apply_gate_sequence(gates)
    qv.synchronize(); e1.record(); torch.cuda.synchronize()
    lib.b200sv_tile_profile(out, 1)
    tiles = out[3]
    print("gates=%d passes=%d tiles=%d  per tile cycles: load-wait %.0f  rounds %.0f  store %.0f   total %.0f" %
          (g, passes, tiles, out[0] / tiles, out[1] / tiles, out[2] / tiles, (out[0] + out[1] + out[2]) / tiles))
