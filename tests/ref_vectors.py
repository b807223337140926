"""Runs the reference's own held vectors (tests/golden/ref_vectors.json, lifted from
/root/reference/test/terra/reference/ref_*.py by tests/golden/make_ref_vectors.py) through the Aer controller module
on a given device and checks counts / statevectors the way the reference's tests do
(test/terra/backends/aer_simulator/test_*.py: deterministic counts exactly, non-deterministic ones within
0.05 * shots, statevectors to 1e-10)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    return json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))["cases"]


def _ops(raw):
    out = []
    for op in raw:
        if op[0] == "gate":
            out.append(("gate", op[1], op[2], op[3]))
        elif op[0] == "unitary":
            out.append(("unitary", op[1], np.array([[complex(z[0], z[1]) for z in row] for row in op[2]])))
        elif op[0] == "measure":
            out.append(("measure", op[1], op[2]))
        elif op[0] == "reset":
            out.append(("reset", op[1]))
    return out


def check_case(be, case, device, **kw):
    n = case["num_qubits"]
    tag = "%s.%s[%d]" % (case["file"], case["function"], case["index"])
    if "counts" in case:
        shots = case["shots"]
        r = be.run_circuit(n, _ops(case["ops"]), device=device, shots=shots, seed=11, measure=False,
                           num_memory=max(case["num_clbits"], 1), **kw)
        got = {k: float(v) for k, v in r["data"].get("counts", {}).items()}
        want = case["counts"]
        if case["deterministic"] and len([v for v in want.values() if v]) == 1:
            assert got == {k: v for k, v in want.items() if v}, (tag, got, want)
        else:
            for k in set(got) | set(want):
                assert abs(got.get(k, 0.0) - want.get(k, 0.0)) <= 0.05 * shots, (tag, k, got, want)
    if "statevector" in case:
        r = be.run_circuit(n, _ops(case["ops_no_measure"]), device=device, shots=1, seed=11, measure=False,
                           save_statevector=True, **kw)
        got = np.asarray(r["data"]["sv"])
        want = np.array([complex(z[0], z[1]) for z in case["statevector"]])
        assert got.shape == want.shape and np.max(np.abs(got - want)) < 1e-10, (tag, got, want)
