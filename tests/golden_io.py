"""Decoder for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dec(x):
    if isinstance(x, dict):
        if "c" in x:
            a = np.asarray(x["c"], dtype=np.float64)
            return a[..., 0] + 1j * a[..., 1]
        if "z" in x:
            return complex(x["z"][0], x["z"][1])
    if isinstance(x, list):
        return [dec(v) for v in x]
    return x


def load_kernel(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(str(z["meta"]))
    ops = [(o[0], tuple(dec(o[1]))) for o in meta["ops"]]
    for i, (m, a) in enumerate(ops):  # permutation pairs back to tuples
        if m == "apply_permutation_matrix":
            ops[i] = (m, (a[0], [tuple(p) for p in a[1]]))
    paulis = [(p[0], p[1]) for p in meta["paulis"]]
    return z, meta, ops, paulis


def load_circuit(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(str(z["meta"]))
    ops = [("unitary", q, u) for q, u in zip(meta["qubits"], z["unitaries"])]
    paulis = [(p[0], p[1]) for p in meta["paulis"]]
    return z, meta, ops, paulis
