"""Depolarizing (Pauli-mixture) noise for the batched-shot executor and, as a plain dict in the
reference's `NoiseModel.to_dict()` schema (qiskit_aer/noise/noise_model.py:921-963,
qiskit_aer/noise/errors/quantum_error.py:308-335), for the reference Controller.

depolarizing_error(p, k) (qiskit_aer/noise/errors/standard_errors.py): every one of the 4^k Pauli strings
has probability p / 4^k, the identity additionally 1 - p."""
import itertools

import numpy as np

PAULI = "IXYZ"


def depolarizing_dict(gate_names, p, k):
    """One `qerror` entry acting on all-qubit gates named in `gate_names`."""
    instructions, probs = [], []
    for combo in itertools.product(range(4), repeat=k):
        circ = [{"name": PAULI[c].lower(), "qubits": [q]} for q, c in enumerate(combo) if c]
        if not circ:
            circ = [{"name": "id", "qubits": [0]}]
        instructions.append(circ)
        probs.append(p / 4 ** k + (1.0 - p if not any(combo) else 0.0))
    return {"type": "qerror", "id": "depol%d_%s" % (k, "_".join(gate_names)), "operations": list(gate_names),
            "instructions": instructions, "probabilities": probs}


def noise_model_dict(p1, p2, gates_1q=("h", "rz", "sx", "x", "u3"), gates_2q=("cx",)):
    """The reference benchmark's model: depolarizing p1 on 1-qubit gates, p2 on cx
    (test/benchmark/simulator_benchmark.py:114-118)."""
    return {"errors": [depolarizing_dict(gates_1q, p1, 1), depolarizing_dict(gates_2q, p2, 2)]}


def sample_pauli_codes(rng, noisy_ops, shots, p1, p2):
    """Per-shot Pauli codes for every noisy gate occurrence.

    noisy_ops: list of qubit tuples (1 or 2 qubits) in circuit order.  Returns (codes, slots) with
    codes[slot, shot] in 0..3 and slots[i] = tuple of slot indices (one per qubit) of occurrence i."""
    nslots = sum(len(q) for q in noisy_ops)
    codes = np.zeros((nslots, shots), dtype=np.uint8)
    slots, s = [], 0
    for qs in noisy_ops:
        k = len(qs)
        p = p1 if k == 1 else p2
        # index of the Pauli string: 0 = identity with prob 1 - p + p/4^k, each other p/4^k
        u = rng.random(shots)
        idx = np.zeros(shots, dtype=np.int64)
        hit = u < p * (1.0 - 1.0 / 4 ** k)
        idx[hit] = rng.integers(1, 4 ** k, size=int(hit.sum()))
        for j in range(k):
            codes[s + j] = (idx >> (2 * j)) & 3
        slots.append(tuple(range(s, s + k)))
        s += k
    return codes, slots
