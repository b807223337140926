"""Synthetic workload generators (no qiskit needed): the circuit families the reference
benchmarks with (test/benchmark/circuit_library_circuits.py:47-48,106-107) expressed as the
op tuples the reference lowers to (qiskit_aer/backends/aer_compiler.py:875-1050):

    ("unitary", qubits, U)        dense matrix, U[i, j] row/col as in numpy (qubits[0] = LSB)
    ("diagonal", qubits, d)
    ("gate", name, qubits, params)   h / cp / swap / x / cx / rz / sx ...
"""
import numpy as np


def haar_unitary(rng, dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diag(r)
    return q * (d / np.abs(d))


def quantum_volume(n, depth, seed):
    """QuantumVolume(n, depth): `depth` layers of floor(n/2) Haar-random SU(4) on a random pairing."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(depth):
        perm = rng.permutation(n)
        for i in range(n // 2):
            ops.append(("unitary", [int(perm[2 * i]), int(perm[2 * i + 1])], haar_unitary(rng, 4)))
    return ops


def qft(n, do_swaps=True):
    """QFT(n): for j = n-1..0: H(j); CP(pi/2^(j-k)) (k, j) for k < j; final bit-reversal swaps."""
    ops = []
    for j in range(n - 1, -1, -1):
        ops.append(("gate", "h", [j], []))
        for k in range(j - 1, -1, -1):
            ops.append(("gate", "cp", [k, j], [np.pi / (1 << (j - k))]))
    if do_swaps:
        for i in range(n // 2):
            ops.append(("gate", "swap", [i, n - 1 - i], []))
    return ops


def amplitudes_written(ops, n):
    """Circuit-level amplitude updates: what an un-fused pass per gate
    would write -- implementation independent, so it is the numerator of amplitude-updates/s
    for every engine regardless of how it fuses."""
    total = 0
    for op in ops:
        if op[0] == "gate" and op[1] == "cp":
            total += 1 << (n - 2)          # mcphase on 2 listed qubits
        elif op[0] == "gate" and op[1] == "swap":
            total += 1 << (n - 1)          # mcswap, 0 controls
        else:
            total += 1 << n
    return total


def random_noisy_circuit(n, depth, seed):
    """Config 5 workload: `depth` layers of {h, rz, sx} on every qubit followed by cx on a random pairing
    (the gate set of the reference noise benchmarks, test/benchmark/noise_20q.py / simulator_benchmark.py)."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(depth):
        for q in range(n):
            kind = int(rng.integers(3))
            if kind == 0:
                ops.append(("gate", "h", [q], []))
            elif kind == 1:
                ops.append(("gate", "rz", [q], [float(rng.uniform(0, 2 * np.pi))]))
            else:
                ops.append(("gate", "sx", [q], []))
        perm = rng.permutation(n)
        for i in range(n // 2):
            ops.append(("gate", "cx", [int(perm[2 * i]), int(perm[2 * i + 1])], []))
    return ops
