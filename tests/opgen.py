"""Seeded random op streams over the QubitVector method set, shared by the
oracle-pinning tests (CPU) and the CUDA parity tests (GPU).

An op is ``(method_name, args_tuple)`` with the reference's argument meaning
(qubitvector.hpp:225-294): controls first, target(s) last; column-major
matrices.  ``apply(qv, op)`` replays it on any object exposing those methods.
"""
import numpy as np


def haar_unitary(rng, dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diag(r)
    return q * (d / np.abs(d))


def random_state(rng, n, dtype=np.complex128):
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def colmajor(U):
    """Aer vectorises matrices column-major: mat[i + dim*j] = U[i, j]."""
    return np.asarray(U, dtype=np.complex128).reshape(-1, order="F")


def pick(rng, n, k):
    return [int(q) for q in rng.choice(n, size=k, replace=False)]


def random_op(rng, n, kinds=None, max_k=5):
    kinds = kinds or ["matrix", "diagonal", "mcx", "mcy", "mcswap", "mcphase", "mcu", "mcu_diag",
                      "pauli", "multiplexer", "permutation"]
    kind = kinds[int(rng.integers(len(kinds)))]
    if kind == "matrix":
        k = int(rng.integers(1, min(max_k, n) + 1))
        return ("apply_matrix", (pick(rng, n, k), colmajor(haar_unitary(rng, 1 << k))))
    if kind == "diagonal":
        k = int(rng.integers(1, min(max_k, n) + 1))
        d = np.exp(2j * np.pi * rng.random(1 << k))
        if rng.random() < 0.3:
            d[int(rng.integers(1 << k))] = 1.0
        return ("apply_diagonal_matrix", (pick(rng, n, k), d))
    if kind in ("mcx", "mcy"):
        k = int(rng.integers(1, min(4, n) + 1))
        return ("apply_" + kind, (pick(rng, n, k),))
    if kind == "mcswap":
        k = int(rng.integers(2, min(4, n) + 1))
        return ("apply_mcswap", (pick(rng, n, k),))
    if kind == "mcphase":
        k = int(rng.integers(1, min(4, n) + 1))
        return ("apply_mcphase", (pick(rng, n, k), complex(np.exp(2j * np.pi * rng.random()))))
    if kind == "mcu":
        k = int(rng.integers(1, min(4, n) + 1))
        return ("apply_mcu", (pick(rng, n, k), colmajor(haar_unitary(rng, 2))))
    if kind == "mcu_diag":
        k = int(rng.integers(1, min(4, n) + 1))
        d = np.exp(2j * np.pi * rng.random(2))
        if rng.random() < 0.5:
            d[0] = 1.0
        return ("apply_mcu", (pick(rng, n, k), colmajor(np.diag(d))))
    if kind == "pauli":
        k = int(rng.integers(1, min(6, n) + 1))
        p = "".join("IXYZ"[int(c)] for c in rng.integers(0, 4, size=k))
        return ("apply_pauli", (pick(rng, n, k), p, complex(np.exp(2j * np.pi * rng.random()))))
    if kind == "multiplexer":
        nc = int(rng.integers(1, min(2, n - 1) + 1))
        nt = int(rng.integers(1, min(2, n - nc) + 1))
        qs = pick(rng, n, nc + nt)
        dim, cols = 1 << (nc + nt), 1 << nt
        mat = np.zeros((dim, cols), dtype=np.complex128)  # stacked blocks, column-major vectorised
        for b in range(1 << nc):
            mat[b * cols:(b + 1) * cols, :] = haar_unitary(rng, cols)
        return ("apply_multiplexer", (qs[:nc], qs[nc:], mat.reshape(-1, order="F")))
    if kind == "permutation":
        k = int(rng.integers(1, min(4, n) + 1))
        dim = 1 << k
        npairs = int(rng.integers(1, dim))
        pairs = [(int(a), int(b)) for a, b in (rng.choice(dim, size=2, replace=False) for _ in range(npairs))]
        return ("apply_permutation_matrix", (pick(rng, n, k), pairs))
    raise ValueError(kind)


def random_ops(seed, n, count, **kw):
    rng = np.random.default_rng(seed)
    return [random_op(rng, n, **kw) for _ in range(count)]


def apply(qv, op):
    return getattr(qv, op[0])(*op[1])


def fidelity_gap(a, b):
    """|1 - <a|b>| as in the north-star tolerance."""
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    return abs(1.0 - np.vdot(a, b))


def random_paulis(seed, n, count, max_weight=5):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        k = int(rng.integers(1, min(max_weight, n) + 1))
        out.append((pick(rng, n, k), "".join("IXYZ"[int(c)] for c in rng.integers(0, 4, size=k))))
    return out
