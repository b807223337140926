"""Host-side gate fusion for the B200 engine.

Plays the role of the reference's transpile pass (src/transpile/fusion.hpp:849, cost-based
`CostBasedFusion` :1002-1136, defaults max_qubit=5 / threshold=14 :762-763) but with a B200 cost
model: a k<=4 double-precision block is HBM bound (one pass = 2*16*2^n bytes) while k=5 leans on
the FP64 pipe (8 flop/byte), so blocks greedily grow up to `max_qubit` and never reorder
non-commuting gates.  Output ops are exactly what `apply_matrix` / `apply_diagonal_matrix`
consume.

Algorithm (list scheduling over open blocks): a gate may join block B iff no block emitted after
B touches any of its qubits; preference: the last block it depends on, then any later block that
stays within max_qubit, else a new block.
"""
import numpy as np

_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_SWAP = np.eye(4, dtype=np.complex128)[[0, 2, 1, 3]]


def gate_matrix(name, params):
    """Matrices of the named gates, numpy convention U[i, j], qubits[0] = least significant bit."""
    if name == "h":
        return _H
    if name == "x":
        return _X
    if name == "swap":
        return _SWAP
    if name == "cp":  # diag(1,1,1,e^{i theta})
        return np.diag([1, 1, 1, np.exp(1j * params[0])]).astype(np.complex128)
    if name == "cx":  # control qubits[0], target qubits[1]
        return np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)
    if name == "rz":
        return np.diag([np.exp(-0.5j * params[0]), np.exp(0.5j * params[0])])
    if name == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
    raise ValueError("unknown gate %s" % name)


def op_matrix(op):
    if op[0] == "unitary":
        return list(op[1]), np.asarray(op[2], dtype=np.complex128)
    if op[0] == "diagonal":
        return list(op[1]), np.diag(np.asarray(op[2], dtype=np.complex128))
    if op[0] == "gate":
        return list(op[2]), gate_matrix(op[1], op[3])
    raise ValueError(op[0])


def embed(U, qubits, block_qubits):
    """Lift U acting on `qubits` to the 2^k space of `block_qubits` (bit i of the index <-> block_qubits[i])."""
    k = len(block_qubits)
    pos = [block_qubits.index(q) for q in qubits]
    dim = 1 << k
    out = np.zeros((dim, dim), dtype=np.complex128)
    m = len(qubits)
    rest = [b for b in range(k) if b not in pos]
    for r in range(1 << len(rest)):
        base = 0
        for i, b in enumerate(rest):
            if (r >> i) & 1:
                base |= 1 << b
        idx = np.empty(1 << m, dtype=np.int64)
        for e in range(1 << m):
            v = base
            for i in range(m):
                if (e >> i) & 1:
                    v |= 1 << pos[i]
            idx[e] = v
        out[np.ix_(idx, idx)] = U
    return out


class Block:
    def __init__(self):
        self.qubits = []
        self.gates = []  # (qubits, U)
        self.diag = True

    def matrix(self):
        M = np.eye(1 << len(self.qubits), dtype=np.complex128)
        for q, U in self.gates:
            M = embed(U, q, self.qubits) @ M
        return M

    def is_diagonal(self):
        return all(np.count_nonzero(U - np.diag(np.diag(U))) == 0 for _, U in self.gates)


def _is_diag(U):
    return np.count_nonzero(U - np.diag(np.diag(U))) == 0


def fuse(ops, max_qubit=5, window=64, max_diag_qubit=10, max_table_qubit=16):
    """Returns a list of ("unitary", qubits, U) / ("diagonal", qubits, d) ops equivalent to `ops`.
    The block assignment and the block matrix products run in the library (b200sv_fuse_assign /
    b200sv_fuse_block_matrix, csrc/planner.cu); `_fuse_py` below is the same algorithm in numpy, kept as the
    cross-check of tests/test_host_logic.py."""
    import ctypes as C
    from . import capi
    lib = capi.lib()
    mats = [op_matrix(op) for op in ops]
    n_ops = len(mats)
    if n_ops == 0:
        return []
    off = np.zeros(n_ops + 1, dtype=np.int32)
    for i, (q, _) in enumerate(mats):
        off[i + 1] = off[i] + len(q)
    qs = np.asarray([x for q, _ in mats for x in q], dtype=np.int32)
    isd = np.asarray([1 if _is_diag(U) else 0 for _, U in mats], dtype=np.uint8)
    blk = np.zeros(n_ops, dtype=np.int32)
    nb = C.c_int(0)
    ip = C.POINTER(C.c_int)
    capi.check(lib.b200sv_fuse_assign(n_ops, off.ctypes.data_as(ip), qs.ctypes.data_as(ip),
                                      isd.ctypes.data_as(C.POINTER(C.c_uint8)), int(max_qubit), int(window),
                                      int(max_diag_qubit), blk.ctypes.data_as(ip), C.byref(nb)))
    members = [[] for _ in range(nb.value)]
    for i in range(n_ops):
        members[blk[i]].append(i)
    out = []
    for ms in members:
        bq = []
        for i in ms:
            for x in mats[i][0]:
                if x not in bq:
                    bq.append(x)
        diag = all(isd[i] for i in ms)
        if diag and len(bq) > max_table_qubit:
            # wider than a 2^k table can be: keep the gate list, the engine applies the whole layer in one pass
            # (b200sv_apply_diagonal_layer); gates on more than two qubits do not occur here (they were tables already)
            if all(len(mats[i][0]) <= 2 and np.all(np.diag(mats[i][1]) != 0) for i in ms):
                out.append(("diag_layer", [(list(mats[i][0]), np.diag(mats[i][1]).copy()) for i in ms]))
                continue
        goff = np.zeros(len(ms) + 1, dtype=np.int32)
        moff = np.zeros(len(ms), dtype=np.int64)
        flat, gq, pos = [], [], 0
        for j, i in enumerate(ms):
            q, U = mats[i]
            goff[j + 1] = goff[j] + len(q)
            gq += q
            moff[j] = pos
            U = np.ascontiguousarray(U, dtype=np.complex128)
            flat.append(U.reshape(-1))
            pos += U.size
        gm = np.concatenate(flat)
        k = len(bq)
        res = np.empty((1 << k) if diag else (1 << k, 1 << k), dtype=np.complex128)
        capi.check(lib.b200sv_fuse_block_matrix(
            k, np.asarray(bq, dtype=np.int32).ctypes.data_as(ip), len(ms), goff.ctypes.data_as(ip),
            np.asarray(gq, dtype=np.int32).ctypes.data_as(ip), moff.ctypes.data_as(C.POINTER(C.c_int64)),
            gm.ctypes.data_as(C.POINTER(C.c_double)), 1 if diag else 0, res.ctypes.data_as(C.POINTER(C.c_double))))
        out.append(("diagonal", list(bq), res) if diag else ("unitary", list(bq), res))
    return out


def _fuse_py(ops, max_qubit=5, window=64, max_diag_qubit=10):
    """Reference implementation of fuse() in numpy (same algorithm).

    Commutation aware: diagonal gates commute with each other, so a diagonal gate only depends on the
    last NON-diagonal block that shares a qubit with it, and purely diagonal blocks may grow to
    `max_diag_qubit` qubits (the streaming diagonal kernel keeps a 2^10 table in shared memory)."""
    blocks = []
    for op in ops:
        q, U = op_matrix(op)
        qs = set(q)
        gdiag = _is_diag(U)
        lo = max(0, len(blocks) - window)
        last_dep = -1
        for i in range(len(blocks) - 1, lo - 1, -1):
            if qs & set(blocks[i].qubits) and not (gdiag and blocks[i].diag):
                last_dep = i
                break
        if lo > 0 and last_dep < 0:
            last_dep = lo - 1  # cannot prove independence from blocks outside the window
        target = None
        cands = ([last_dep] if last_dep >= lo else []) + list(range(max(last_dep + 1, lo), len(blocks)))
        for i in cands:
            b = blocks[i]
            union = len(set(b.qubits) | qs)
            if b.diag and gdiag:
                ok = union <= max(max_diag_qubit, max_qubit)
            else:
                ok = union <= max_qubit
                # a non-diagonal gate may only join a later diagonal block if it overlaps nothing there
                # that is ordered before it -- blocks after last_dep never overlap non-commutingly, fine.
            if ok:
                target = b
                break
        if target is None:
            target = Block()
            target.diag = True
            blocks.append(target)
        for x in q:
            if x not in target.qubits:
                target.qubits.append(x)
        target.gates.append((q, U))
        target.diag = target.diag and gdiag
    out = []
    for b in blocks:
        if b.diag:
            d = np.ones(1 << len(b.qubits), dtype=np.complex128)
            idx = np.arange(1 << len(b.qubits))
            for gq, U in b.gates:  # product of diagonals without building 2^k x 2^k matrices
                sub = np.zeros_like(idx)
                for i, x in enumerate(gq):
                    sub |= ((idx >> b.qubits.index(x)) & 1) << i
                d *= np.diag(U)[sub]
            out.append(("diagonal", list(b.qubits), d))
        else:
            out.append(("unitary", list(b.qubits), b.matrix()))
    return out
