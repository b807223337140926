"""Runs op lists on a QubitVectorB200 -- the slice of Statevector::State::apply_op
(/root/reference/src/simulators/statevector/statevector_state.hpp:491-568, gate table :314-384)
that the benchmark circuits need: matrix / diagonal_matrix ops plus the named gates h, x, cx,
cp, swap, rz, sx mapped to the same QubitVector calls the reference uses."""
import numpy as np

from .fusion import gate_matrix


def colmajor(U):
    return np.asarray(U, dtype=np.complex128).reshape(-1, order="F")


def apply_op(qv, op):
    kind = op[0]
    if kind == "unitary":
        qv.apply_matrix(op[1], colmajor(op[2]))
    elif kind == "diagonal":
        qv.apply_diagonal_matrix(op[1], op[2])
    elif kind == "diag_layer":  # commuting diagonal 1-/2-qubit gates: one pass on the engine, gate by gate elsewhere
        if hasattr(qv, "apply_diagonal_layer"):
            qv.apply_diagonal_layer(op[1])
        else:
            for q, d in op[1]:
                qv.apply_diagonal_matrix(q, d)
    elif kind == "gate":
        name, qubits, params = op[1], op[2], op[3]
        if name == "h":      # apply_mcu(u4(pi/2,0,pi,0)), statevector_state.hpp:809-811
            qv.apply_mcu(qubits, colmajor(gate_matrix("h", [])))
        elif name in ("x", "cx", "ccx", "mcx"):
            qv.apply_mcx(qubits)
        elif name in ("y", "cy"):
            qv.apply_mcy(qubits)
        elif name in ("z", "cz", "ccz"):
            qv.apply_mcphase(qubits, -1.0)
        elif name == "cp":   # apply_mcphase, statevector_state.hpp:769-772
            qv.apply_mcphase(qubits, np.exp(1j * params[0]))
        elif name in ("swap", "cswap"):
            qv.apply_mcswap(qubits)
        elif name == "rz":
            qv.apply_diagonal_matrix(qubits, np.diag(gate_matrix("rz", params)))
        elif name == "sx":
            qv.apply_mcu(qubits, colmajor(gate_matrix("sx", [])))
        else:
            raise ValueError("unsupported gate %s" % name)
    else:
        raise ValueError(kind)


def apply_ops(qv, ops):
    for op in ops:
        apply_op(qv, op)


def op_h2d_bytes(op):
    """Bytes of gate data that cross the ABI from host memory for this op."""
    if op[0] == "unitary":
        return 16 * np.asarray(op[2]).size
    if op[0] == "diagonal":
        return 16 * len(op[2])
    if op[0] == "diag_layer":
        return 64 * len(op[1])
    return 64


def _small_dense(op):
    """(qubits, column-major matrix) if `op` is a dense 1- or 2-qubit gate the tile engine can queue."""
    if op[0] == "unitary" and len(op[1]) <= 2:
        return list(op[1]), colmajor(op[2])
    if op[0] == "gate" and op[1] in ("h", "sx", "x") and len(op[2]) == 1:
        return list(op[2]), colmajor(gate_matrix(op[1], op[3]))
    return None


def apply_ops_queued(qv, ops, stats=None, special=None):
    """Like apply_ops, but consecutive dense 1-/2-qubit gates are queued and flushed through
    b200sv_apply_gate_sequence (tile-blocked multi-gate passes) -- the B200 engine's analogue of the
    reference's blocked-gate queue (qubitvector_thrust.hpp:1102-1111,1511-1512).  `stats`, if given,
    accumulates {"passes": n, "launches": n}.  `special` maps extra op kinds (e.g. the sharded runner's
    "swap") to handlers returning the number of launches they issued."""
    queue = []

    def flush():
        if queue:
            passes = qv.apply_gate_sequence(queue)
            if stats is not None:
                stats["passes"] = stats.get("passes", 0) + passes
                stats["launches"] = stats.get("launches", 0) + passes
            queue.clear()

    for op in ops:
        g = _small_dense(op)
        if g is not None:
            queue.append(g)
            continue
        flush()
        if special is not None and op[0] in special:
            n_launch = special[op[0]](op)
        else:
            apply_op(qv, op)
            n_launch = 1
        if stats is not None:
            stats["launches"] = stats.get("launches", 0) + n_launch
            stats[op[0]] = stats.get(op[0], 0) + 1
    flush()
