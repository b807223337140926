"""Lift the reference's OWN held vectors into tests/golden/ref_vectors.json.

The reference keeps its known-answer tests as Python functions under /root/reference/test/terra/reference/
(`<name>_circuits_deterministic()` builds qiskit circuits, `<name>_counts_*()` / `<name>_statevector_*()` return the
expected outputs; e.g. ref_non_clifford.py:418-457,948-991, ref_measure.py, ref_reset.py).  qiskit is not installed here,
so this script executes those functions against a small RECORDING stand-in for `qiskit` (QuantumRegister /
ClassicalRegister / QuantumCircuit that note every gate call) and writes circuits + targets as JSON.  Nothing of the
reference is copied: the fixture holds gate lists and numbers produced by running the reference's functions.

    python tests/golden/make_ref_vectors.py          # dev container only (needs /root/reference)

tests/test_golden_ref_vectors.py (CPU: through the reference CPU controller, pins the lifting itself) and
tests/test_gpu_ref_vectors.py (GPU: through the Aer adapter on the B200 engine) consume the fixture.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/test/terra/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")
FILES = ["ref_1q_clifford", "ref_2q_clifford", "ref_non_clifford", "ref_measure", "ref_reset", "ref_rotation",
         "ref_unitary_gate"]
SHOTS = 4000


class Unsupported(Exception):
    pass


class Bit:
    def __init__(self, reg, index):
        self.reg, self.index = reg, index


class Register:
    def __init__(self, size, name=None):
        self.size, self.name = size, name
        self.offset = None

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [Bit(self, j) for j in range(*i.indices(self.size))]
        return Bit(self, i if i >= 0 else self.size + i)

    def __iter__(self):
        return iter(Bit(self, j) for j in range(self.size))


class QuantumRegister(Register):
    pass


class ClassicalRegister(Register):
    pass


class QuantumCircuit:
    def __init__(self, *regs, **kw):
        self.ops, self.nq, self.nc = [], 0, 0
        self.qregs, self.cregs = [], []
        for r in regs:
            if isinstance(r, int):  # QuantumCircuit(n[, m])
                r = QuantumRegister(r) if not self.qregs else ClassicalRegister(r)
            if isinstance(r, QuantumRegister):
                r.offset_in = getattr(r, "offset_in", {})
                r.offset_in[id(self)] = self.nq
                self.nq += r.size
                self.qregs.append(r)
            else:
                r.offset_in = getattr(r, "offset_in", {})
                r.offset_in[id(self)] = self.nc
                self.nc += r.size
                self.cregs.append(r)

    # ---- argument resolution: a bit, an int, a register (broadcast) or a list of those
    def _bits(self, x, classical=False):
        if isinstance(x, Bit):
            return [x.reg.offset_in[id(self)] + x.index]
        if isinstance(x, Register):
            return [x.offset_in[id(self)] + j for j in range(x.size)]
        if isinstance(x, (int, np.integer)):
            return [int(x)]
        if isinstance(x, (list, tuple, range)):
            out = []
            for y in x:
                out += self._bits(y, classical)
            return out
        raise Unsupported("argument %r" % (x,))

    def _gate(self, name, params, *args):
        lists = [self._bits(a) for a in args]
        width = max(len(l) for l in lists)
        for i in range(width):  # qiskit broadcasting: registers expand element-wise, single bits repeat
            qs = [l[i] if len(l) > 1 else l[0] for l in lists]
            self.ops.append(["gate", name, qs, [float(p) for p in params]])

    def barrier(self, *a, **k):
        pass

    def delay(self, *a, **k):
        pass

    def h(self, q): self._gate("h", [], q)
    def x(self, q): self._gate("x", [], q)
    def y(self, q): self._gate("y", [], q)
    def z(self, q): self._gate("z", [], q)
    def s(self, q): self._gate("s", [], q)
    def sdg(self, q): self._gate("sdg", [], q)
    def t(self, q): self._gate("t", [], q)
    def tdg(self, q): self._gate("tdg", [], q)
    def id(self, q): self._gate("id", [], q)
    def sx(self, q): self._gate("sx", [], q)
    def rx(self, th, q): self._gate("rx", [th], q)
    def ry(self, th, q): self._gate("ry", [th], q)
    def rz(self, th, q): self._gate("rz", [th], q)
    def p(self, th, q): self._gate("p", [th], q)
    def _ctrl(self, name, params, ctrls, targets, ctrl_state):
        """controlled gate; controls whose ctrl_state bit is 0 are conjugated with X (qiskit's open controls)"""
        n = len(ctrls)
        cs = (1 << n) - 1 if ctrl_state is None else (int(ctrl_state, 2) if isinstance(ctrl_state, str) else int(ctrl_state))
        opened = [c for i, c in enumerate(ctrls) if not (cs >> i) & 1]
        for c in opened:
            self._gate("x", [], c)
        self._gate(name, params, *ctrls, *targets)
        for c in opened:
            self._gate("x", [], c)

    def cx(self, a, b, ctrl_state=None): self._ctrl("cx", [], [a], [b], ctrl_state)
    def cy(self, a, b, ctrl_state=None): self._ctrl("cy", [], [a], [b], ctrl_state)
    def cz(self, a, b, ctrl_state=None): self._ctrl("cz", [], [a], [b], ctrl_state)
    def swap(self, a, b): self._gate("swap", [], a, b)
    def ecr(self, a, b): self._gate("ecr", [], a, b)
    def cp(self, th, a, b, ctrl_state=None): self._ctrl("cp", [th], [a], [b], ctrl_state)
    def cu(self, th, ph, lam, gam, a, b, ctrl_state=None): self._ctrl("cu", [th, ph, lam, gam], [a], [b], ctrl_state)
    def ccx(self, a, b, c, ctrl_state=None): self._ctrl("ccx", [], [a, b], [c], ctrl_state)
    def cswap(self, a, b, c, ctrl_state=None): self._ctrl("cswap", [], [a], [b, c], ctrl_state)

    def unitary(self, mat, qubits, label=None):
        m = np.asarray(mat, dtype=np.complex128)
        self.ops.append(["unitary", self._bits(qubits), [[[float(z.real), float(z.imag)] for z in row] for row in m]])

    def measure(self, q, c):
        qs, cs = self._bits(q), self._bits(c, True)
        if len(qs) != len(cs):
            raise Unsupported("measure broadcast")
        self.ops.append(["measure", qs, cs])

    def reset(self, q):
        for b in self._bits(q):
            self.ops.append(["reset", [b]])

    def __getattr__(self, name):  # initialize, append, pauli, inverse, ...: circuits using them are skipped
        def missing(*a, **k):
            raise Unsupported(name)
        return missing


def install_shim():
    q = types.ModuleType("qiskit")
    q.QuantumRegister, q.ClassicalRegister, q.QuantumCircuit = QuantumRegister, ClassicalRegister, QuantumCircuit
    q.transpile = lambda c, *a, **k: c
    sys.modules["qiskit"] = q
    circ = types.ModuleType("qiskit.circuit")
    circ.Instruction = type("Instruction", (), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(Unsupported("Instruction"))})
    sys.modules["qiskit.circuit"] = circ
    lib = types.ModuleType("qiskit.circuit.library")
    sys.modules["qiskit.circuit.library"] = lib
    qi = types.ModuleType("qiskit.quantum_info")
    sys.modules["qiskit.quantum_info"] = qi
    rnd = types.ModuleType("qiskit.quantum_info.random")

    def random_unitary(dim, seed=None):
        from scipy.stats import unitary_group
        return unitary_group.rvs(dim, random_state=seed)
    rnd.random_unitary = random_unitary
    sys.modules["qiskit.quantum_info.random"] = rnd
    st = types.ModuleType("qiskit.quantum_info.states")
    sys.modules["qiskit.quantum_info.states"] = st
    # the reference's own helper package path (test.terra.utils.utils.list2dict)
    for name in ("test", "test.terra", "test.terra.utils", "test.terra.utils.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))

    def list2dict(counts_list, hex_counts=True):
        out = {}
        for i, c in enumerate(counts_list):
            if c:
                out[hex(i) if hex_counts else bin(i)[2:]] = c
        return out
    sys.modules["test.terra.utils.utils"].list2dict = list2dict


def load(name):
    spec = importlib.util.spec_from_file_location("refmod_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def call(fn, **prefer):
    import inspect
    kw = {k: v for k, v in prefer.items() if k in inspect.signature(fn).parameters}
    return fn(**kw)


def main():
    install_shim()
    cases, skipped = [], []
    for fname in FILES:
        try:
            mod = load(fname)
        except Exception as e:  # a file that needs more of qiskit than the stand-in offers
            skipped.append("%s: %s" % (fname, e))
            continue
        for attr in sorted(dir(mod)):
            if "_circuits_" not in attr:
                continue
            base, kind = attr.split("_circuits_")
            if kind not in ("deterministic", "nondeterministic"):
                continue
            counts_fn = getattr(mod, "%s_counts_%s" % (base, kind), None)
            sv_fn = getattr(mod, "%s_statevector_%s" % (base, kind), None)
            try:
                measured = call(getattr(mod, attr), final_measure=True, allow_sampling=True)
                counts = call(counts_fn, shots=SHOTS, hex_counts=True) if counts_fn else None
            except Unsupported as e:
                skipped.append("%s.%s: %s" % (fname, attr, e))
                continue
            svs, plain = None, None
            if sv_fn is not None:
                try:
                    plain = call(getattr(mod, attr), final_measure=False, allow_sampling=True)
                    svs = sv_fn()
                except Unsupported:
                    svs = None
            for i, c in enumerate(measured):
                case = {"file": fname, "function": attr, "index": i, "deterministic": kind == "deterministic",
                        "num_qubits": c.nq, "num_clbits": c.nc, "ops": c.ops}
                if counts is not None and len(counts) == len(measured):
                    case["shots"] = SHOTS
                    case["counts"] = {k: float(v) for k, v in counts[i].items()}
                if (svs is not None and plain is not None and len(svs) == len(plain) == len(measured)
                        and "final_measure" in getattr(mod, attr).__code__.co_varnames):
                    case["ops_no_measure"] = plain[i].ops
                    v = np.asarray(svs[i], dtype=np.complex128).reshape(-1)
                    case["statevector"] = [[float(z.real), float(z.imag)] for z in v]
                cases.append(case)
    json.dump({"source": "functions under /root/reference/test/terra/reference, executed by tests/golden/make_ref_vectors.py",
               "shots": SHOTS, "cases": cases, "skipped": skipped}, open(OUT, "w"))
    print("wrote %d cases from %d files to %s (%d skipped)" % (len(cases), len(FILES), OUT, len(skipped)))
    for s in skipped:
        print("  skipped:", s)


if __name__ == "__main__":
    main()
