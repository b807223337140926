"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the two CPU checkers.

* ``OracleQV``  : our plain-C restatement (oracle/sv_oracle.c), state held in a
                  numpy array.  ``cpu_baseline.kind == "port"``.
* ``RefQV``     : the UNMODIFIED reference ``AER::QV::QubitVector`` behind the
                  shim oracle/ref_harness.cpp (oracle/_ref/libaer_qv_ref.so).
                  ``cpu_baseline.kind == "reference"``.

Both expose the reference's QubitVector method names and argument meaning
(/root/reference/src/simulators/statevector/qubitvector.hpp:225-411): qubit
lists with controls first / target last, column-major matrices, Pauli strings
with ``pauli[N-1-i] <-> qubits[i]``.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def build(ref=True):
    """Compile the C restatement and (when the reference tree is present) _ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libsv_oracle.so"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _q(qubits):
    a = np.ascontiguousarray(qubits, dtype=np.uint64)
    return a, a.ctypes.data_as(_u64p), int(a.size)


def _m(mat):
    a = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128).reshape(-1))
    return a, a.ctypes.data_as(_f64p)


def pauli_masks(qubits, pauli):
    """x_mask, z_mask, num_y, x_max -- qubitvector.hpp:2236-2273."""
    n = len(qubits)
    x = z = ny = xm = 0
    for i, q in enumerate(qubits):
        p = pauli[n - 1 - i]
        if p == "I":
            continue
        if p in "XY":
            x |= 1 << q
            xm = max(xm, q)
        if p in "ZY":
            z |= 1 << q
        if p == "Y":
            ny += 1
        if p not in "XYZ":
            raise ValueError("Invalid Pauli")
    return x, z, ny, xm


class OracleQV:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "libsv_oracle.so")
            if not os.path.exists(path):
                build(ref=False)
            cls._lib = C.CDLL(path)
            for sfx in ("_f64", "_f32"):
                for name in ("svo_norm", "svo_norm_matrix", "svo_expval_pauli"):
                    getattr(cls._lib, name + sfx).restype = C.c_double
        return cls._lib

    def __init__(self, num_qubits, dtype=np.complex128, index_size=10):
        self.n = int(num_qubits)
        self.dtype = np.dtype(dtype)
        self.sfx = "_f64" if self.dtype == np.complex128 else "_f32"
        self.index_size = index_size
        self.psi = np.zeros(1 << self.n, dtype=self.dtype)
        self.psi[0] = 1

    def _f(self, name):
        return getattr(self.lib(), name + self.sfx)

    @property
    def _p(self):
        return C.c_void_p(self.psi.ctypes.data)

    def set_state(self, vec):
        self.psi = np.ascontiguousarray(vec, dtype=self.dtype).copy()

    def vector(self):
        return self.psi.copy()

    def apply_matrix(self, qubits, mat):
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        self._f("svo_apply_matrix")(self._p, self.n, qp, k, mp)

    def apply_diagonal_matrix(self, qubits, diag):
        qa, qp, k = _q(qubits)
        ma, mp = _m(diag)
        self._f("svo_apply_diagonal")(self._p, self.n, qp, k, mp)

    def apply_mcx(self, qubits):
        qa, qp, k = _q(qubits)
        self._f("svo_apply_mcx")(self._p, self.n, qp, k)

    def apply_mcy(self, qubits):
        qa, qp, k = _q(qubits)
        self._f("svo_apply_mcy")(self._p, self.n, qp, k)

    def apply_mcswap(self, qubits):
        qa, qp, k = _q(qubits)
        self._f("svo_apply_mcswap")(self._p, self.n, qp, k)

    def apply_mcphase(self, qubits, phase):
        qa, qp, k = _q(qubits)
        ph = complex(phase)
        self._f("svo_apply_mcphase")(self._p, self.n, qp, k, C.c_double(ph.real), C.c_double(ph.imag))

    def apply_mcu(self, qubits, mat):
        qa, qp, k = _q(qubits)
        if k == 1:  # qubitvector.hpp:1637,1676: N == 1 is a plain 1-qubit matrix / diagonal
            m = np.asarray(mat, dtype=np.complex128).reshape(-1)
            if m[1] == 0 and m[2] == 0 and m[0] != 1:
                return self.apply_diagonal_matrix(qubits, [m[0], m[3]])
            if not (m[1] == 0 and m[2] == 0):
                return self.apply_matrix(qubits, mat)
        ma, mp = _m(mat)
        self._f("svo_apply_mcu")(self._p, self.n, qp, k, mp)

    def apply_multiplexer(self, control_qubits, target_qubits, mat):
        ca, cp, nc = _q(control_qubits)
        ta, tp, nt = _q(target_qubits)
        ma, mp = _m(mat)
        self._f("svo_apply_multiplexer")(self._p, self.n, cp, nc, tp, nt, mp)

    def apply_permutation_matrix(self, qubits, pairs):
        qa, qp, k = _q(qubits)
        pa = np.ascontiguousarray(np.asarray(pairs, dtype=np.uint64).reshape(-1))
        self._f("svo_apply_permutation")(self._p, self.n, qp, k, pa.ctypes.data_as(_u64p), len(pa) // 2)

    def apply_pauli(self, qubits, pauli, coeff=1.0):
        x, z, ny, xm = pauli_masks(list(qubits), pauli)
        c = complex(coeff)
        self._f("svo_apply_pauli")(self._p, self.n, C.c_uint64(x), C.c_uint64(z), ny, xm,
                                   C.c_double(c.real), C.c_double(c.imag))

    def norm(self, qubits=None, mat=None):
        if qubits is None:
            return self._f("svo_norm")(self._p, self.n)
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        return self._f("svo_norm_matrix")(self._p, self.n, qp, k, mp)

    def probabilities(self, qubits=None):
        if qubits is None:
            qubits = list(range(self.n))
        qa, qp, k = _q(qubits)
        out = np.zeros(1 << k, dtype=np.float64)
        self._f("svo_probabilities")(self._p, self.n, qp, k, out.ctypes.data_as(_f64p))
        return out

    def sample_measure(self, rnds):
        r = np.ascontiguousarray(rnds, dtype=np.float64)
        out = np.zeros(r.size, dtype=np.uint64)
        self._f("svo_sample_measure")(self._p, self.n, self.index_size, r.ctypes.data_as(_f64p),
                                      C.c_int64(r.size), out.ctypes.data_as(_u64p))
        return out

    def expval_pauli(self, qubits, pauli, initial_phase=1.0):
        x, z, ny, xm = pauli_masks(list(qubits), pauli)
        c = complex(initial_phase)
        return self._f("svo_expval_pauli")(self._p, self.n, C.c_uint64(x), C.c_uint64(z), ny, xm,
                                           C.c_double(c.real), C.c_double(c.imag))


class RefQV:
    """The reference's own QubitVector<double|float> (unmodified), via ctypes."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(_HERE, "_ref", "libaer_qv_ref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            cls._lib = C.CDLL(os.path.join(_HERE, "_ref", "libaer_qv_ref.so"))
            for sfx in ("_f64", "_f32"):
                getattr(cls._lib, "refqv_create" + sfx).restype = C.c_void_p
                getattr(cls._lib, "refqv_data" + sfx).restype = C.c_void_p
                for name in ("refqv_norm", "refqv_norm_matrix", "refqv_expval_pauli"):
                    getattr(cls._lib, name + sfx).restype = C.c_double
        return cls._lib

    @classmethod
    def max_threads(cls):
        return int(cls.lib().refqv_max_threads())

    def __init__(self, num_qubits, dtype=np.complex128, threads=None, index_size=10):
        self.n = int(num_qubits)
        self.dtype = np.dtype(dtype)
        self.sfx = "_f64" if self.dtype == np.complex128 else "_f32"
        if threads is None:
            threads = self.max_threads()
        self.threads = threads
        self.h = C.c_void_p(self._f("refqv_create")(self.n, int(threads)))
        self._f("refqv_set_sample_index_size")(self.h, int(index_size))

    def __del__(self):
        try:
            if self.h:
                self._f("refqv_destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def _f(self, name):
        return getattr(self.lib(), name + self.sfx)

    def _view(self):
        ptr = self._f("refqv_data")(self.h)
        buf = (C.c_char * ((1 << self.n) * self.dtype.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=self.dtype)

    def set_state(self, vec):
        self._view()[:] = np.asarray(vec, dtype=self.dtype)

    def vector(self):
        return self._view().copy()

    def initialize(self):
        self._f("refqv_initialize")(self.h)

    def apply_matrix(self, qubits, mat):
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        self._f("refqv_apply_matrix")(self.h, qp, k, mp)

    def apply_diagonal_matrix(self, qubits, diag):
        qa, qp, k = _q(qubits)
        ma, mp = _m(diag)
        self._f("refqv_apply_diagonal")(self.h, qp, k, mp)

    def apply_mcx(self, qubits):
        qa, qp, k = _q(qubits)
        self._f("refqv_apply_mcx")(self.h, qp, k)

    def apply_mcy(self, qubits):
        qa, qp, k = _q(qubits)
        self._f("refqv_apply_mcy")(self.h, qp, k)

    def apply_mcswap(self, qubits):
        qa, qp, k = _q(qubits)
        self._f("refqv_apply_mcswap")(self.h, qp, k)

    def apply_mcphase(self, qubits, phase):
        qa, qp, k = _q(qubits)
        ph = complex(phase)
        self._f("refqv_apply_mcphase")(self.h, qp, k, C.c_double(ph.real), C.c_double(ph.imag))

    def apply_mcu(self, qubits, mat):
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        self._f("refqv_apply_mcu")(self.h, qp, k, mp)

    def apply_multiplexer(self, control_qubits, target_qubits, mat):
        ca, cp, nc = _q(control_qubits)
        ta, tp, nt = _q(target_qubits)
        ma, mp = _m(mat)
        self._f("refqv_apply_multiplexer")(self.h, cp, nc, tp, nt, mp)

    def apply_permutation_matrix(self, qubits, pairs):
        qa, qp, k = _q(qubits)
        pa = np.ascontiguousarray(np.asarray(pairs, dtype=np.uint64).reshape(-1))
        self._f("refqv_apply_permutation")(self.h, qp, k, pa.ctypes.data_as(_u64p), len(pa) // 2)

    def apply_pauli(self, qubits, pauli, coeff=1.0):
        qa, qp, k = _q(qubits)
        c = complex(coeff)
        self._f("refqv_apply_pauli")(self.h, qp, k, pauli.encode(), C.c_double(c.real), C.c_double(c.imag))

    def norm(self, qubits=None, mat=None):
        if qubits is None:
            return self._f("refqv_norm")(self.h)
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        return self._f("refqv_norm_matrix")(self.h, qp, k, mp)

    def probabilities(self, qubits=None):
        if qubits is None:
            qubits = list(range(self.n))
        qa, qp, k = _q(qubits)
        out = np.zeros(1 << k, dtype=np.float64)
        self._f("refqv_probabilities")(self.h, qp, k, out.ctypes.data_as(_f64p))
        return out

    def sample_measure(self, rnds):
        r = np.ascontiguousarray(rnds, dtype=np.float64)
        out = np.zeros(r.size, dtype=np.uint64)
        self._f("refqv_sample_measure")(self.h, r.ctypes.data_as(_f64p), C.c_int64(r.size),
                                        out.ctypes.data_as(_u64p))
        return out

    def expval_pauli(self, qubits, pauli, initial_phase=1.0):
        qa, qp, k = _q(qubits)
        c = complex(initial_phase)
        return self._f("refqv_expval_pauli")(self.h, qp, k, pauli.encode(), C.c_double(c.real),
                                             C.c_double(c.imag))
