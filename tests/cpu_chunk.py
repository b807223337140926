"""CPU stand-in for one sharded chunk, used only by the gloo world_size>1 tests of the host-side
scheduling logic (qiskit-aer_b200/sharded.py).  Arithmetic = the oracle; the global-qubit
resolution mirrors what the C ABI does on the device (api.cu: b200sv_apply_diagonal)."""
import numpy as np
import torch

from oracle.oracle import OracleQV


def _insert_zero(v, pos):
    low = v & ((1 << pos) - 1)
    return ((v >> pos) << (pos + 1)) | low


class CpuChunk:
    def __init__(self, n_local):
        self.n = n_local
        self.buf = torch.zeros((1 << n_local) * 2, dtype=torch.float64)
        self.ora = OracleQV(n_local)
        self.ora.psi = self.buf.numpy().view(np.complex128)  # share memory with the torch tensor
        self.ora.psi[0] = 1.0
        self.global_n, self.chunk_index = n_local, 0

    def chunk_setup(self, global_n, chunk_index):
        self.global_n, self.chunk_index = global_n, chunk_index

    def initialize(self):
        self.ora.psi[:] = 0
        self.ora.psi[0] = 1.0

    def zero(self):
        self.ora.psi[:] = 0

    def apply_matrix(self, qubits, mat):
        assert all(q < self.n for q in qubits)
        self.ora.apply_matrix(qubits, mat)

    def apply_diagonal_matrix(self, qubits, diag):
        diag = np.asarray(diag, dtype=np.complex128)
        lq, lbit, fixed = [], [], 0
        for j, q in enumerate(qubits):
            if q < self.n:
                lq.append(q)
                lbit.append(j)
            elif (self.chunk_index >> (q - self.n)) & 1:
                fixed |= 1 << j
        d2 = np.empty(1 << len(lq), dtype=np.complex128)
        for i in range(1 << len(lq)):
            src = fixed
            for b in range(len(lq)):
                if (i >> b) & 1:
                    src |= 1 << lbit[b]
            d2[i] = diag[src]
        if lq:
            self.ora.apply_diagonal_matrix(lq, d2)
        else:
            self.ora.psi *= d2[0]

    def apply_mcphase(self, qubits, phase):
        d = np.ones(1 << len(qubits), dtype=np.complex128)
        d[-1] = phase
        self.apply_diagonal_matrix(qubits, d)

    def apply_mcu(self, qubits, mat):
        self.ora.apply_mcu(qubits, mat)

    def apply_mcswap(self, qubits):
        self.ora.apply_mcswap(qubits)

    def pack_half(self, q, bit, begin, count, ptr):
        dst = self._tensor_at(ptr, count)
        idx = np.array([_insert_zero(begin + j, q) | (bit << q) for j in range(count)])
        dst[:] = self.ora.psi[idx]

    def unpack_half(self, q, bit, begin, count, ptr):
        src = self._tensor_at(ptr, count)
        idx = np.array([_insert_zero(begin + j, q) | (bit << q) for j in range(count)])
        self.ora.psi[idx] = src

    @staticmethod
    def _tensor_at(ptr, count):
        import ctypes
        buf = (ctypes.c_char * (count * 16)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.complex128)

    def norm(self):
        return self.ora.norm()

    def expval_pauli(self, qubits, pauli):
        return self.ora.expval_pauli(qubits, pauli)

    def sample_measure(self, rnds):
        return self.ora.sample_measure(rnds)
