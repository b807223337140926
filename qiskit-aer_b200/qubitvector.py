"""Host-side mirror of the reference's QubitVector interface over the C ABI.

Same method names, argument meaning and error behaviour as
``AER::QV::QubitVector`` (/root/reference/src/simulators/statevector/
qubitvector.hpp:62-651): qubit lists with controls first / targets last,
column-major vectorised matrices, ``pauli[N-1-i] <-> qubits[i]``, errors raised
as exceptions.  All arithmetic happens in libb200sv.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import capi

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def _q(qubits):
    a = np.ascontiguousarray(list(qubits), dtype=np.uint64)
    return a, a.ctypes.data_as(_u64p), int(a.size)


def _m(mat):
    a = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128).reshape(-1))
    return a, a.ctypes.data_as(_f64p)


class QubitVectorB200:
    """``num_states`` statevectors of ``num_qubits`` qubits resident on one B200."""

    @staticmethod
    def name():
        return "statevector_b200"  # cf. QubitVector::name() "statevector"

    def __init__(self, num_qubits, dtype=np.complex128, num_states=1, device=0, external_ptr=None, stream=None):
        self._lib = capi.lib()
        self.n = int(num_qubits)
        self.num_states = int(num_states)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.complex128), np.dtype(np.complex64)):
            raise ValueError("precision must be complex128 (double) or complex64 (single)")
        prec = 64 if self.dtype == np.complex128 else 32
        self.h = C.c_void_p()
        if external_ptr is None:
            capi.check(self._lib.b200sv_create(C.byref(self.h), self.n, self.num_states, prec, int(device)))
        else:
            capi.check(self._lib.b200sv_create_external(C.byref(self.h), self.n, self.num_states, prec, int(device),
                                                        C.c_void_p(int(external_ptr)), C.c_void_p(int(stream or 0))))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.b200sv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ---------------------------------------------------
    def num_qubits(self):
        return self.n

    def size(self):
        return 1 << self.n

    def chunk_setup(self, global_num_qubits, chunk_index):
        capi.check(self._lib.b200sv_set_chunk(self.h, int(global_num_qubits), int(chunk_index)))

    def synchronize(self):
        capi.check(self._lib.b200sv_synchronize(self.h))

    def device_ptr(self):
        p = C.c_void_p()
        capi.check(self._lib.b200sv_device_ptr(self.h, C.byref(p)))
        return p.value

    # ---- data ------------------------------------------------------------
    def initialize(self):
        capi.check(self._lib.b200sv_initialize(self.h))

    def zero(self):
        capi.check(self._lib.b200sv_zero(self.h))

    def initialize_from_vector(self, vec, offset=0):
        a = np.ascontiguousarray(vec, dtype=self.dtype).reshape(-1)
        capi.check(self._lib.b200sv_upload(self.h, C.c_void_p(a.ctypes.data), int(offset), int(a.size)))

    set_state = initialize_from_vector

    def initialize_component(self, qubits, state):
        qa, qp, k = _q(qubits)
        ma, mp = _m(state)
        capi.check(self._lib.b200sv_initialize_component(self.h, qp, k, mp))

    def vector(self, offset=0, count=None):
        if count is None:
            count = (self.num_states << self.n) - offset
        out = np.empty(int(count), dtype=self.dtype)
        capi.check(self._lib.b200sv_download(self.h, C.c_void_p(out.ctypes.data), int(offset), int(count)))
        return out

    copy_to_vector = vector

    def checkpoint(self):
        capi.check(self._lib.b200sv_checkpoint(self.h))

    def revert(self, keep):
        capi.check(self._lib.b200sv_revert(self.h, int(bool(keep))))

    def inner_product(self):
        re = np.zeros(self.num_states)
        im = np.zeros(self.num_states)
        capi.check(self._lib.b200sv_inner_product(self.h, re.ctypes.data_as(_f64p), im.ctypes.data_as(_f64p)))
        z = re + 1j * im
        return complex(z[0]) if self.num_states == 1 else z

    # ---- gates -------------------------------------------------------------
    def apply_matrix(self, qubits, mat):
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        if ma.size != 1 << (2 * k):
            raise ValueError("matrix size does not match the qubit count")
        capi.check(self._lib.b200sv_apply_matrix(self.h, qp, k, mp))

    def apply_diagonal_matrix(self, qubits, diag):
        qa, qp, k = _q(qubits)
        ma, mp = _m(diag)
        if ma.size != 1 << k:
            raise ValueError("diagonal size does not match the qubit count")
        capi.check(self._lib.b200sv_apply_diagonal(self.h, qp, k, mp))

    def apply_multiplexer(self, control_qubits, target_qubits, mat):
        ca, cp, nc = _q(control_qubits)
        ta, tp, nt = _q(target_qubits)
        ma, mp = _m(mat)
        capi.check(self._lib.b200sv_apply_multiplexer(self.h, cp, nc, tp, nt, mp))

    def apply_permutation_matrix(self, qubits, pairs):
        qa, qp, k = _q(qubits)
        pa = np.ascontiguousarray(np.asarray(pairs, dtype=np.uint64).reshape(-1))
        capi.check(self._lib.b200sv_apply_permutation(self.h, qp, k, pa.ctypes.data_as(_u64p), pa.size // 2))

    def apply_mcx(self, qubits):
        qa, qp, k = _q(qubits)
        capi.check(self._lib.b200sv_apply_mcx(self.h, qp, k))

    def apply_mcy(self, qubits):
        qa, qp, k = _q(qubits)
        capi.check(self._lib.b200sv_apply_mcy(self.h, qp, k))

    def apply_mcswap(self, qubits):
        qa, qp, k = _q(qubits)
        capi.check(self._lib.b200sv_apply_mcswap(self.h, qp, k))

    def apply_mcphase(self, qubits, phase):
        qa, qp, k = _q(qubits)
        ph = complex(phase)
        capi.check(self._lib.b200sv_apply_mcphase(self.h, qp, k, ph.real, ph.imag))

    def apply_mcu(self, qubits, mat):
        qa, qp, k = _q(qubits)
        ma, mp = _m(mat)
        capi.check(self._lib.b200sv_apply_mcu(self.h, qp, k, mp))

    def apply_pauli(self, qubits, pauli, coeff=1.0):
        qa, qp, k = _q(qubits)
        c = complex(coeff)
        capi.check(self._lib.b200sv_apply_pauli(self.h, qp, k, pauli.encode(), c.real, c.imag))

    def apply_gate_sequence(self, gates):
        """gates: list of (qubits, column-major matrix) with 1 or 2 qubits each; applied in order, as few
        HBM passes as possible (b200sv_apply_gate_sequence).  Returns the number of passes used."""
        ng = len(gates)
        nq = np.zeros(ng, dtype=np.int32)
        qs = np.zeros(2 * ng, dtype=np.uint64)
        mats = np.zeros((ng, 16), dtype=np.complex128)
        for i, (q, m) in enumerate(gates):
            nq[i] = len(q)
            qs[2 * i:2 * i + len(q)] = q
            m = np.asarray(m, dtype=np.complex128).reshape(-1)
            mats[i, :m.size] = m
        passes = C.c_int(0)
        capi.check(self._lib.b200sv_apply_gate_sequence(self.h, ng, nq.ctypes.data_as(C.POINTER(C.c_int)),
                                                        qs.ctypes.data_as(_u64p), mats.ctypes.data_as(_f64p),
                                                        C.byref(passes)))
        return passes.value

    def apply_op_sequence(self, ops, codes=None):
        """ops: list of ("dense", qubits, column-major matrix) | ("pauli", qubit, slot); codes: uint8 array
        [nslots][num_states] of per-state Pauli codes (0..3 = I, X, Y, Z).  Returns the passes used."""
        nops = len(ops)
        kind = np.zeros(nops, dtype=np.int32)
        qs = np.zeros(2 * nops, dtype=np.uint64)
        mats = np.zeros((nops, 16), dtype=np.complex128)
        slot = np.zeros(nops, dtype=np.int32)
        for i, op in enumerate(ops):
            if op[0] == "dense":
                q = list(op[1])
                kind[i] = len(q)
                qs[2 * i:2 * i + len(q)] = q
                m = np.asarray(op[2], dtype=np.complex128).reshape(-1)
                mats[i, :m.size] = m
            elif op[0] == "pauli":
                kind[i] = 3
                qs[2 * i] = op[1]
                slot[i] = op[2]
            else:
                raise ValueError(op[0])
        if codes is None:
            codes_arr, nslots, cp = None, 0, None
        else:
            codes_arr = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1, self.num_states)
            nslots = codes_arr.shape[0]
            cp = codes_arr.ctypes.data_as(C.POINTER(C.c_uint8))
        passes = C.c_int(0)
        capi.check(self._lib.b200sv_apply_op_sequence(self.h, nops, kind.ctypes.data_as(C.POINTER(C.c_int)),
                                                      qs.ctypes.data_as(_u64p), mats.ctypes.data_as(_f64p),
                                                      slot.ctypes.data_as(C.POINTER(C.c_int)), cp, nslots,
                                                      C.byref(passes)))
        return passes.value

    def apply_batched_pauli_ops(self, masks4):
        """masks4: [num_states][4] = x_mask, z_mask, num_y, apply (qubitvector_thrust.hpp:2892)."""
        a = np.ascontiguousarray(masks4, dtype=np.uint64).reshape(-1)
        if a.size != 4 * self.num_states:
            raise ValueError("need 4 words per state")
        capi.check(self._lib.b200sv_apply_batched_pauli(self.h, a.ctypes.data_as(_u64p)))

    def apply_diagonal_layer(self, gates):
        """gates: [(qubits (1 or 2), diagonal (2 or 4 complex))] -- commuting diagonal gates, one streaming pass."""
        ng = len(gates)
        nq = np.zeros(ng, dtype=np.int32)
        qs = np.zeros(2 * ng, dtype=np.uint64)
        dg = np.zeros((ng, 4), dtype=np.complex128)
        for i, (q, d) in enumerate(gates):
            nq[i] = len(q)
            qs[2 * i:2 * i + len(q)] = q
            dg[i, :1 << len(q)] = np.asarray(d, dtype=np.complex128).reshape(-1)
        capi.check(self._lib.b200sv_apply_diagonal_layer(self.h, ng, nq.ctypes.data_as(C.POINTER(C.c_int)),
                                                         qs.ctypes.data_as(_u64p), dg.ctypes.data_as(_f64p)))

    def apply_batched_matrix(self, qubits, mats, index, scale=None):
        """State s applies the column-major matrix mats[index[s]] * scale[s] (index < 0: untouched) -- one launch
        (batched Kraus / per-parameter matrices, qubitvector_thrust.hpp:1578-1611,2996-3177)."""
        qa, qp, k = _q(qubits)
        m = np.ascontiguousarray(np.asarray(mats, dtype=np.complex128).reshape(-1))
        nm = m.size >> (2 * k)
        ix = np.ascontiguousarray(index, dtype=np.int32)
        sc = np.ascontiguousarray(np.ones(self.num_states) if scale is None else scale, dtype=np.float64)
        if ix.size != self.num_states or sc.size != self.num_states:
            raise ValueError("need one index / scale per state")
        capi.check(self._lib.b200sv_apply_batched_matrix(self.h, qp, k, m.ctypes.data_as(_f64p), int(nm),
                                                         ix.ctypes.data_as(C.POINTER(C.c_int)), sc.ctypes.data_as(_f64p)))

    # ---- reductions ----------------------------------------------------------
    def _ret(self, out):
        return float(out[0]) if self.num_states == 1 else out

    def norm(self, qubits=None, mat=None):
        out = np.zeros(self.num_states)
        if qubits is None:
            capi.check(self._lib.b200sv_norm(self.h, out.ctypes.data_as(_f64p)))
        else:
            qa, qp, k = _q(qubits)
            ma, mp = _m(mat)
            capi.check(self._lib.b200sv_norm_matrix(self.h, qp, k, mp, out.ctypes.data_as(_f64p)))
        return self._ret(out)

    def probabilities(self, qubits=None):
        if qubits is None:
            qubits = list(range(self.n))
        qa, qp, k = _q(qubits)
        out = np.zeros((self.num_states, 1 << k))
        capi.check(self._lib.b200sv_probabilities(self.h, qp, k, out.ctypes.data_as(_f64p)))
        return out[0] if self.num_states == 1 else out

    def sample_measure(self, rnds):
        r = np.ascontiguousarray(rnds, dtype=np.float64)
        shots = r.size // self.num_states
        out = np.zeros(r.size, dtype=np.uint64)
        capi.check(self._lib.b200sv_sample_measure(self.h, r.ctypes.data_as(_f64p), int(shots),
                                                   out.ctypes.data_as(_u64p)))
        return out.reshape(r.shape)

    def expval_pauli(self, qubits, pauli, initial_phase=1.0):
        qa, qp, k = _q(qubits)
        c = complex(initial_phase)
        out = np.zeros(self.num_states)
        capi.check(self._lib.b200sv_expval_pauli(self.h, qp, k, pauli.encode(), c.real, c.imag,
                                                 out.ctypes.data_as(_f64p)))
        return self._ret(out)

    # ---- exchange --------------------------------------------------------------
    def set_stream(self, cuda_stream):
        capi.check(self._lib.b200sv_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def ipc_export(self):
        buf = C.create_string_buffer(64)
        capi.check(self._lib.b200sv_ipc_export(self.h, buf))
        return bytes(buf.raw)

    def ipc_open(self, handle_bytes):
        p = C.c_void_p()
        capi.check(self._lib.b200sv_ipc_open(self.h, C.c_char_p(handle_bytes), C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        capi.check(self._lib.b200sv_ipc_close(self.h, C.c_void_p(int(ptr))))

    def torch_view(self):
        """Zero-copy torch tensor (float64/float32 pairs) over the library-owned amplitudes."""
        import torch
        comp = "<f8" if self.dtype == np.complex128 else "<f4"

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": ((self.num_states << self.n) * 2,), "typestr": comp,
                                      "data": (self.device_ptr(), False), "version": 3}
        return torch.as_tensor(a, device="cuda")

    def pack_half(self, local_q, bit, begin, count, dev_buf):
        capi.check(self._lib.b200sv_pack_half(self.h, int(local_q), int(bit), int(begin), int(count),
                                              C.c_void_p(int(dev_buf))))

    def unpack_half(self, local_q, bit, begin, count, dev_buf):
        capi.check(self._lib.b200sv_unpack_half(self.h, int(local_q), int(bit), int(begin), int(count),
                                                C.c_void_p(int(dev_buf))))

    def multi_swap_peer(self, local_qs, my_g, peer_ptrs):
        """k local qubits <-> k global bits in one pass; peer_ptrs[v] = mapped slice of the rank whose k global
        bits read v (b200sv_multi_swap_peer)."""
        k = len(local_qs)
        lq = (C.c_int * k)(*[int(q) for q in local_qs])
        pp = (C.c_void_p * (1 << k))(*[C.c_void_p(int(p) if p else 0) for p in peer_ptrs])
        capi.check(self._lib.b200sv_multi_swap_peer(self.h, k, lq, int(my_g), pp))

    def chunk_swap_peer(self, local_q, peer_ptr, this_is_upper, half):
        capi.check(self._lib.b200sv_chunk_swap_peer(self.h, int(local_q), C.c_void_p(int(peer_ptr)),
                                                    int(this_is_upper), int(half)))


def rng_uniform(seed, n):
    """n draws of Aer's RngEngine::rand(0,1) for ``seed`` (framework/rng.hpp)."""
    out = np.zeros(int(n))
    capi.check(capi.lib().b200sv_rng_uniform(int(seed), int(n), out.ctypes.data_as(_f64p)))
    return out
