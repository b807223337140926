"""ctypes binding of include/b200sv.h (the same stub a reference maintainer would
write for a Python consumer; the C++ consumer is include/qubitvector_b200.hpp)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200sv.so")

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# every symbol include/b200sv.h declares, with its argument types
SIGNATURES = {
    "b200sv_version": [],
    "b200sv_device_count": [C.POINTER(C.c_int)],
    "b200sv_trim": [],
    "b200sv_create": [C.POINTER(_vp), C.c_int, C.c_int64, C.c_int, C.c_int],
    "b200sv_create_external": [C.POINTER(_vp), C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _vp],
    "b200sv_destroy": [_vp],
    "b200sv_num_qubits": [_vp, C.POINTER(C.c_int)],
    "b200sv_device_ptr": [_vp, C.POINTER(_vp)],
    "b200sv_stream": [_vp, C.POINTER(_vp)],
    "b200sv_set_chunk": [_vp, C.c_int, C.c_uint64],
    "b200sv_synchronize": [_vp],
    "b200sv_initialize": [_vp],
    "b200sv_zero": [_vp],
    "b200sv_upload": [_vp, _vp, C.c_uint64, C.c_uint64],
    "b200sv_download": [_vp, _vp, C.c_uint64, C.c_uint64],
    "b200sv_download_line": [_vp, C.c_int, C.c_uint64, _vp],
    "b200sv_dm_expval_pauli": [_vp, C.c_int, _u64p, C.c_int, C.c_char_p, C.c_double, C.c_double, _f64p],
    "b200sv_dm_probabilities": [_vp, C.c_int, _u64p, C.c_int, _f64p],
    "b200sv_initialize_component": [_vp, _u64p, C.c_int, _f64p],
    "b200sv_checkpoint": [_vp],
    "b200sv_revert": [_vp, C.c_int],
    "b200sv_inner_product": [_vp, _f64p, _f64p],
    "b200sv_apply_matrix": [_vp, _u64p, C.c_int, _f64p],
    "b200sv_apply_diagonal": [_vp, _u64p, C.c_int, _f64p],
    "b200sv_apply_diagonal_layer": [_vp, C.c_int, C.POINTER(C.c_int), _u64p, _f64p],
    "b200sv_apply_multiplexer": [_vp, _u64p, C.c_int, _u64p, C.c_int, _f64p],
    "b200sv_apply_permutation": [_vp, _u64p, C.c_int, _u64p, C.c_int],
    "b200sv_apply_mcx": [_vp, _u64p, C.c_int],
    "b200sv_apply_mcy": [_vp, _u64p, C.c_int],
    "b200sv_apply_mcswap": [_vp, _u64p, C.c_int],
    "b200sv_apply_mcphase": [_vp, _u64p, C.c_int, C.c_double, C.c_double],
    "b200sv_apply_mcu": [_vp, _u64p, C.c_int, _f64p],
    "b200sv_apply_pauli": [_vp, _u64p, C.c_int, C.c_char_p, C.c_double, C.c_double],
    "b200sv_apply_batched_pauli": [_vp, _u64p],
    "b200sv_apply_op_sequence": [_vp, C.c_int, C.POINTER(C.c_int), _u64p, _f64p, C.POINTER(C.c_int),
                                 C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int)],
    "b200sv_apply_gate_sequence": [_vp, C.c_int, C.POINTER(C.c_int), _u64p, _f64p, C.POINTER(C.c_int)],
    "b200sv_plan_epochs": [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint8),
                           C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)],
    "b200sv_fuse_assign": [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                           C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "b200sv_fuse_block_matrix": [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64), _f64p, C.c_int, _f64p],
    "b200sv_selftest_op_sequence": [C.c_int, C.c_int64, C.c_int, _vp, C.c_int, C.POINTER(C.c_int), _u64p, _f64p,
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int)],
    "b200sv_collapse": [_vp, _u64p, C.c_int, _u64p, _f64p, C.POINTER(C.c_uint8)],
    "b200sv_apply_batched_matrix": [_vp, _u64p, C.c_int, _f64p, C.c_int, C.POINTER(C.c_int), _f64p],
    "b200sv_create_view": [C.POINTER(_vp), _vp, C.c_int64, C.c_int64],
    "b200sv_norm": [_vp, _f64p],
    "b200sv_norm_matrix": [_vp, _u64p, C.c_int, _f64p, _f64p],
    "b200sv_probabilities": [_vp, _u64p, C.c_int, _f64p],
    "b200sv_sample_measure": [_vp, _f64p, C.c_int64, _u64p],
    "b200sv_expval_pauli": [_vp, _u64p, C.c_int, C.c_char_p, C.c_double, C.c_double, _f64p],
    "b200sv_expval_pauli_pair": [_vp, _u64p, C.c_int, C.c_char_p, _vp, C.c_uint64, C.c_uint64, C.c_double,
                                 C.c_double, _f64p],
    "b200sv_chunk_swap_peer": [_vp, C.c_int, _vp, C.c_int, C.c_int],
    "b200sv_multi_swap_peer": [_vp, C.c_int, C.POINTER(C.c_int), C.c_uint32, C.POINTER(_vp)],
    "b200sv_swap_range_peer": [_vp, C.c_uint64, _vp, C.c_uint64, C.c_uint64],
    "b200sv_copy_range_peer": [_vp, C.c_uint64, _vp, C.c_uint64, C.c_uint64],
    "b200sv_mem_info": [C.c_int, _u64p, _u64p],
    "b200sv_measure_fp64_peak": [C.c_int, C.c_double, _f64p, _f64p],
    "b200sv_pack_half": [_vp, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _vp],
    "b200sv_unpack_half": [_vp, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _vp],
    "b200sv_ipc_export": [_vp, _vp],
    "b200sv_ipc_open": [_vp, _vp, C.POINTER(_vp)],
    "b200sv_ipc_close": [_vp, _vp],
    "b200sv_set_stream": [_vp, _vp],
    "b200sv_rng_uniform": [C.c_uint64, C.c_int64, _f64p],
    # sharded executor (csrc/sharded.cu)
    "b200sv_sharded_create": [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                              C.c_uint64],
    "b200sv_sharded_destroy": [_vp],
    "b200sv_sharded_ipc_export": [_vp, C.c_int, _vp],
    "b200sv_sharded_ipc_attach": [_vp, C.c_int, _vp],
    "b200sv_sharded_shard_handle": [_vp, C.c_int, C.POINTER(_vp)],
    "b200sv_sharded_initialize": [_vp],
    "b200sv_sharded_synchronize": [_vp],
    "b200sv_sharded_apply_ops": [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64), _f64p],
    "b200sv_sharded_plan_only": [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int64), _f64p, _f64p],
    "b200sv_sharded_selftest": [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.POINTER(C.c_int64), _f64p, _vp, _f64p],
    "b200sv_sharded_stats": [_vp, _f64p],
    "b200sv_sharded_profile": [_vp, C.c_int],
    "b200sv_sharded_profile_read": [_vp, _f64p],
    "b200sv_sharded_elapsed_ms": [_vp, _f64p],
    "b200sv_sharded_qubit_map": [_vp, C.POINTER(C.c_int)],
    "b200sv_sharded_restore_order": [_vp],
    "b200sv_sharded_norms": [_vp, _f64p],
    "b200sv_sharded_expval_pauli": [_vp, _u64p, C.c_int, C.c_char_p, _f64p],
    "b200sv_sharded_sample_measure": [_vp, _f64p, C.c_int64, _f64p, _u64p],
}

_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    """Load libb200sv.so; fail loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            f = getattr(l, name)
            f.argtypes = args
            f.restype = C.c_int
        l.b200sv_last_error.argtypes = []
        l.b200sv_last_error.restype = C.c_char_p
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise B200Error(lib().b200sv_last_error().decode())
