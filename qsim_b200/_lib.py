"""ctypes binding of libqsim_b200.so (the C ABI declared in include/qsim_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing this
module raises at import of the symbols, and every compute entry point fails
with a status code when no GPU is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqsim_b200.so")

OK, ERR_CUDA, ERR_OOM, ERR_INVALID, ERR_UNSUPPORTED = range(5)
F32, F64 = 0, 1

_vp, _u, _u64, _i, _d = C.c_void_p, C.c_uint, C.c_uint64, C.c_int, C.c_double
_pu, _pd, _pu64 = C.POINTER(C.c_uint), C.POINTER(C.c_double), C.POINTER(C.c_uint64)

# name -> (restype, argtypes); kept in sync with include/qsim_b200.h
# (tests/test_abi.py parses the header and checks every declared symbol).
SIGNATURES = {
    "qb200_abi_version": (_i, []),
    "qb200_device_count": (_i, [C.POINTER(_i)]),
    "qb200_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "qb200_ctx_destroy": (_i, [_vp]),
    "qb200_ctx_set_stream": (_i, [_vp, _vp]),
    "qb200_last_cuda_error": (_i, [_vp]),
    "qb200_last_cuda_error_string": (C.c_char_p, [_vp]),
    "qb200_launch_count": (_u64, [_vp]),
    "qb200_ctx_set_tuning": (_i, [_vp, C.c_char_p, _i]),
    "qb200_timer_start": (_i, [_vp]),
    "qb200_timer_stop_ms": (_i, [_vp, C.POINTER(C.c_float)]),
    "qb200_min_size": (_u64, [_u]),
    "qb200_state_alloc": (_i, [_u, _i, C.POINTER(_vp)]),
    "qb200_state_free": (_i, [_vp]),
    "qb200_copy_d2d": (_i, [_vp, _i, _vp, _vp, _u64]),
    "qb200_copy_d2d_async": (_i, [_vp, _i, _vp, _vp, _u64]),
    "qb200_copy_d2h": (_i, [_vp, _i, _vp, _vp, _u64]),
    "qb200_copy_h2d": (_i, [_vp, _i, _vp, _vp, _u64]),
    "qb200_sync": (_i, [_vp]),
    "qb200_device_sync": (_i, []),
    "qb200_apply_gate": (_i, [_vp, _i, _vp, _u, _pu, _u, _vp]),
    "qb200_apply_controlled_gate": (_i, [_vp, _i, _vp, _u, _pu, _u, _pu, _u, _u64, _vp]),
    "qb200_expectation_value": (_i, [_vp, _i, _vp, _u, _pu, _u, _vp, _pd]),
    "qb200_one_qubit_moments": (_i, [_vp, _i, _vp, _u, _pd]),
    "qb200_reduce_batch_begin": (_i, [_vp, C.c_uint32]),
    "qb200_reduce_batch_end": (_i, [_vp, _pd, C.c_uint32, C.POINTER(C.c_uint32)]),
    "qb200_set_all_zeros": (_i, [_vp, _i, _vp, _u]),
    "qb200_set_state_zero": (_i, [_vp, _i, _vp, _u]),
    "qb200_set_state_uniform": (_i, [_vp, _i, _vp, _u]),
    "qb200_get_ampl": (_i, [_vp, _i, _vp, _u64, _pd]),
    "qb200_set_ampl": (_i, [_vp, _i, _vp, _u64, _d, _d]),
    "qb200_bulk_set_ampl": (_i, [_vp, _i, _vp, _u, _u64, _u64, _d, _d, _i]),
    "qb200_add": (_i, [_vp, _i, _vp, _vp, _u]),
    "qb200_multiply": (_i, [_vp, _i, _d, _vp, _u]),
    "qb200_inner_product": (_i, [_vp, _i, _vp, _vp, _u, _pd]),
    "qb200_real_inner_product": (_i, [_vp, _i, _vp, _vp, _u, _pd]),
    "qb200_norm": (_i, [_vp, _i, _vp, _u, _pd]),
    "qb200_sample": (_i, [_vp, _i, _vp, _u, _pd, _u64, _pu64]),
    "qb200_generate_random_values": (_i, [_u64, _u, _d, _pd]),
    "qb200_mutation_epoch": (_u64, []),
    "qb200_expectation_values_multi": (_i, [_vp, _i, _vp, _u, _pu, _u, _vp, _u, _pd]),
    "qb200_sample_seeded": (_i, [_vp, _i, _vp, _u, _u64, _u, _d, _pu64]),
    "qb200_generate_random_values_device": (_i, [_vp, _u64, _u, _d, _pd]),
    "qb200_partial_norms_count": (_u64, [_u]),
    "qb200_partial_norms": (_i, [_vp, _i, _vp, _u, _pd]),
    "qb200_find_measured_bits": (_i, [_vp, _i, _vp, _u, _u64, _d, _u64, _pu64]),
    "qb200_collapse": (_i, [_vp, _i, _vp, _u, _u64, _u64, _pd]),
    "qb200_internal_to_normal_order": (_i, [_vp, _i, _vp, _u]),
    "qb200_normal_to_internal_order": (_i, [_vp, _i, _vp, _u]),
    "qb200_ipc_export": (_i, [_vp, C.POINTER(C.c_ubyte)]),
    "qb200_ipc_import": (_i, [C.POINTER(C.c_ubyte), C.POINTER(_vp)]),
    "qb200_ipc_close": (_i, [_vp]),
    "qb200_swap_global_local": (_i, [_vp, _i, _vp, _u, C.POINTER(_vp), _u, _pu, _u]),
}


class Comm(C.Structure):
    """qb200_comm: host-side collectives for a multi-process sharded state."""
    ALLGATHER = C.CFUNCTYPE(_i, _vp, _vp, _vp, _u64)
    ALLREDUCE = C.CFUNCTYPE(_i, _vp, _pd, _u64)
    BARRIER = C.CFUNCTYPE(_i, _vp)
    _fields_ = [("user", _vp), ("allgather", ALLGATHER), ("allreduce_sum_f64", ALLREDUCE), ("barrier", BARRIER)]


class Gate(C.Structure):
    """qb200_gate"""
    _fields_ = [("num_targets", _u), ("qs", _pu), ("num_controls", _u), ("cqs", _pu), ("cvals", _u64), ("matrix", _vp)]


class SvStats(C.Structure):
    """qb200_sv_stats"""
    _fields_ = [("swaps", _u64), ("local_swap_passes", _u64), ("gate_passes", _u64),
                ("bytes_sent_per_shard", _d), ("exchange_ms", _d), ("barrier_wait_ms", _d),
                ("overlapped_swaps", _u64), ("overlapped_gate_passes", _u64), ("overlap_ms", _d),
                ("copy_engine_swaps", _u64)]


SIGNATURES.update({
    "qb200_device_sync_on": (_i, [_i]),
    "qb200_state_alloc_on": (_i, [_vp, _u, _i, C.POINTER(_vp)]),
    "qb200_last_kernel_name": (C.c_char_p, [_vp]),
    "qb200_ctx_set_sm_limit": (_i, [_vp, _i]),
    "qb200_ctx_set_occupancy_reduction": (_i, [_vp, _i]),
    "qb200_masked_norm": (_i, [_vp, _i, _vp, _u, _u64, _u64, _pd]),
    "qb200_collapse_scaled": (_i, [_vp, _i, _vp, _u, _u64, _u64, _d]),
    "qb200_sv_create": (_i, [C.POINTER(_i), _u, _u, _i, C.POINTER(_vp)]),
    "qb200_sv_create_mp": (_i, [_i, _u, _u, C.POINTER(Comm), _u, _i, C.POINTER(_vp)]),
    "qb200_sv_destroy": (_i, [_vp]),
    "qb200_sv_num_qubits": (_u, [_vp]),
    "qb200_sv_num_shards": (_u, [_vp]),
    "qb200_sv_num_local_qubits": (_u, [_vp]),
    "qb200_sv_num_local_shards": (_u, [_vp]),
    "qb200_sv_last_cuda_error": (_i, [_vp]),
    "qb200_sv_qubit_map": (_i, [_vp, _pu]),
    "qb200_sv_shard": (_i, [_vp, _u, _pu, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp)]),
    "qb200_sv_set_option": (_i, [_vp, C.c_char_p, _i]),
    "qb200_sv_sync": (_i, [_vp]),
    "qb200_sv_launch_count": (_u64, [_vp]),
    "qb200_sv_get_stats": (_i, [_vp, C.POINTER(SvStats)]),
    "qb200_sv_reset_stats": (_i, [_vp]),
    "qb200_sv_set_all_zeros": (_i, [_vp]),
    "qb200_sv_set_state_zero": (_i, [_vp]),
    "qb200_sv_set_state_uniform": (_i, [_vp]),
    "qb200_sv_reset_map": (_i, [_vp]),
    "qb200_sv_get_ampl": (_i, [_vp, _u64, _pd]),
    "qb200_sv_get_ampls": (_i, [_vp, _pu64, _u64, _pd]),
    "qb200_sv_set_ampl": (_i, [_vp, _u64, _d, _d]),
    "qb200_sv_bulk_set_ampl": (_i, [_vp, _u64, _u64, _d, _d, _i]),
    "qb200_sv_norm": (_i, [_vp, _pd]),
    "qb200_sv_inner_product": (_i, [_vp, _vp, _pd]),
    "qb200_sv_add": (_i, [_vp, _vp]),
    "qb200_sv_copy": (_i, [_vp, _vp]),
    "qb200_sv_multiply": (_i, [_vp, _d]),
    "qb200_sv_sample": (_i, [_vp, _pd, _u64, _pu64]),
    "qb200_sv_partial_norms_count": (_u64, [_vp]),
    "qb200_sv_partial_norms": (_i, [_vp, _pd]),
    "qb200_sv_find_measured_bits": (_i, [_vp, _u64, _d, _u64, _pu64]),
    "qb200_sv_collapse": (_i, [_vp, _u64, _u64, _pd]),
    "qb200_sv_copy_to_host": (_i, [_vp, _vp]),
    "qb200_sv_copy_from_host": (_i, [_vp, _vp]),
    "qb200_sv_apply_gate": (_i, [_vp, _pu, _u, _vp]),
    "qb200_sv_apply_controlled_gate": (_i, [_vp, _pu, _u, _pu, _u, _u64, _vp]),
    "qb200_sv_expectation_value": (_i, [_vp, _pu, _u, _vp, _pd]),
    "qb200_sv_run": (_i, [_vp, C.POINTER(Gate), _u64]),
    "qb200_sv_plan": (_i, [_u, _u, C.POINTER(Gate), _u64, _pu, _i, C.POINTER(C.c_int64), _u64, _pu64]),
    "qb200_sv_plan_initial": (_i, [_u, _u, _vp, _u64, _i, _pu]),
    "qb200_sv_swap": (_i, [_vp, _pu, _pu, _u]),
    "qb200_sv_canonicalize": (_i, [_vp]),
})

_lib = None


def load():
    """Loads libqsim_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the qsim_b200 product path has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class QB200Error(RuntimeError):
    def __init__(self, status, what, detail=""):
        names = {1: "CUDA error", 2: "out of device memory", 3: "invalid argument",
                 4: "unsupported gate size"}
        super().__init__(f"{what}: {names.get(status, status)} {detail}".strip())
        self.status = status
