"""Loader for libqcs.so (the CUDA engine).  There is no fallback: if the
library is missing this raises, and the library itself refuses to create a
register without a CUDA device."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QCS_LIB_PATH: development builds of the same library (e.g. the -DQCS_PIPE_TIMING build of `make timing`)
LIB_PATH = os.environ.get("QCS_LIB_PATH") or os.path.join(_HERE, "lib", "libqcs.so")

_u = C.c_uint
_ull = C.c_ulonglong
_ll = C.c_longlong
_vp = C.c_void_p
_dp = C.POINTER(C.c_double)

# (name, restype, argtypes): every symbol declared in include/qcs.h
SIGNATURES = [
    ("qcs_version", C.c_char_p, []),
    ("qcs_error_string", C.c_char_p, [C.c_int]),
    ("qcs_device_count", C.c_int, []),
    ("qcs_register_create", C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    ("qcs_register_destroy", None, [_vp]),
    ("qcs_comm_unique_id", C.c_int, [_vp]),
    ("qcs_register_create_sharded", C.c_int,
     [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    ("qcs_register_create_multi", C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    ("qcs_num_gpus", C.c_int, [_vp]),
    ("qcs_L_size", C.c_int, [_vp]),
    ("qcs_M_size", C.c_int, [_vp]),
    ("qcs_num_qubits", _u, [_vp]),
    ("qcs_num_states", _ull, [_vp]),
    ("qcs_local_states", _ull, [_vp]),
    ("qcs_rank", C.c_int, [_vp]),
    ("qcs_world_size", C.c_int, [_vp]),
    ("qcs_peer_memory", C.c_int, [_vp]),
    ("qcs_set_option", C.c_int, [_vp, C.c_int, _ll]),
    ("qcs_get_option", _ll, [_vp, C.c_int]),
    ("qcs_synchronize", C.c_int, [_vp]),
    ("qcs_reset_register", C.c_int, [_vp]),
    ("qcs_hadamard_gate", C.c_int, [_vp, _u]),
    ("qcs_c_phase_shift_gate", C.c_int, [_vp, _u, _u, C.c_double]),
    ("qcs_c_amodc_gate", C.c_int, [_vp, _u, _ull, _u]),
    ("qcs_inverse_QFT", C.c_int, [_vp]),
    ("qcs_QFT", C.c_int, [_vp]),
    ("qcs_inverse_QFT_range", C.c_int, [_vp, _u, _u]),
    ("qcs_QFT_range", C.c_int, [_vp, _u, _u]),
    ("qcs_quantum_computation", C.c_int, [_vp, _u, _u, C.c_int]),
    ("qcs_measure_state", C.c_int, [_vp, C.c_double, C.POINTER(_ull)]),
    ("qcs_norm2", C.c_int, [_vp, _dp]),
    ("qcs_nonzero_states", C.c_int, [_vp, _ull, C.POINTER(_ull), _dp, C.POINTER(_ull)]),
    ("qcs_get_state", C.c_int, [_vp, _ull, _ull, _vp]),
    ("qcs_set_state", C.c_int, [_vp, _ull, _ull, _vp]),
    ("qcs_set_state_async", C.c_int, [_vp, _ull, _ull, _vp]),
    ("qcs_sample_states", C.c_int, [_vp, _ull, _vp, _vp]),
    ("qcs_apply_gate", C.c_int, [_vp, _u, _vp]),
    ("qcs_apply_controlled_gate", C.c_int, [_vp, _u, _u, _vp]),
    ("qcs_save_state", C.c_int, [_vp, C.c_char_p]),
    ("qcs_load_state", C.c_int, [_vp, C.c_char_p]),
    ("qcs_schedule_describe", C.c_int, [_u, C.c_int, C.c_int, _ull, _vp, _vp, _vp, _vp, _ull]),
    ("qcs_plan_describe", C.c_int, [_u, _u, _u, C.c_int, C.c_int, _vp, _ull]),
    ("qcs_pair_selfcheck", C.c_longlong, [_u, _u, _u, C.c_int]),
    ("qcs_fuse_begin", C.c_int, [_vp]),
    ("qcs_fuse_end", C.c_int, [_vp]),
    ("qcs_fuse_pending", _ull, [_vp]),
    ("qcs_apply_dense_block", C.c_int, [_vp, _u, _vp]),
    ("qcs_fill_synthetic", C.c_int, [_vp, _ull]),
    ("qcs_scale", C.c_int, [_vp, C.c_double]),
    ("qcs_int_pow", _u, [_u, _u]),
    ("qcs_modpow2k", _ull, [_u, _u, _u]),
    ("qcs_host_alloc", C.c_int, [C.POINTER(_vp), C.c_size_t]),
    ("qcs_host_alloc_near", C.c_int, [C.POINTER(_vp), C.c_size_t, C.c_int]),
    ("qcs_host_free", C.c_int, [_vp]),
    ("qcs_timer_start", C.c_int, [_vp]),
    ("qcs_timer_stop", C.c_int, [_vp, _dp]),
    ("qcs_launch_count", _ull, [_vp]),
    ("qcs_profile_reset", C.c_int, [_vp]),
    ("qcs_profile_get", C.c_int, [_vp, C.c_int, C.POINTER(_ull), _dp, _dp]),
    ("qcs_kernel_class_name", C.c_char_p, [C.c_int]),
]

_lib = None


def load():
    """Load libqcs.so and attach prototypes.  Raises OSError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            f"{LIB_PATH} is missing: build it with `make lib` (or __graft_entry__.build()). "
            "quantumcomputer_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, restype, argtypes in SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
