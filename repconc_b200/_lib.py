"""ctypes binding of librepconc_b200.so (the C ABI declared in include/repconc_b200.h).

There is NO fallback: if the shared library is missing or a call fails, the product raises.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librepconc_b200.so")

_lib = None

c_f32p = ctypes.c_void_p
c_ptr = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f64 = ctypes.c_double
c_f32 = ctypes.c_float
c_size = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/repconc_b200.h one to one
SIGNATURES = {
    "rc_last_error": (ctypes.c_char_p, []),
    "rc_version": (ctypes.c_char_p, []),
    "rc_launch_count": (c_i64, []),
    "rc_nn_assign": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "rc_minmax_init": (c_int, [c_ptr, c_int, c_ptr]),
    "rc_dist_table": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "rc_sinkhorn_set_dense": (c_int, [c_int]),
    "rc_sinkhorn_state_bytes": (c_size, [c_i64, c_int, c_int]),
    "rc_sinkhorn_rowsum_ptr": (c_ptr, [c_ptr, c_i64, c_int, c_int]),
    "rc_sinkhorn_begin": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_f64, c_ptr, c_ptr, c_ptr]),
    "rc_sinkhorn_debug_pool_entries": (c_i64, [c_i64]),
    "rc_sinkhorn_step": (c_int, [c_ptr, c_i64, c_i64, c_int, c_int, c_f64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "rc_sinkhorn_expand": (c_int, [c_ptr, c_i64, c_i64, c_int, c_int, c_f64, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "rc_sinkhorn_finish": (c_int, [c_ptr, c_i64, c_i64, c_int, c_int, c_f64, c_int, c_int, c_int, c_ptr, c_ptr,
                                   c_ptr, c_ptr, c_ptr]),
    "rc_sinkhorn_solve": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_f64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                  c_ptr]),
    "rc_sinkhorn_peer_buffer_bytes": (c_size, [c_int, c_int]),
    "rc_sinkhorn_solve_peer": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_f64, c_int, c_ptr, c_ptr, c_int,
                                       c_int, ctypes.c_uint32, c_ptr, c_ptr, c_ptr, c_ptr]),
    "rc_sinkhorn_debug_cta_times": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_int]),
    "rc_sinkhorn_list_stats": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr]),
    "rc_sinkhorn_debug_drift": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr]),
    "rc_peer_allreduce_buffer_bytes": (c_size, [c_i64]),
    "rc_peer_allreduce_f64": (c_int, [c_ptr, c_int, c_int, c_i64, ctypes.c_uint32, c_ptr, c_ptr, c_ptr]),
    "rc_decode": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "rc_code_histogram": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "rc_decode_bwd_workspace_bytes": (c_size, [c_i64, c_int, c_int, c_int]),
    "rc_decode_bwd": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "rc_mse_workspace_bytes": (c_size, [c_i64, c_int, c_int, c_int]),
    "rc_mse_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int,
                           c_int, c_int, c_f32, c_ptr, c_ptr, c_ptr]),
    "rc_mse_bwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int,
                           c_int, c_int, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "rc_encode_assign": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_ptr, c_i64, c_ptr,
                                 c_ptr, c_ptr]),
    "rc_adc_search_workspace_bytes": (c_size, [c_i64, c_i64, c_int, c_int, c_i64]),
    "rc_adc_search": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_i64, c_i64, c_ptr,
                              c_ptr, c_ptr, c_size, c_ptr]),
    "rc_adc_lut": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr]),
    "rc_adc_scores": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_ptr, c_ptr]),
    "rc_map_ids": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "rc_topk_merge": (c_int, [c_ptr, c_ptr, c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "rc_adc_last_stats": (None, [ctypes.POINTER(c_i64)]),
    "rc_adc_enable_timing": (None, [c_int]),
    "rc_adc_last_scan_ms": (c_f64, []),
    "rc_adc_last_scan_launches": (c_int, []),
    "rc_adc_last_scan_wavefronts": (c_f64, []),
    "rc_adc_last_scan_kernel": (ctypes.c_char_p, []),
}


class RepconcLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RepconcLibraryError(
                f"{LIB_PATH} not found: build it with `python -m repconc_b200.build` "
                "(repconc_b200 has no CPU or PyTorch fallback path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so is stale: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().rc_last_error().decode(errors="replace")
        raise RepconcLibraryError(f"{what} failed (rc={rc}): {msg}")


def launch_count():
    return int(load().rc_launch_count())
