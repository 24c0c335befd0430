"""
ctypes binding of libphylo_b200.so (the C ABI declared in include/phylo_b200.h).

The library is the only compute back end: if it cannot be loaded this module raises - there is
no Python / CPU fallback for the likelihood path.  ``check()`` turns status codes into the Python
exceptions the reference raises in the same situations (ValueError for bad arguments,
RuntimeError otherwise).
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_size_t, c_uint, c_double, c_void_p, c_char_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHB_LIBRARY") or os.path.join(_HERE, "libphylo_b200.so")   # PHB_LIBRARY: developer override (what-if builds)

PHB_OK, PHB_ERR_INVALID, PHB_ERR_CUDA, PHB_ERR_NO_DEVICE, PHB_ERR_STATE, PHB_ERR_NOMEM, PHB_ERR_UNSUPPORTED = range(7)
PHB_FLAG_UP_PARTIALS = 0x1
PHB_FLAG_NO_PARTIALS = 0x2
PHB_MODE_AUTO, PHB_MODE_TILE, PHB_MODE_LEVEL, PHB_MODE_RESIDENT = 0, 1, 2, 3

_dp = POINTER(c_double)
_ip = POINTER(c_int32)
_lp = POINTER(c_int64)
_bp = POINTER(ctypes.c_uint8)

# name -> (restype, argtypes); kept in one table so tests can compare it with the header
SIGNATURES = {
    "phb_version": (c_int, []),
    "phb_status_name": (c_char_p, [c_int]),
    "phb_last_error": (c_char_p, [c_void_p]),
    "phb_reload_tuning": (c_int, []),
    "phb_launch_count": (c_int64, [c_void_p]),
    "phb_discrete_gamma": (c_int, [c_double, c_double, c_int, c_int, _dp, _dp]),
    "phb_compress_patterns": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, _lp, _lp, _lp, _lp]),
    "phb_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int, c_uint]),
    "phb_create": (c_int, [c_int, c_int, c_int64, c_int, c_int, c_uint, c_void_p, c_size_t, c_void_p,
                           POINTER(c_void_p)]),
    "phb_destroy": (c_int, [c_void_p]),
    "phb_sync": (c_int, [c_void_p]),
    "phb_set_tips": (c_int, [c_void_p, c_void_p, c_int, c_int, _dp, _ip]),
    "phb_set_pattern_weights": (c_int, [c_void_p, _lp]),
    "phb_set_model": (c_int, [c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "phb_set_mixture": (c_int, [c_void_p, _dp, _dp, _dp]),
    "phb_set_schedule": (c_int, [c_void_p, c_int, _ip, c_int, _ip]),
    "phb_set_edge_lengths": (c_int, [c_void_p, _dp]),
    "phb_build_pmatrices": (c_int, [c_void_p]),
    "phb_set_pmatrices": (c_int, [c_void_p, _dp]),
    "phb_get_pmatrix": (c_int, [c_void_p, c_int, c_int, _dp]),
    "phb_compute_partials": (c_int, [c_void_p, c_int]),
    "phb_root_lnl": (c_int, [c_void_p, c_int, c_int, c_double, _dp, _dp, _dp, _dp]),
    "phb_lnl_resident": (c_int, [c_void_p, c_int, c_int, c_double, _dp, _dp]),
    "phb_lnl_from_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, _dp, _dp]),
    "phb_lnl_from_host_packed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, _dp, _dp]),
    "phb_pack_codes": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    "phb_lnl_from_host_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, _dp, _dp]),
    "phb_split_codes": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "phb_lnl_from_host_split_async": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double]),
    "phb_get_partials": (c_int, [c_void_p, c_int, _dp]),
    "phb_get_scalers": (c_int, [c_void_p, c_int, _dp]),
    "phb_get_root_partials": (c_int, [c_void_p, _dp, _dp]),
    "phb_compute_up_partials": (c_int, [c_void_p, c_int, c_int, c_double]),
    "phb_edge_derivatives": (c_int, [c_void_p, c_int, _ip, _dp, c_int, _dp]),
    "phb_update_node": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_double]),
    "phb_branch_derivatives": (c_int, [c_void_p, c_int, c_int, c_int, _dp, c_int, _dp]),
    "phb_lnl_resident_async": (c_int, [c_void_p, c_int, c_int, c_double]),
    "phb_root_lnl_async": (c_int, [c_void_p, c_int, c_int, c_double]),
    "phb_lnl_from_host_packed_async": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double]),
    "phb_edge_derivatives_async": (c_int, [c_void_p, c_int, _ip, _dp, c_int]),
    "phb_lnl_from_host_submit": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, POINTER(c_int)]),
    "phb_result_post": (c_int, [c_void_p, c_int]),
    "phb_result_wait": (c_int, [c_void_p, c_int, _dp]),
    "phb_device_result": (c_int, [c_void_p, POINTER(c_void_p), _lp]),
    "phb_result_fetch": (c_int, [c_void_p, c_int, _dp]),
    "phb_peer_buffer": (c_int, [c_void_p, c_void_p]),
    "phb_peer_connect": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "phb_peer_sum_next": (c_int, [c_void_p]),
    "phb_op_clv": (c_int, [c_int, c_int64, c_int, c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "phb_op_lnl_node": (c_int, [c_int, c_int64, c_int, c_int, _dp, _dp, _dp, _dp]),
    "phb_op_lnl_branch": (c_int, [c_int, c_int64, c_int, c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "phb_op_pmatrices": (c_int, [c_int, c_int, c_int, _dp, _dp, _dp, _dp, c_int, _dp]),
    "phb_op_fp64_peak": (c_int, [c_int, c_int, _dp]),
}

_lib = None


class EngineError(RuntimeError):
    def __init__(self, status, message):
        RuntimeError.__init__(self, message)
        self.status = status


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "{} is missing - build it with `python phylo_utils_b200/csrc/build.py` "
                "(or __graft_entry__.build()); phylo_utils_b200 has no CPU fallback".format(LIB_PATH))
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status, ctx=None):
    if status == PHB_OK:
        return
    handle = lib()
    raw = handle.phb_last_error(ctx)
    msg = raw.decode("utf-8", "replace") if raw else ""
    name = handle.phb_status_name(status).decode()
    text = "{}: {}".format(name, msg) if msg else name
    if status in (PHB_ERR_INVALID, PHB_ERR_UNSUPPORTED):
        raise ValueError(text)
    if status == PHB_ERR_NOMEM:
        raise MemoryError(text)
    raise EngineError(status, text)


def dptr(arr):
    """numpy float64 C-contiguous array -> double*"""
    return arr.ctypes.data_as(_dp)


def iptr(arr):
    return arr.ctypes.data_as(_ip)
