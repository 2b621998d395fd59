"""ctypes loader of libnsparse_b200.so (the C ABI of include/nsparse_b200.h).

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NSP_LIB_PATH") or os.path.join(_HERE, "lib", "libnsparse_b200.so")

_lib = None

i32p = C.c_void_p   # device / host pointers are passed as plain addresses
vp = C.c_void_p
ll = C.c_longlong


class NsparseError(RuntimeError):
    pass


class nsp_amb(C.Structure):
    """Device part of sfAMB (cuda-c/inc/nsparse.h:78-107), see include/nsparse_b200.h."""
    _fields_ = [
        ("d_cs", vp), ("d_cl", vp), ("d_sellcs_col", vp), ("d_sellcs_val", vp),
        ("d_s_write_permutation", vp), ("d_s_write_permutation_offset", vp), ("d_write_permutation", vp),
        ("block_size", C.c_int), ("nnz", C.c_int), ("M", C.c_int), ("N", C.c_int), ("pad_M", C.c_int),
        ("chunk", C.c_int), ("SIGMA", C.c_int), ("c_size", C.c_int),
        ("seg_size", ll), ("seg_num", ll), ("thread_grid", ll), ("thread_block", ll),
    ]


# every symbol include/nsparse_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "nsp_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "nsp_destroy": (C.c_int, [vp]),
    "nsp_last_error": (C.c_char_p, [vp]),
    "nsp_set_stream": (C.c_int, [vp, vp]),
    "nsp_sync": (C.c_int, [vp]),
    "nsp_set_option": (C.c_int, [vp, C.c_char_p, ll]),
    "nsp_launch_count": (ll, [vp]),
    "nsp_profile_dump": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
    "nsp_spgemm_flop": (C.c_int, [vp, C.c_int, vp, vp, vp, C.POINTER(ll)]),
    "nsp_spgemm_symbolic": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                      C.POINTER(ll), C.POINTER(ll)]),
    "nsp_spgemm_numeric_s": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 9),
    "nsp_spgemm_numeric_d": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 9),
    "nsp_spgemm_numeric_rows_s": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 9),
    "nsp_spgemm_numeric_rows_d": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 9),
    "nsp_rpt64_to_rpt32": (C.c_int, [vp, C.c_int, vp, ll, vp]),
    "nsp_csr_fold_s": (C.c_int, [vp, C.c_int, ll, vp, vp, vp, C.POINTER(C.c_ulonglong * 2), C.POINTER(C.c_double * 2)]),
    "nsp_csr_fold_d": (C.c_int, [vp, C.c_int, ll, vp, vp, vp, C.POINTER(C.c_ulonglong * 2), C.POINTER(C.c_double * 2)]),
    "nsp_spgemm_host_s": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [C.POINTER(ll)]),
    "nsp_spgemm_host_d": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [C.POINTER(ll)]),
    "nsp_spgemm_host_fetch_s": (C.c_int, [vp, vp, vp, vp]),
    "nsp_spgemm_host_fetch_d": (C.c_int, [vp, vp, vp, vp]),
    "nsp_spgemm_host_drain": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_ulonglong), C.POINTER(ll)]),
    "nsp_spgemm_host_stream_s": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [vp, C.c_size_t, C.c_int, C.POINTER(ll),
                                                                                      C.POINTER(C.c_ulonglong), C.POINTER(ll)]),
    "nsp_spgemm_host_stream_d": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [vp, C.c_size_t, C.c_int, C.POINTER(ll),
                                                                                      C.POINTER(C.c_ulonglong), C.POINTER(ll)]),
    "nsp_spgemm_host_release": (C.c_int, [vp]),
    "nsp_csr2amb_s": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, ll, C.c_int, C.c_int, vp,
                                C.POINTER(nsp_amb)]),
    "nsp_csr2amb_d": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, ll, C.c_int, C.c_int, vp,
                                C.POINTER(nsp_amb)]),
    "nsp_amb_free": (C.c_int, [vp, C.POINTER(nsp_amb)]),
    "nsp_spmv_amb_s": (C.c_int, [vp, C.POINTER(nsp_amb), vp, vp]),
    "nsp_spmv_amb_d": (C.c_int, [vp, C.POINTER(nsp_amb), vp, vp]),
    "nsp_spmv_amb_host_s": (C.c_int, [vp, C.POINTER(nsp_amb), vp, vp]),
    "nsp_spmv_amb_host_d": (C.c_int, [vp, C.POINTER(nsp_amb), vp, vp]),
    "nsp_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "nsp_peer_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]),
    "nsp_peer_open": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "nsp_peer_close": (C.c_int, [vp, vp]),
    "nsp_peer_free": (C.c_int, [vp, vp]),
    "nsp_copy_async": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "nsp_spgemm_set_peers": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), ll]),
    "nsp_spgemm_peers_status": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "nsp_spgemm_peers_stats": (C.c_int, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "nsp_mgpu_create": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
    "nsp_mgpu_destroy": (C.c_int, [vp]),
    "nsp_mgpu_last_error": (C.c_char_p, [vp]),
    "nsp_mgpu_ngpu": (C.c_int, [vp]),
    "nsp_mgpu_context": (vp, [vp, C.c_int]),
    "nsp_mgpu_block": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(ll),
                                 C.POINTER(ll), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nsp_mgpu_spgemm_symbolic_s": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [C.POINTER(ll), C.POINTER(ll)]),
    "nsp_mgpu_spgemm_symbolic_d": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [C.POINTER(ll), C.POINTER(ll)]),
    "nsp_mgpu_spgemm_numeric_s": (C.c_int, [vp, vp, vp, vp]),
    "nsp_mgpu_spgemm_numeric_d": (C.c_int, [vp, vp, vp, vp]),
    "nsp_push_multicast": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t]),
    "nsp_push_to_peers": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_size_t, vp, C.c_size_t]),
    "nsp_read_mtx": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(ll),
                               C.POINTER(C.c_int), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "nsp_free_host": (None, [vp]),
    "nsp_gen_rmat_edges": (C.c_int, [C.c_int, ll, C.c_ulonglong, vp, vp]),
}


# symbols declared in the header whose implementation has not landed yet (must be empty at release)
_PENDING = set()


def load():
    """Load the shared library and bind every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NsparseError(
                f"{LIB_PATH} is not built: run `make lib` (or `python -c 'import __graft_entry__ as g; "
                "g.build()'`).  nsparse_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if name in _PENDING and not hasattr(L, name):
                continue
            fn = getattr(L, name)   # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
