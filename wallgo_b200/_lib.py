"""ctypes binding of libwallgo_b200.so (C ABI declared in include/wallgo_b200.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 GPU is
usable when a solver is created, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwallgo_b200.so")

# every symbol include/wallgo_b200.h declares: name -> (restype, argtypes)
_vp, _dp, _ip = C.c_void_p, C.c_void_p, C.c_void_p
_hd = C.POINTER(C.c_double)  # host double*
SYMBOLS = {
    "wg_version": (C.c_int, []),
    "wg_last_error": (C.c_char_p, []),
    "wg_launch_count": (C.c_longlong, []),
    "wg_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "wg_solver_set_tuning": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "wg_set_lookahead": (C.c_int, [C.c_int]),
    "wg_set_bulk_priority_rest": (C.c_int, [C.c_int]),
    "wg_set_bulk_split": (C.c_int, [C.c_int]),
    "wg_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "wg_destroy": (C.c_int, [_vp]),
    "wg_dims": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "wg_set_grid": (C.c_int, [_vp] + [_hd] * 7),
    "wg_set_statistics": (C.c_int, [_vp, _hd]),
    "wg_build_basis": (C.c_int, [_vp]),
    "wg_set_tables": (C.c_int, [_vp] + [_hd] * 6),
    "wg_get_tables": (C.c_int, [_vp] + [_hd] * 6),
    "wg_set_transforms": (C.c_int, [_vp, C.POINTER(_hd)]),
    "wg_set_collision": (C.c_int, [_vp, _vp, C.c_int]),
    "wg_interpolate_collision": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _hd, _hd, _hd, _hd]),
    "wg_assemble": (C.c_int, [_vp, _dp, C.c_double, C.c_double, C.c_int, _dp, C.c_longlong, _dp, _vp]),
    "wg_assemble_rows": (C.c_int, [_vp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _dp, C.c_longlong, _dp, _vp]),
    "wg_factor": (C.c_int, [_vp, _dp, C.c_longlong, _vp]),
    "wg_factor_solve": (C.c_int, [_vp, _dp, C.c_longlong, _dp, _vp]),
    "wg_solve": (C.c_int, [_vp, _dp, C.c_longlong, _dp, _vp]),
    "wg_refine": (C.c_int, [_vp, _dp, C.c_double, C.c_double, _dp, C.c_longlong, _dp, _dp, _vp]),
    "wg_assemble_factor_solve_batched": (C.c_int, [C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_double),
                                                   C.POINTER(C.c_double), C.POINTER(_vp), C.POINTER(C.c_longlong),
                                                   C.POINTER(_vp), C.POINTER(_vp)]),
    "wg_info": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "wg_spectra": (C.c_int, [_vp, _dp, _dp, _vp]),
    "wg_moments": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _vp]),
    "wg_profile_gemm": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "wg_set_dofs": (C.c_int, [_vp, _hd]),
    "wg_feedback": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_double, _dp, _vp]),
    "wg_collision_apply": (C.c_int, [_vp, _dp, _dp, _vp]),
    "wg_change_basis": (C.c_int, [_vp, C.c_int, _dp, _dp, _vp]),
    "wg_batch_create": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "wg_batch_destroy": (C.c_int, [_vp]),
    "wg_batch_assemble": (C.c_int, [_vp, _dp, _dp, C.c_double, C.c_int, _dp, C.c_longlong, C.c_longlong, _dp, _vp]),
    "wg_batch_factor": (C.c_int, [_vp, _dp, C.c_longlong, C.c_longlong, _vp]),
    "wg_batch_solve": (C.c_int, [_vp, _dp, C.c_longlong, C.c_longlong, _dp, _vp]),
    "wg_batch_refine": (C.c_int, [_vp, _dp, _dp, C.c_double, _dp, C.c_longlong, C.c_longlong, _dp, _dp, _vp]),
    "wg_batch_info": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "wg_batch_spectra": (C.c_int, [_vp, _dp, _dp, _vp]),
    "wg_batch_moments": (C.c_int, [_vp, _dp, _dp, _ip, C.c_int, _dp, _dp, _dp, _vp]),
    "wg_lu_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "wg_lu_create_batched": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "wg_lu_factor_batched": (C.c_int, [_vp, _dp, C.c_longlong, C.c_longlong, _ip, _ip, _vp]),
    "wg_lu_solve_batched": (C.c_int, [_vp, _dp, C.c_longlong, C.c_longlong, _dp, C.c_longlong, _vp]),
    "wg_lu_destroy": (C.c_int, [_vp]),
    "wg_lu_factor": (C.c_int, [_vp, _dp, C.c_longlong, _ip, _ip, _vp]),
    "wg_lu_solve": (C.c_int, [_vp, _dp, C.c_longlong, _dp, _vp]),
    "wg_dgemm_sub": (C.c_int, [_dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                               C.c_longlong, _vp]),
}

_lib = None


class WallGoB200Error(RuntimeError):
    """Raised for every non-zero return code of the C ABI."""


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WallGoB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or wallgo_b200/csrc/build.sh.  wallgo_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().wg_last_error().decode(errors="replace")
        raise WallGoB200Error(f"{what} failed with code {rc}: {msg}")


def host_ptr(arr):
    """double* of a C-contiguous float64 numpy array (kept alive by the caller)."""
    return arr.ctypes.data_as(_hd)
