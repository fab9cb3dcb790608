"""ctypes binding of libnbctc.so (the C ABI declared in include/nbctc.h).

There is no CPU fallback: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libnbctc.so")

OK = 0
FLAG_DEFAULT = 0
FLAG_GENERIC = 1
FLAG_NO_GRAD = 2
FLAG_ALIGNED16 = 4
FLAG_LOCKSTEP = 8
FLAG_SEQWARP = 32
FLAG_SUM_WEIGHTED = 64

_lib = None

# name -> (restype, argtypes); must list every symbol include/nbctc.h declares
_i64, _u32, _f32, _vp, _sz, _int = C.c_int64, C.c_uint32, C.c_float, C.c_void_p, C.c_size_t, C.c_int
SIGNATURES = {
    "nbctc_version": (_int, []),
    "nbctc_last_error": (C.c_char_p, []),
    "nbctc_kernel_launch_count": (C.c_uint64, []),
    "nbctc_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64, _int, _u32]),
    "nbctc_loss_grad_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _sz, _u32, _vp]),
    "nbctc_loss_grad_lse_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _sz, _u32, _vp]),
    "nbbctc_loss_grad_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _sz, _u32, _vp]),
    "nbctc_aux_ce_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp]),
    "nbctc_scale_grad_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "nbctc_best_path_i32": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nbctc_best_path_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "nbctc_frame_topk_i32": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _int, _vp, _vp]),
    "nbctc_match_time_i32": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _i64, _i64, _int, _vp, _vp, _vp]),
    "nbctc_match_frame_i32": (_int, [_vp, _vp, _vp, _i64, _int, _i64, _vp, _vp]),
    "nbctc_host_release": (_int, [_int]),
    "nbctc_debug_set_prof": (_int, [_vp]),
    "nbctc_loss_grad_host_f32": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _u32]),
    "nbbctc_loss_grad_host_f32": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _u32]),
}


class NbctcError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libnbctc.so (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NbctcError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. "
                "Run `python -m ctc_b200.build` (needs nvcc). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = lib().nbctc_last_error().decode("utf-8", "replace")
        raise NbctcError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().nbctc_kernel_launch_count())
