"""ctypes binding of ``include/snowtri.h``.  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as ct
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsnowtri.so")

OK, E_ARG, E_CUDA, E_UNSUPPORTED, E_NOMEM = 0, -1, -2, -3, -4
PREC_F64, PREC_F32, PREC_MIXED, PREC_F32_EXPERIMENTAL = 0, 1, 2, 3

# Every symbol include/snowtri.h declares: name -> (restype, argtypes)
_P, _I, _D = ct.c_void_p, ct.c_int, ct.c_double
SYMBOLS = {
    "snowtri_create": (_I, [ct.POINTER(_P), _I, _I, _P, _P, _P]),
    "snowtri_destroy": (_I, [_P]),
    "snowtri_set_params": (_I, [_P, _D, _D, _D, _D, _I, _D, _I]),
    "snowtri_set_precision": (_I, [_P, _I]),
    "snowtri_set_tuning": (_I, [_P, _I, _I, _I]),
    "snowtri_set_pipeline": (_I, [_P, _I]),
    "snowtri_set_general_kernels": (_I, [_P, _I]),
    "snowtri_set_jit": (_I, [_P, _I]),
    "snowtri_jit_status": (ct.c_char_p, [_P]),
    "snowtri_run": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "snowtri_run_host": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "snowtri_candidates": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "snowtri_condense": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "snowtri_skew_ray": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "snowtri_smooth_create": (_I, [_P, ct.POINTER(_P), _I, _I, _D, _D, _D]),
    "snowtri_smooth_destroy": (_I, [_P]),
    "snowtri_smooth_reset": (_I, [_P, _P, _P]),
    "snowtri_smooth_set_chunked": (_I, [_P, _I]),
    "snowtri_smooth_run": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _D, _P]),
    "snowtri_smooth_run_f64": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _D, _P]),
    "snowtri_pack_ragged": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "snowtri_dlt_run": (_I, [_P, _P, _P, _I, _I, _P, _I, _P]),
    "snowtri_blender_run": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "snowtri_blender_run_f64": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "snowtri_blender_smooth_create": (_I, [_P, ct.POINTER(_P), _I, _P]),
    "snowtri_blender_smooth_destroy": (_I, [_P]),
    "snowtri_blender_smooth_reset": (_I, [_P, _P, _P]),
    "snowtri_blender_smooth_set_chunked": (_I, [_P, _I]),
    "snowtri_blender_smooth_run": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _D, _P]),
    "snowtri_blender_smooth_run_f64": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _D, _P]),
    "snowtri_clip_run": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _D, _P]),
    "snowtri_comm_unique_id": (_I, [_P]),
    "snowtri_comm_init": (_I, [_P, _P, _I, _I]),
    "snowtri_comm_destroy": (_I, [_P]),
    "snowtri_allgather": (_I, [_P, _P, _P, ct.c_size_t, _P, _P]),
    "snowtri_last_error": (ct.c_char_p, [_P]),
    "snowtri_launch_count": (ct.c_longlong, [_P]),
    "snowtri_last_launch_info": (_I, [_P, ct.POINTER(_I), ct.POINTER(_I), ct.POINTER(_I), ct.POINTER(_I)]),
    "snowtri_last_kernel": (ct.c_char_p, [_P]),
    "snowtri_version": (_I, []),
}

_lib = None


class SnowtriError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"snowtri error {code}: {message}")
        self.code = code


def load():
    """Load libsnowtri.so (built in-tree by ``snowmocap_b200.build``).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "snowmocap_b200 has no CPU or PyTorch fallback.")
        lib = ct.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the library does not export the symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, handle=None):
    if rc != OK:
        msg = load().snowtri_last_error(handle)
        raise SnowtriError(rc, msg.decode() if msg else "unknown")
