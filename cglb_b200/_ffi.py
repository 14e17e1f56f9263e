"""ctypes binding of libcglb_b200.so (include/cglb_b200.h).  There is NO fallback: if the shared library
or a CUDA device is missing every product call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcglb_b200.so")

MATERN32, RBF = 0, 1
KIND_IDS = {"matern32": MATERN32, "rbf": RBF}
ROW_PAD = 128

c_dp = C.c_void_p      # device pointers travel as integers
c_long, c_int, c_dbl = C.c_long, C.c_int, C.c_double

# name -> (restype, argtypes); must list every symbol include/cglb_b200.h declares (tests check this)
SIGNATURES = {
    "cglb_abi_version": (c_int, []),
    "cglb_last_error": (C.c_char_p, []),
    "cglb_create": (c_int, [C.POINTER(C.c_void_p), c_int]),
    "cglb_destroy": (c_int, [C.c_void_p]),
    "cglb_set_option": (c_int, [C.c_void_p, C.c_char_p, c_long]),
    "cglb_launch_count": (C.c_ulonglong, [C.c_void_p]),
    "cglb_num_sms": (c_int, [C.c_void_p]),
    "cglb_packed_width": (c_int, [c_int]),
    "cglb_padded_rows": (c_long, [c_long]),
    "cglb_pack_inputs": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_int, c_dp, c_dp, c_dp, C.c_void_p]),
    "cglb_kmv_sym_variant": (c_int, [C.c_void_p, c_int, c_long, c_int]),
    "cglb_packed_width_f32": (c_int, [c_int]),
    "cglb_pack_inputs_f32": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_int, c_dp, c_dp, c_dp, C.c_void_p]),
    "cglb_kmv_sym_f32": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_int, c_dp, c_dp, c_dbl, c_dbl, c_int, c_int, C.c_void_p]),
    "cglb_kmv_bwd_sym_f32": (c_int, [C.c_void_p, c_int, c_dp, c_dp, c_long, c_int, c_dp, c_dp, c_dbl, c_dp, c_dp, c_int, c_int, C.c_void_p]),
    "cglb_kmv_sym": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_int, c_dp, c_dp, c_dbl, c_dbl, c_int, c_int, C.c_void_p]),
    "cglb_kmv_sym_multi": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_int, c_dp, c_int, c_dp, c_dbl, c_dbl, c_int, c_int, C.c_void_p]),
    "cglb_kmv_rect": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_dp, c_long, c_int, c_dp, c_dp, c_dbl, C.c_void_p]),
    "cglb_kmv_bwd_sym": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_int, c_dp, c_dp, c_dbl, c_dp, c_dp, c_int, c_int, C.c_void_p]),
    "cglb_knm_build": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_dp, c_long, c_int, c_dbl, c_dp, c_long, C.c_void_p]),
    "cglb_potrf": (c_int, [C.c_void_p, c_dp, c_long, c_long, c_dp, C.c_void_p]),
    "cglb_tri_inverse": (c_int, [C.c_void_p, c_dp, c_long, c_long, c_dp, c_long, C.c_void_p]),
    "cglb_trsm_left_lower": (c_int, [C.c_void_p, c_dp, c_long, c_long, c_dp, c_long, c_long, c_dbl, C.c_void_p]),
    "cglb_syrk": (c_int, [C.c_void_p, c_dp, c_long, c_long, c_long, c_dp, c_long, c_int, C.c_void_p]),
    "cglb_gemm": (c_int, [C.c_void_p, c_int, c_long, c_long, c_long, c_dbl, c_dp, c_long, c_dp, c_long, c_dbl, c_dp, c_long, C.c_void_p]),
    "cglb_precond_project": (c_int, [C.c_void_p, c_dp, c_long, c_long, c_long, c_dp, c_dp, C.c_void_p]),
    "cglb_precond_finish": (c_int, [C.c_void_p, c_dp, c_long, c_long, c_long, c_dp, c_dp, c_dp, c_dbl, c_dp, c_dp, c_dp, C.c_void_p]),
    "cglb_dot": (c_int, [C.c_void_p, c_dp, c_dp, c_long, c_dp, C.c_void_p]),
    "cglb_cg_step": (c_int, [C.c_void_p, c_long, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_int, C.c_void_p]),
    "cglb_residual": (c_int, [C.c_void_p, c_long, c_dp, c_dp, c_dp, C.c_void_p]),
    "cglb_cg_direction": (c_int, [C.c_void_p, c_long, c_dp, c_dp, c_dp, c_dp, c_int, C.c_void_p]),
    "cglb_quad_terms": (c_int, [C.c_void_p, c_long, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "cglb_knm_backward": (c_int, [C.c_void_p, c_int, c_dp, c_long, c_dp, c_long, c_int, c_dbl, c_dp, c_dp, c_long, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
}

_lib = None


class CglbError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree library and attach signatures.  Raises CglbError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CglbError(
            f"{LIB_PATH} not found: build it with `python -m cglb_b200.build` (or __graft_entry__.build()). "
            "cglb_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load_library().cglb_last_error()
        raise CglbError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
