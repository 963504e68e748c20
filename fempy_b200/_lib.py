"""ctypes binding of libfemb200.so (the C ABI declared in include/femb200.h).

There is no fallback: if the shared library is missing this module raises, and every compute call
fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

from ._build import LIB_PATH

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class Material(C.Structure):
    """Mirror of `femb_material`."""

    _fields_ = [("E", C.c_double), ("nu", C.c_double), ("rho", C.c_double), ("model", C.c_int32), ("nonlinear", C.c_int32)]


class FembError(RuntimeError):
    pass


_SIGNATURES = {
    "femb_last_error": (C.c_char_p, []),
    "femb_version": (C.c_char_p, []),
    "femb_device_count": (C.c_int, []),
    "femb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "femb_host_free": (C.c_int, [C.c_void_p]),
    "femb_plan_create": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "femb_plan_add_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p, C.c_int64, c_int64_p]),
    "femb_plan_add_block_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_void_p]),
    "femb_plan_finalize": (C.c_int, [C.c_void_p]),
    "femb_plan_destroy": (None, [C.c_void_p]),
    "femb_plan_set_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "femb_plan_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "femb_plan_get_info": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "femb_plan_sync": (C.c_int, [C.c_void_p]),
    "femb_plan_num_dof": (C.c_int64, [C.c_void_p]),
    "femb_plan_num_elems": (C.c_int64, [C.c_void_p]),
    "femb_plan_nnz": (C.c_int64, [C.c_void_p]),
    "femb_plan_build_ms": (C.c_double, [C.c_void_p]),
    "femb_plan_last_launches": (C.c_int64, [C.c_void_p]),
    "femb_plan_pattern": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "femb_plan_pattern_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "femb_assemble": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.POINTER(Material), c_int64_p, c_double_p,
                                C.c_int64, C.c_int32, c_double_p, C.c_int32, c_double_p, c_double_p]),
    "femb_plan_set_geometry": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "femb_plan_set_bcs": (C.c_int, [C.c_void_p, c_int64_p, c_double_p, C.c_int64]),
    "femb_assemble_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Material), C.c_int32,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "femb_plan_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "femb_plan_profile_read": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "femb_plan_row_stats": (C.c_int, [C.c_void_p, C.c_int32, c_int64_p, C.c_int64]),
    "femb_function": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.POINTER(Material), C.c_int32, C.c_int32, c_double_p]),
    "femb_point_quantities": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_double_p]),
    "femb_function_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Material), C.c_int32, C.c_int32, C.c_void_p]),
    "femb_solve_pcg": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int32, C.POINTER(C.c_int32), c_double_p]),
    "femb_solve_pcg_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.POINTER(C.c_int32),
                                     c_double_p]),
    "femb_assemble_solve": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.POINTER(Material), c_int64_p, c_double_p,
                                      C.c_int64, c_double_p, C.c_double, C.c_int32, c_double_p, c_double_p, C.POINTER(C.c_int32),
                                      c_double_p]),
    "femb_element_blocks": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.POINTER(Material), c_double_p, c_double_p]),
    "femb_local_matrices_to_coo": (C.c_int, [c_double_p, c_int64_p, C.c_int64, C.c_int32, C.c_int32, c_int64_p, c_int64_p,
                                             c_double_p, c_int64_p]),
    "femb_scatter_local_residuals": (C.c_int, [c_double_p, c_int64_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                               c_double_p]),
}  # fmt: skip

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load libfemb200.so (raises if it has not been built -- there is no CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FEMB200_LIB", LIB_PATH)  # measurement builds of the same library (tools/build_variant.sh)
    if not os.path.exists(path):
        raise FembError(
            f"{path} not found: build it with `python -m fempy_b200._build` (or __graft_entry__.build()). "
            "fempy_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise FembError(load().femb_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int64_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
