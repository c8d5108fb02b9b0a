"""ctypes view of include/hrweno_b200.h and the loader of libhrweno_b200.so.

The shared library is the product; this file only declares its C ABI.  Loading
fails loudly when the library has not been built -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("HRWENO_B200_LIB") or os.path.join(PKG_DIR, "lib", "libhrweno_b200.so")  # override: tuning variants

ABI_VERSION = 1
OK, EINVAL, ECUDA, ENOMEM, ESTATE, ECOMM = range(6)

FLUX_BURGERS, FLUX_LINEAR = 0, 1
SCHEME_GODUNOV, SCHEME_LAX_FRIEDRICHS = 0, 1
BC_COPY_NEIGHBOUR, BC_ZERO_FLUX = 0, 1
GRID_WIDTH_ARRAY, GRID_LINEAR = 0, 1
MODE_STRICT, MODE_FAST = 0, 1
IPC_HANDLE_BYTES = 64

c_double_p = C.POINTER(C.c_double)


class FvDesc(C.Structure):
    """struct hrweno_fv_desc (include/hrweno_b200.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("ndim", C.c_int32),
        ("n", C.c_int64 * 2),
        ("rows", C.c_int64),
        ("k", C.c_int32),
        ("flux_model", C.c_int32),
        ("flux_scheme", C.c_int32),
        ("bc", C.c_int32),
        ("grid_kind", C.c_int32),
        ("mode", C.c_int32),
        ("eps", C.c_double),
        ("flux_coef", C.c_double * 2),
        ("alpha", C.c_double),
        ("xmin", C.c_double),
        ("xmax", C.c_double),
        ("width", c_double_p * 2),
        ("rank", C.c_int32),
        ("nranks", C.c_int32),
        ("global_n", C.c_int64),
        ("global_offset", C.c_int64),
    ]


class FvDesc32(C.Structure):
    """struct hrweno_fv_desc_f32 (include/hrweno_b200.h): the descriptor of the REAL32 build (hrweno_kinds.F90:9-17)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("ndim", C.c_int32),
        ("n", C.c_int64 * 2),
        ("rows", C.c_int64),
        ("k", C.c_int32),
        ("flux_model", C.c_int32),
        ("flux_scheme", C.c_int32),
        ("bc", C.c_int32),
        ("grid_kind", C.c_int32),
        ("mode", C.c_int32),
        ("eps", C.c_float),
        ("flux_coef", C.c_float * 2),
        ("alpha", C.c_float),
        ("xmin", C.c_float),
        ("xmax", C.c_float),
        ("width", C.POINTER(C.c_float) * 2),
        ("rank", C.c_int32),
        ("nranks", C.c_int32),
        ("global_n", C.c_int64),
        ("global_offset", C.c_int64),
    ]


TIME_FN32 = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_float)
FLUX_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double, c_double_p, C.c_int, C.c_double)
RHS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p)
RHS_HOST_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_int64, c_double_p, c_double_p)
TIME_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double)

# name -> (restype, argtypes); this table is also what the symbol-export test walks
PROTOTYPES = {
    "hrweno_abi_version": (C.c_int, []),
    "hrweno_last_error": (C.c_char_p, []),
    "hrweno_device_count": (C.c_int, []),
    "hrweno_weno_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int, C.c_double, C.c_void_p]),
    "hrweno_weno_destroy": (None, [C.c_void_p]),
    "hrweno_weno_info": (
        C.c_int,
        [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)],
    ),
    "hrweno_weno_get_cnu": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hrweno_weno_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "hrweno_weno_reconstruct": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_weno_reconstruct_s": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_weno_reconstruct_batch": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64],
    ),
    "hrweno_weno_reconstruct_dev": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "hrweno_lax_friedrichs": (
        C.c_double,
        [FLUX_FN, C.c_void_p, C.c_double, C.c_double, c_double_p, C.c_int, C.c_double, C.c_double],
    ),
    "hrweno_godunov": (C.c_double, [FLUX_FN, C.c_void_p, C.c_double, C.c_double, c_double_p, C.c_int, C.c_double]),
    "hrweno_flux_faces": (
        C.c_int,
        [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "hrweno_fv_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(FvDesc)]),
    "hrweno_fv_destroy": (None, [C.c_void_p]),
    "hrweno_fv_neq": (C.c_int64, [C.c_void_p]),
    "hrweno_fv_rhs": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "hrweno_fv_rhs_dev": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_fv_max_wavespeed_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_fv_set_alpha": (C.c_int, [C.c_void_p, C.c_double]),
    "hrweno_fv_set_xedges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hrweno_fv_set_flux_coef": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hrweno_fv_set_flux_time_fn": (C.c_int, [C.c_void_p, TIME_FN, C.c_void_p]),
    "hrweno_fv_export_halo": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hrweno_fv_import_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_fv_halo_status": (C.c_int, [C.c_void_p]),
    "hrweno_rktvd_create": (C.c_int, [C.POINTER(C.c_void_p), RHS_FN, C.c_void_p, C.c_int64, C.c_int]),
    "hrweno_mstvd_create": (C.c_int, [C.POINTER(C.c_void_p), RHS_FN, C.c_void_p, C.c_int64]),
    "hrweno_rktvd_create_host": (C.c_int, [C.POINTER(C.c_void_p), RHS_HOST_FN, C.c_void_p, C.c_int64, C.c_int]),
    "hrweno_mstvd_create_host": (C.c_int, [C.POINTER(C.c_void_p), RHS_HOST_FN, C.c_void_p, C.c_int64]),
    "hrweno_rktvd_create_fused": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]),
    "hrweno_mstvd_create_fused": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p]),
    "hrweno_ode_destroy": (None, [C.c_void_p]),
    "hrweno_ode_integrate": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int],
    ),
    "hrweno_ode_integrate_dev": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int, C.c_void_p],
    ),
    "hrweno_ode_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_ode_integrate_attached": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "hrweno_ode_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_ode_fevals": (C.c_int64, [C.c_void_p]),
    "hrweno_ode_istate": (C.c_int, [C.c_void_p]),
    "hrweno_ode_order": (C.c_int, [C.c_void_p]),
    "hrweno_ode_neq": (C.c_int64, [C.c_void_p]),
    "hrweno_ode_launches": (C.c_int64, [C.c_void_p]),
    "hrweno_mgpu_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(FvDesc), C.c_int, C.POINTER(C.c_int)]),
    "hrweno_mgpu_destroy": (None, [C.c_void_p]),
    "hrweno_mgpu_ngpus": (C.c_int, [C.c_void_p]),
    "hrweno_mgpu_slab": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hrweno_mgpu_set_xedges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hrweno_mgpu_set_flux_coef": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hrweno_mgpu_set_flux_time_fn": (C.c_int, [C.c_void_p, TIME_FN, C.c_void_p]),
    "hrweno_mgpu_rktvd": (C.c_int, [C.c_void_p, C.c_int]),
    "hrweno_mgpu_mstvd": (C.c_int, [C.c_void_p]),
    "hrweno_mgpu_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int]),
    "hrweno_mgpu_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hrweno_mgpu_integrate_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int]),
    "hrweno_mgpu_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hrweno_mgpu_max_wavespeed": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "hrweno_mgpu_set_alpha": (C.c_int, [C.c_void_p, C.c_double]),
    "hrweno_mgpu_fevals": (C.c_int64, [C.c_void_p]),
    "hrweno_mgpu_launches": (C.c_int64, [C.c_void_p]),
    "hrweno_weno_f32_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int, C.c_float, C.c_void_p]),
    "hrweno_weno_f32_destroy": (None, [C.c_void_p]),
    "hrweno_weno_f32_reconstruct": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_weno_f32_reconstruct_s": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_weno_f32_get_cnu": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hrweno_weno_f32_reconstruct_dev": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "hrweno_fv_f32_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(FvDesc32)]),
    "hrweno_fv_f32_destroy": (None, [C.c_void_p]),
    "hrweno_fv_f32_neq": (C.c_int64, [C.c_void_p]),
    "hrweno_fv_f32_rhs": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "hrweno_fv_f32_rhs_dev": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hrweno_fv_f32_set_xedges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hrweno_fv_f32_set_flux_coef": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hrweno_fv_f32_set_flux_time_fn": (C.c_int, [C.c_void_p, TIME_FN32, C.c_void_p]),
    "hrweno_rktvd_f32_create_fused": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]),
    "hrweno_mstvd_f32_create_fused": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p]),
    "hrweno_ode_f32_destroy": (None, [C.c_void_p]),
    "hrweno_ode_f32_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int]),
    "hrweno_ode_f32_integrate_dev": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_void_p],
    ),
    "hrweno_ode_f32_fevals": (C.c_int64, [C.c_void_p]),
    "hrweno_ode_f32_istate": (C.c_int, [C.c_void_p]),
    "hrweno_ode_f32_launches": (C.c_int64, [C.c_void_p]),
    "hrweno_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "hrweno_host_free": (None, [C.c_void_p]),
}

_lib = None


class HrwenoError(RuntimeError):
    """Raised where the reference would `error stop` (or on a CUDA failure)."""

    def __init__(self, status, msg):
        super().__init__(f"hrweno status {status}: {msg}")
        self.status = status
        self.msg = msg


def lib():
    """Load libhrweno_b200.so (once) and attach the prototypes.  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if handle.hrweno_abi_version() != ABI_VERSION:
        raise ImportError("libhrweno_b200.so ABI version mismatch")
    _lib = handle
    return _lib


def check(status):
    if status != OK:
        raise HrwenoError(status, lib().hrweno_last_error().decode())
