"""ctypes wrapper of the REAL32 build of the CPU oracle (oracle/libhrweno_oracle_f32.so: hrweno_oracle.c compiled with
-DHRW_REAL32, i.e. rk = real32 as src/hrweno_kinds.F90:9-17 selects it).  TEST INFRASTRUCTURE, like ref.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import __graft_entry__ as graft

_abi = graft.load_package()._abi
HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhrweno_oracle_f32.so")
F32 = np.float32
_lib = None
# the callback types of hrweno_oracle.h / hrweno_b200.h as the REAL32 build reads them (every `double` a `float`)
FLUX_FN32 = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_float, C.POINTER(C.c_float), C.c_int, C.c_float)
RHS_FN32 = C.CFUNCTYPE(None, C.c_void_p, C.c_float, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make"], cwd=HERE, check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(LIB_PATH)
    vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int
    protos = {
        "hrweno_ref_set_threads": (None, [i32]),
        "hrweno_ref_weno_calc_cnu": (i32, [i64, i32, vp, vp]),
        "hrweno_ref_weno_reconstruct": (i32, [i64, i32, f32, vp, vp, i64, vp, vp]),
        "hrweno_ref_fv_create": (i32, [C.POINTER(vp), C.POINTER(_abi.FvDesc32)]),
        "hrweno_ref_fv_destroy": (None, [vp]),
        "hrweno_ref_fv_neq": (i64, [vp]),
        "hrweno_ref_fv_rhs": (i32, [vp, f32, vp, vp]),
        "hrweno_ref_fv_set_xedges": (i32, [vp, i32, vp]),
        "hrweno_ref_fv_set_flux_coef": (i32, [vp, i32, vp, vp]),
        "hrweno_ref_fv_set_flux_time_fn": (i32, [vp, _abi.TIME_FN32, vp]),
        "hrweno_ref_rktvd_create_fv": (i32, [C.POINTER(vp), vp, i32]),
        "hrweno_ref_mstvd_create_fv": (i32, [C.POINTER(vp), vp]),
        "hrweno_ref_ode_destroy": (None, [vp]),
        "hrweno_ref_ode_integrate": (i32, [vp, vp, C.POINTER(f32), f32, f32, i32]),
        "hrweno_ref_ode_fevals": (i64, [vp]),
        "hrweno_ref_ode_istate": (i32, [vp]),
        "hrweno_ref_lax_friedrichs": (f32, [FLUX_FN32, vp, f32, f32, vp, i32, f32, f32]),
        "hrweno_ref_godunov": (f32, [FLUX_FN32, vp, f32, f32, vp, i32, f32]),
        "hrweno_ref_face_flux": (f32, [i32, i32, f32, f32, f32, f32]),
        "hrweno_ref_rktvd_create": (i32, [C.POINTER(vp), RHS_FN32, vp, i64, i32]),
        "hrweno_ref_mstvd_create": (i32, [C.POINTER(vp), RHS_FN32, vp, i64]),
        "hrweno_ref_is_done": (i32, [f32, f32, f32]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _ok(st):
    if st != 0:
        raise _abi.HrwenoError(st, "oracle (real32): invalid input (the reference would error stop)")


def _f32(a):
    return np.ascontiguousarray(a, dtype=F32)


def calc_cnu(xedges, k):
    xe = _f32(xedges)
    nc = xe.size - 1
    out = np.empty((nc, k + 1, k), dtype=F32)
    _ok(lib().hrweno_ref_weno_calc_cnu(nc, k, xe.ctypes.data, out.ctypes.data))
    return out


def reconstruct(v, k=3, eps=1e-6, cnu=None):
    v = _f32(v)
    vl, vr = np.empty_like(v), np.empty_like(v)
    c = None if cnu is None else _f32(cnu)
    _ok(lib().hrweno_ref_weno_reconstruct(v.size, k, eps, None if c is None else c.ctypes.data, v.ctypes.data, 1, vl.ctypes.data, vr.ctypes.data))
    return vl, vr


class FV:
    def __init__(self, desc):
        self.desc = desc
        self._h = C.c_void_p()
        _ok(lib().hrweno_ref_fv_create(C.byref(self._h), C.byref(desc)))
        self.neq = lib().hrweno_ref_fv_neq(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().hrweno_ref_fv_destroy(self._h)
            self._h = C.c_void_p()

    def rhs(self, t, v):
        v = _f32(v)
        out = np.empty_like(v)
        _ok(lib().hrweno_ref_fv_rhs(self._h, float(t), v.ctypes.data, out.ctypes.data))
        return out

    def set_xedges(self, axis, xedges):
        xe = _f32(xedges)
        _ok(lib().hrweno_ref_fv_set_xedges(self._h, axis, xe.ctypes.data))

    def set_flux_coef(self, axis, face=None, cross=None):
        f = None if face is None else _f32(face)
        c = None if cross is None else _f32(cross)
        _ok(lib().hrweno_ref_fv_set_flux_coef(self._h, axis, None if f is None else f.ctypes.data, None if c is None else c.ctypes.data))

    def set_flux_time_fn(self, g):
        self._tfn = _abi.TIME_FN32(lambda _ctx, t: float(g(F32(t)))) if g is not None else C.cast(None, _abi.TIME_FN32)
        _ok(lib().hrweno_ref_fv_set_flux_time_fn(self._h, self._tfn, None))


class _ode:
    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().hrweno_ref_ode_destroy(self._h)
            self._h = C.c_void_p()

    fevals = property(lambda s: lib().hrweno_ref_ode_fevals(s._h))

    def integrate(self, u, t, tout, dt, itask=1):
        assert isinstance(u, np.ndarray) and u.dtype == F32 and u.flags.c_contiguous
        tt = C.c_float(t)
        _ok(lib().hrweno_ref_ode_integrate(self._h, u.ctypes.data, C.byref(tt), float(tout), float(dt), int(itask)))
        return tt.value


def _wrap_rhs(fu):
    def cb(_ctx, t, neq, up, dp):
        u = np.ctypeslib.as_array(up, shape=(neq,))
        d = np.ctypeslib.as_array(dp, shape=(neq,))
        d[:] = fu(F32(t), u)

    return RHS_FN32(cb)


class rktvd(_ode):
    """rktvd(fu, order): fu is an oracle FV or a callable fu(t, u) -> udot on float32 (host)"""

    def __init__(self, fv, order, neq=None):
        super().__init__()
        if isinstance(fv, FV):
            self._fv = fv
            _ok(lib().hrweno_ref_rktvd_create_fv(C.byref(self._h), fv._h, order))
        else:
            self._cb = _wrap_rhs(fv)
            _ok(lib().hrweno_ref_rktvd_create(C.byref(self._h), self._cb, None, neq, order))


class mstvd(_ode):
    def __init__(self, fv, neq=None):
        super().__init__()
        if isinstance(fv, FV):
            self._fv = fv
            _ok(lib().hrweno_ref_mstvd_create_fv(C.byref(self._h), fv._h))
        else:
            self._cb = _wrap_rhs(fv)
            _ok(lib().hrweno_ref_mstvd_create(C.byref(self._h), self._cb, None, neq))


def _wrap_flux(f):
    return FLUX_FN32(lambda _ctx, u, xptr, nx, t: float(f(F32(u), np.ctypeslib.as_array(xptr, shape=(nx,)), F32(t))))


def lax_friedrichs(f, vm, vp, x, t, alpha):
    x = _f32(np.atleast_1d(x))
    return F32(lib().hrweno_ref_lax_friedrichs(_wrap_flux(f), None, vm, vp, x.ctypes.data, x.size, t, alpha))


def godunov(f, vm, vp, x, t):
    x = _f32(np.atleast_1d(x))
    return F32(lib().hrweno_ref_godunov(_wrap_flux(f), None, vm, vp, x.ctypes.data, x.size, t))


def face_flux(scheme, model, coef, alpha, vm, vp):
    f = lib().hrweno_ref_face_flux
    return np.array([f(scheme, model, coef, alpha, a, b) for a, b in zip(np.ravel(vm), np.ravel(vp))], dtype=F32)


def is_done(t, tout, dt):
    return bool(lib().hrweno_ref_is_done(t, tout, dt))
