"""ctypes wrapper of oracle/libhrweno_oracle.so (TEST INFRASTRUCTURE -- see hrweno_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  It borrows the descriptor struct from the product's ABI module so the
same hrweno_fv_desc drives both sides.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhrweno_oracle.so")

_abi = sys.modules["hrweno_b200"]._abi if "hrweno_b200" in sys.modules else None
if _abi is None:  # pragma: no cover - load through __graft_entry__.load_oracle()
    sys.path.insert(0, os.path.dirname(HERE))
    import __graft_entry__ as _g

    _abi = _g.load_package()._abi

REF_RHS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double))

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make"], cwd=HERE, check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(LIB_PATH)
    vp, i64, dbl, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    protos = {
        "hrweno_ref_set_threads": (None, [i32]),
        "hrweno_ref_max_threads": (i32, []),
        "hrweno_ref_tables": (i32, [i32, vp, vp]),
        "hrweno_ref_weno_calc_cnu": (i32, [i64, i32, vp, vp]),
        "hrweno_ref_weno_reconstruct": (i32, [i64, i32, dbl, vp, vp, i64, vp, vp]),
        "hrweno_ref_weno_check": (i32, [i64, i32, dbl]),
        "hrweno_ref_lax_friedrichs": (dbl, [_abi.FLUX_FN, vp, dbl, dbl, _abi.c_double_p, i32, dbl, dbl]),
        "hrweno_ref_godunov": (dbl, [_abi.FLUX_FN, vp, dbl, dbl, _abi.c_double_p, i32, dbl]),
        "hrweno_ref_flux_model": (dbl, [i32, dbl, dbl]),
        "hrweno_ref_face_flux": (dbl, [i32, i32, dbl, dbl, dbl, dbl]),
        "hrweno_ref_grid_linear": (None, [dbl, dbl, i64, vp, vp, vp]),
        "hrweno_ref_fv_create": (i32, [C.POINTER(vp), C.POINTER(_abi.FvDesc)]),
        "hrweno_ref_fv_destroy": (None, [vp]),
        "hrweno_ref_fv_neq": (i64, [vp]),
        "hrweno_ref_fv_rhs": (i32, [vp, dbl, vp, vp]),
        "hrweno_ref_fv_set_xedges": (i32, [vp, i32, vp]),
        "hrweno_ref_fv_set_flux_coef": (i32, [vp, i32, vp, vp]),
        "hrweno_ref_fv_set_flux_time_fn": (i32, [vp, _abi.TIME_FN, vp]),
        "hrweno_ref_rktvd_create": (i32, [C.POINTER(vp), REF_RHS_FN, vp, i64, i32]),
        "hrweno_ref_mstvd_create": (i32, [C.POINTER(vp), REF_RHS_FN, vp, i64]),
        "hrweno_ref_rktvd_create_fv": (i32, [C.POINTER(vp), vp, i32]),
        "hrweno_ref_mstvd_create_fv": (i32, [C.POINTER(vp), vp]),
        "hrweno_ref_ode_destroy": (None, [vp]),
        "hrweno_ref_ode_integrate": (i32, [vp, vp, C.POINTER(dbl), dbl, dbl, i32]),
        "hrweno_ref_ode_fevals": (i64, [vp]),
        "hrweno_ref_ode_rhs_calls": (i64, [vp]),
        "hrweno_ref_ode_istate": (i32, [vp]),
        "hrweno_ref_is_done": (i32, [dbl, dbl, dbl]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    L.hrweno_ref_test_ode_rhs.restype = None
    _lib = L
    return L


def _ok(st):
    if st != 0:
        raise _abi.HrwenoError(st, "oracle: invalid input (the reference would error stop)")


def set_threads(n):
    lib().hrweno_ref_set_threads(int(n))


def max_threads():
    return lib().hrweno_ref_max_threads()


def tables(k):
    d, c = np.empty(k), np.empty(k * (k + 1))
    _ok(lib().hrweno_ref_tables(k, d.ctypes.data, c.ctypes.data))
    return d, c.reshape(k + 1, k)  # [r+1, j]


def calc_cnu(xedges, k):
    xe = np.ascontiguousarray(xedges, dtype=np.float64)
    nc = xe.size - 1
    out = np.empty((nc, k + 1, k))
    _ok(lib().hrweno_ref_weno_calc_cnu(nc, k, xe.ctypes.data, out.ctypes.data))
    return out


def reconstruct(v, k=3, eps=1e-6, cnu=None, incv=1):
    v = np.asarray(v, dtype=np.float64)
    if incv == 1:
        v = np.ascontiguousarray(v)
        nc = v.size
        ptr = v.ctypes.data
    else:
        nc = (v.size + incv - 1) // incv
        ptr = v.ctypes.data
    vl, vr = np.empty(nc), np.empty(nc)
    cp = None
    if cnu is not None:
        cnu = np.ascontiguousarray(cnu, dtype=np.float64)
        cp = cnu.ctypes.data
    _ok(lib().hrweno_ref_weno_reconstruct(nc, k, eps, cp, ptr, incv, vl.ctypes.data, vr.ctypes.data))
    return vl, vr


def _wrap_flux(f):
    def cb(_ctx, u, xptr, nx, t):
        return float(f(u, np.ctypeslib.as_array(xptr, shape=(nx,)), t))

    return _abi.FLUX_FN(cb)


def lax_friedrichs(f, vm, vp, x, t, alpha):
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    return lib().hrweno_ref_lax_friedrichs(_wrap_flux(f), None, vm, vp, x.ctypes.data_as(_abi.c_double_p), x.size, t, alpha)


def godunov(f, vm, vp, x, t):
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    return lib().hrweno_ref_godunov(_wrap_flux(f), None, vm, vp, x.ctypes.data_as(_abi.c_double_p), x.size, t)


def face_flux(scheme, model, coef, alpha, vm, vp):
    f = lib().hrweno_ref_face_flux
    return np.array([f(scheme, model, coef, alpha, a, b) for a, b in zip(np.ravel(vm), np.ravel(vp))])


def grid_linear(xmin, xmax, n):
    e, c, w = np.empty(n + 1), np.empty(n), np.empty(n)
    lib().hrweno_ref_grid_linear(xmin, xmax, n, e.ctypes.data, c.ctypes.data, w.ctypes.data)
    return e, c, w


def is_done(t, tout, dt):
    return bool(lib().hrweno_ref_is_done(t, tout, dt))


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
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(v)
        _ok(lib().hrweno_ref_fv_rhs(self._h, t, v.ctypes.data, out.ctypes.data))
        return out

    def set_xedges(self, axis, xedges):
        xe = np.ascontiguousarray(xedges, dtype=np.float64)
        assert xe.size == self.desc.n[axis] + 1
        _ok(lib().hrweno_ref_fv_set_xedges(self._h, axis, xe.ctypes.data))

    def set_flux_coef(self, axis, face=None, cross=None):
        f = None if face is None else np.ascontiguousarray(face, dtype=np.float64)
        c = None if cross is None else np.ascontiguousarray(cross, dtype=np.float64)
        assert f is None or f.size == self.desc.n[axis] + 1
        assert c is None or c.size == self.desc.n[1 - axis]
        _ok(lib().hrweno_ref_fv_set_flux_coef(self._h, axis, None if f is None else f.ctypes.data,
                                             None if c is None else c.ctypes.data))


    def set_flux_time_fn(self, g):
        self._tfn = _abi.TIME_FN(lambda _ctx, t: float(g(t))) if g is not None else C.cast(None, _abi.TIME_FN)
        _ok(lib().hrweno_ref_fv_set_flux_time_fn(self._h, self._tfn, None))


class _ode:
    def __init__(self):
        self._h = C.c_void_p()
        self._cb = None

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().hrweno_ref_ode_destroy(self._h)
            self._h = C.c_void_p()

    fevals = property(lambda s: lib().hrweno_ref_ode_fevals(s._h))
    rhs_calls = property(lambda s: lib().hrweno_ref_ode_rhs_calls(s._h))
    istate = property(lambda s: lib().hrweno_ref_ode_istate(s._h))

    def integrate(self, u, t, tout, dt, itask=1):
        assert isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous
        tt = C.c_double(t)
        _ok(lib().hrweno_ref_ode_integrate(self._h, u.ctypes.data, C.byref(tt), tout, dt, itask))
        return tt.value


def _wrap_rhs(fu):
    def cb(_ctx, t, neq, up, dp):
        u = np.ctypeslib.as_array(up, shape=(neq,))
        d = np.ctypeslib.as_array(dp, shape=(neq,))
        d[:] = fu(t, u)

    return REF_RHS_FN(cb)


class rktvd(_ode):
    """rktvd(fu, order): fu is an oracle FV, the string 'test_ode', or a callable fu(t, u)->udot (host)."""

    def __init__(self, fu, order, neq=None):
        super().__init__()
        if isinstance(fu, FV):
            self._fv = fu
            _ok(lib().hrweno_ref_rktvd_create_fv(C.byref(self._h), fu._h, order))
        else:
            self._cb = C.cast(lib().hrweno_ref_test_ode_rhs, REF_RHS_FN) if fu == "test_ode" else _wrap_rhs(fu)
            _ok(lib().hrweno_ref_rktvd_create(C.byref(self._h), self._cb, None, neq, order))


class mstvd(_ode):
    def __init__(self, fu, neq=None):
        super().__init__()
        if isinstance(fu, FV):
            self._fv = fu
            _ok(lib().hrweno_ref_mstvd_create_fv(C.byref(self._h), fu._h))
        else:
            self._cb = C.cast(lib().hrweno_ref_test_ode_rhs, REF_RHS_FN) if fu == "test_ode" else _wrap_rhs(fu)
            _ok(lib().hrweno_ref_mstvd_create(C.byref(self._h), self._cb, None, neq))
