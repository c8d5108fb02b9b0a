"""The REAL32 build of the path (src/hrweno_kinds.F90:9-17: rk = real32) over the C ABI (hrweno_*_f32, csrc/real32.cu).

Same names as the fp64 wrappers (weno, FV, rktvd, mstvd, make_desc); every real is a float32.  One GPU, the reference's
operation order in single precision, bit-identical to the REAL32 build of the oracle (oracle/ref32.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi

F32 = np.float32
c_float_p = C.POINTER(C.c_float)


def _f32(a):
    return np.ascontiguousarray(a, dtype=F32)


def grid1():
    """type(grid1) of the REAL32 build (grids.f90:10-37 with rk = real32): edges / center / width evaluated in float32"""
    from .hrweno_grids import grid1 as _grid1

    return _grid1(F32)


def make_desc(n, k=3, eps=1e-6, rows=1, flux_model=_abi.FLUX_BURGERS, flux_scheme=_abi.SCHEME_GODUNOV, flux_coef=(1.0, 1.0), alpha=1.0,
              bc=_abi.BC_COPY_NEIGHBOUR, width=None, linear=None):
    """hrweno_fv_desc_f32; `width` = per-axis float32 arrays (grid1 evaluated in real32), or `linear` = (xmin, xmax)"""
    n = (int(n),) if np.isscalar(n) else tuple(int(x) for x in n)
    d = _abi.FvDesc32()
    d.abi_version, d.ndim = _abi.ABI_VERSION, len(n)
    d.n[0], d.n[1] = n[0], (n[1] if len(n) > 1 else 1)
    d.rows, d.k, d.eps = rows, k, eps
    d.flux_model, d.flux_scheme, d.bc, d.mode = flux_model, flux_scheme, bc, _abi.MODE_STRICT
    d.flux_coef[0], d.flux_coef[1] = flux_coef
    d.alpha = alpha
    keep = []
    if linear is not None:
        d.grid_kind = _abi.GRID_LINEAR
        d.xmin, d.xmax = linear
    else:
        d.grid_kind = _abi.GRID_WIDTH_ARRAY
        for a in range(d.ndim):
            w = _f32(width[a])
            keep.append(w)
            d.width[a] = w.ctypes.data_as(c_float_p)
    d.rank, d.nranks, d.global_n, d.global_offset = 0, 1, 0, 0
    d._keepalive = keep
    return d


class weno:
    """``weno(ncells, k, eps[, xedges])`` / ``%reconstruct`` in real32 (weno.f90:54-219)"""

    def __init__(self, ncells, k=3, eps=1e-6, xedges=None):
        self._h = C.c_void_p()
        xe = None if xedges is None else _f32(xedges)
        if xe is not None and xe.size != ncells + 1:
            raise _abi.HrwenoError(_abi.EINVAL, "Invalid input 'xedges': size(xedges) /= ncells + 1.")
        _abi.check(_abi.lib().hrweno_weno_f32_create(C.byref(self._h), int(ncells), int(k), float(eps), None if xe is None else xe.ctypes.data))
        self.ncells, self.k = int(ncells), int(k)
        self.uniform_grid = xe is None

    @property
    def cnu(self):
        """cnu(0:k-1, -1:k-1, 1:ncells) as a float32 array [i-1, r+1, j] (weno.f90:41); None on a uniform grid (:110-112)"""
        if self.uniform_grid:
            return None
        out = np.empty((self.ncells, self.k + 1, self.k), dtype=F32)
        _abi.check(_abi.lib().hrweno_weno_f32_get_cnu(self._h, out.ctypes.data))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_weno_f32_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def reconstruct(self, v):
        v = _f32(v)
        if v.shape != (self.ncells,):
            raise ValueError("size(v) /= ncells")
        vl, vr = np.empty_like(v), np.empty_like(v)
        _abi.check(_abi.lib().hrweno_weno_f32_reconstruct(self._h, v.ctypes.data, vl.ctypes.data, vr.ctypes.data))
        return vl, vr


class FV:
    def __init__(self, desc):
        self.desc = desc
        self._h = C.c_void_p()
        _abi.check(_abi.lib().hrweno_fv_f32_create(C.byref(self._h), C.byref(desc)))
        self.neq = _abi.lib().hrweno_fv_f32_neq(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_fv_f32_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def rhs(self, t, v):
        v = _f32(v)
        out = np.empty_like(v)
        _abi.check(_abi.lib().hrweno_fv_f32_rhs(self._h, float(t), v.ctypes.data, out.ctypes.data))
        return out

    def set_xedges(self, axis, xedges):
        xe = _f32(xedges)
        if xe.size != self.desc.n[axis] + 1:
            raise _abi.HrwenoError(_abi.EINVAL, "Invalid input 'xedges'. Valid range: size(xedges) = ncells + 1.")
        _abi.check(_abi.lib().hrweno_fv_f32_set_xedges(self._h, axis, xe.ctypes.data))

    def set_flux_coef(self, axis, face=None, cross=None):
        f = None if face is None else _f32(face)
        c = None if cross is None else _f32(cross)
        _abi.check(_abi.lib().hrweno_fv_f32_set_flux_coef(self._h, axis, None if f is None else f.ctypes.data, None if c is None else c.ctypes.data))

    def set_flux_time_fn(self, g):
        self._tfn = _abi.TIME_FN32(lambda _ctx, t: float(g(F32(t)))) if g is not None else C.cast(None, _abi.TIME_FN32)
        _abi.check(_abi.lib().hrweno_fv_f32_set_flux_time_fn(self._h, self._tfn, None))


class _ode:
    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_ode_f32_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    fevals = property(lambda s: _abi.lib().hrweno_ode_f32_fevals(s._h))
    istate = property(lambda s: _abi.lib().hrweno_ode_f32_istate(s._h))
    launches = property(lambda s: _abi.lib().hrweno_ode_f32_launches(s._h))

    def integrate(self, u, t, tout, dt, itask=1):
        """u: contiguous float32 ndarray, updated in place; t, tout, dt are rounded to real32; returns t (a float32 value)"""
        if not (isinstance(u, np.ndarray) and u.dtype == F32 and u.flags.c_contiguous):
            raise TypeError("u must be a contiguous float32 ndarray (updated in place)")
        tt = C.c_float(t)
        _abi.check(_abi.lib().hrweno_ode_f32_integrate(self._h, u.ctypes.data, C.byref(tt), float(tout), float(dt), int(itask)))
        return tt.value

    def integrate_dev(self, u_ptr, t, tout, dt, itask=1, stream=None):
        tt = C.c_float(t)
        _abi.check(_abi.lib().hrweno_ode_f32_integrate_dev(self._h, u_ptr, C.byref(tt), float(tout), float(dt), int(itask), stream))
        return tt.value


class rktvd(_ode):
    def __init__(self, fv, order):
        super().__init__()
        self._fv = fv
        _abi.check(_abi.lib().hrweno_rktvd_f32_create_fused(C.byref(self._h), fv._h, int(order)))


class mstvd(_ode):
    def __init__(self, fv):
        super().__init__()
        self._fv = fv
        _abi.check(_abi.lib().hrweno_mstvd_f32_create_fused(C.byref(self._h), fv._h))

    def integrate(self, u, t, tout, dt, itask=1):
        return super().integrate(u, t, tout, dt, 1)
