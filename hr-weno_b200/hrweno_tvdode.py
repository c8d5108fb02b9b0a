"""type(rktvd) / type(mstvd) (src/hrweno_tvdode.f90) over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .fv import FV


class _tvdode:
    def __init__(self):
        self._h = C.c_void_p()
        self._cb = None
        self.msg = ""

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_ode_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:  # interpreter shutdown: module globals are already gone
            pass

    def _created(self, st):
        if st != _abi.OK:
            self.msg = _abi.lib().hrweno_last_error().decode()
            raise _abi.HrwenoError(st, self.msg)

    # public fields of type(tvdode), tvdode.f90:18-29
    @property
    def fevals(self):
        return _abi.lib().hrweno_ode_fevals(self._h)

    @property
    def istate(self):
        return _abi.lib().hrweno_ode_istate(self._h)

    @property
    def order(self):
        return _abi.lib().hrweno_ode_order(self._h)

    @property
    def neq(self):
        return _abi.lib().hrweno_ode_neq(self._h)

    @property
    def launches(self):
        return _abi.lib().hrweno_ode_launches(self._h)

    def integrate(self, u, t, tout, dt, itask=1):
        """``call ode%integrate(u, t, tout, dt[, itask])``; u (host ndarray) is updated in place, returns t."""
        if not (isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous):
            raise TypeError("u must be a contiguous float64 ndarray (updated in place)")
        tt = C.c_double(t)
        _abi.check(_abi.lib().hrweno_ode_integrate(self._h, u.ctypes.data, C.byref(tt), tout, dt, itask))
        return tt.value

    def integrate_dev(self, u_ptr, t, tout, dt, itask=1, stream=None):
        tt = C.c_double(t)
        _abi.check(_abi.lib().hrweno_ode_integrate_dev(self._h, u_ptr, C.byref(tt), tout, dt, itask, stream))
        return tt.value


    # device callers that keep the state inside the integrator between output times (fused integrators)
    def attach(self, u_ptr, stream=None):
        _abi.check(_abi.lib().hrweno_ode_attach(self._h, u_ptr, stream))

    def integrate_attached(self, t, tout, dt, itask=1, stream=None):
        tt = C.c_double(t)
        _abi.check(_abi.lib().hrweno_ode_integrate_attached(self._h, C.byref(tt), tout, dt, itask, stream))
        return tt.value

    def fetch(self, u_ptr, stream=None):
        _abi.check(_abi.lib().hrweno_ode_fetch(self._h, u_ptr, stream))


def _wrap_rhs(fu):
    def cb(_ctx, t, neq, u_ptr, udot_ptr, stream):
        fu(t, neq, u_ptr, udot_ptr, stream)

    return _abi.RHS_FN(cb)


def _wrap_host_rhs(fu):
    """the reference's integrand `fu(t, u(:), udot(:))` (tvdode.f90:50-57) on host arrays: fu(t, u, udot) fills udot"""

    def cb(_ctx, t, neq, u_ptr, udot_ptr):
        u = np.ctypeslib.as_array(u_ptr, shape=(neq,))
        udot = np.ctypeslib.as_array(udot_ptr, shape=(neq,))
        fu(t, u, udot)

    return _abi.RHS_HOST_FN(cb)


class rktvd(_tvdode):
    """``rktvd(fu, neq, order)`` (tvdode.f90:69-95).  `fu` is an FV operator (fused path), a host integrand
    fu(t, u, udot) on NumPy arrays as in the reference (``host=True``), or a callable
    fu(t, neq, u_dev_ptr, udot_dev_ptr, stream) working on device memory."""

    def __init__(self, fu, neq, order, host=False):
        super().__init__()
        if isinstance(fu, FV):
            if neq != fu.neq:
                raise _abi.HrwenoError(_abi.EINVAL, "neq does not match the FV operator")
            self._fv = fu
            self._created(_abi.lib().hrweno_rktvd_create_fused(C.byref(self._h), fu._h, order))
        elif host:
            self._cb = _wrap_host_rhs(fu)
            self._created(_abi.lib().hrweno_rktvd_create_host(C.byref(self._h), self._cb, None, neq, order))
        else:
            self._cb = _wrap_rhs(fu)
            self._created(_abi.lib().hrweno_rktvd_create(C.byref(self._h), self._cb, None, neq, order))


class mstvd(_tvdode):
    """``mstvd(fu, neq)`` (tvdode.f90:180-201)"""

    def __init__(self, fu, neq, host=False):
        super().__init__()
        if isinstance(fu, FV):
            if neq != fu.neq:
                raise _abi.HrwenoError(_abi.EINVAL, "neq does not match the FV operator")
            self._fv = fu
            self._created(_abi.lib().hrweno_mstvd_create_fused(C.byref(self._h), fu._h))
        elif host:
            self._cb = _wrap_host_rhs(fu)
            self._created(_abi.lib().hrweno_mstvd_create_host(C.byref(self._h), self._cb, None, neq))
        else:
            self._cb = _wrap_rhs(fu)
            self._created(_abi.lib().hrweno_mstvd_create(C.byref(self._h), self._cb, None, neq))

    def integrate(self, u, t, tout, dt, itask=1):  # mstvd has no itask (tvdode.f90:203)
        return super().integrate(u, t, tout, dt, 1)
