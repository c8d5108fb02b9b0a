"""godunov / lax_friedrichs (src/hrweno_fluxes.f90) over the C ABI."""
from __future__ import annotations

import numpy as np

from . import _abi


def _wrap(f):
    def cb(_ctx, u, xptr, nx, t):
        return float(f(u, np.ctypeslib.as_array(xptr, shape=(nx,)), t))

    return _abi.FLUX_FN(cb)


def lax_friedrichs(f, vm, vp, x, t, alpha):
    """fluxes.f90:22-45 -- pointwise, user flux callback f(u, x(:), t)"""
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    xp = x.ctypes.data_as(_abi.c_double_p)
    return _abi.lib().hrweno_lax_friedrichs(_wrap(f), None, vm, vp, xp, x.size, t, alpha)


def godunov(f, vm, vp, x, t):
    """fluxes.f90:47-76"""
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    xp = x.ctypes.data_as(_abi.c_double_p)
    return _abi.lib().hrweno_godunov(_wrap(f), None, vm, vp, xp, x.size, t)


def flux_faces(scheme, model, vm, vp, coef=1.0, alpha=1.0):
    """closed-set device evaluation of one numerical flux per face"""
    vm = np.ascontiguousarray(vm, dtype=np.float64)
    vp = np.ascontiguousarray(vp, dtype=np.float64)
    h = np.empty_like(vm)
    _abi.check(
        _abi.lib().hrweno_flux_faces(scheme, model, coef, alpha, vm.size, vm.ctypes.data, vp.ctypes.data, h.ctypes.data)
    )
    return h
