"""One process, N GPUs: ctypes view of the hrweno_mgpu_* entry points (include/hrweno_b200.h; csrc/mgpu.cu).

The reference program is a single process; this mirrors `ode = rktvd(rhs, neq, order)` /
`call ode%integrate(u, t, tout, dt)` (tvdode.f90:69-178) for a problem spread over the GPUs of one box, with the
decomposition, peer access, halo mailboxes and the alpha max-reduction done inside the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


class MultiGPU:
    """`desc` is the GLOBAL problem (fv.make_desc); slabs along the slowest axis, or rows dealt out when rows > 1."""

    def __init__(self, desc, ngpus=0, devices=None):
        self.desc = desc
        self._h = C.c_void_p()
        dev = None if devices is None else (C.c_int * len(devices))(*devices)
        _abi.check(_abi.lib().hrweno_mgpu_create(C.byref(self._h), C.byref(desc), int(ngpus), dev))
        self.ngpus = _abi.lib().hrweno_mgpu_ngpus(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_mgpu_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def slab(self, rank):
        dev, off, cnt = C.c_int(), C.c_int64(), C.c_int64()
        _abi.check(_abi.lib().hrweno_mgpu_slab(self._h, rank, C.byref(dev), C.byref(off), C.byref(cnt)))
        return dev.value, off.value, cnt.value

    # general operators: GLOBAL arrays, cut like the grid inside the library
    def set_xedges(self, axis, xedges):
        xe = np.ascontiguousarray(xedges, dtype=np.float64)
        if xe.size != self.desc.n[axis] + 1:  # weno.f90:101-108
            raise _abi.HrwenoError(_abi.EINVAL, "Invalid input 'xedges'. Valid range: size(xedges) = ncells + 1.")
        _abi.check(_abi.lib().hrweno_mgpu_set_xedges(self._h, axis, xe.ctypes.data))

    def set_flux_coef(self, axis, face=None, cross=None):
        f = None if face is None else np.ascontiguousarray(face, dtype=np.float64)
        c = None if cross is None else np.ascontiguousarray(cross, dtype=np.float64)
        if f is not None and f.size != self.desc.n[axis] + 1:
            raise _abi.HrwenoError(_abi.EINVAL, "face coefficients: size must be ncells + 1 (indexed like edges(0:n))")
        if c is not None and (self.desc.ndim != 2 or c.size != self.desc.n[1 - axis]):
            raise _abi.HrwenoError(_abi.EINVAL, "cross coefficients: ndim == 2 and one value per cell of the other axis")
        _abi.check(_abi.lib().hrweno_mgpu_set_flux_coef(self._h, axis, None if f is None else f.ctypes.data,
                                                        None if c is None else c.ctypes.data))

    def set_flux_time_fn(self, g):
        self._tfn = _abi.TIME_FN(lambda _ctx, t: float(g(t))) if g is not None else C.cast(None, _abi.TIME_FN)
        _abi.check(_abi.lib().hrweno_mgpu_set_flux_time_fn(self._h, self._tfn, None))

    def rktvd(self, order=3):
        _abi.check(_abi.lib().hrweno_mgpu_rktvd(self._h, int(order)))
        return self

    def mstvd(self):
        _abi.check(_abi.lib().hrweno_mgpu_mstvd(self._h))
        return self

    def integrate(self, u, t, tout, dt, itask=1):
        """u: global host vector (advanced in place); returns the new t"""
        if not (isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous):
            raise TypeError("u must be a C-contiguous float64 array (it is updated in place)")
        tt = C.c_double(t)
        _abi.check(_abi.lib().hrweno_mgpu_integrate(self._h, u.ctypes.data, C.byref(tt), float(tout), float(dt), int(itask)))
        return tt.value

    def upload(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        _abi.check(_abi.lib().hrweno_mgpu_upload(self._h, u.ctypes.data))

    def download(self, u):
        _abi.check(_abi.lib().hrweno_mgpu_download(self._h, u.ctypes.data))
        return u

    def integrate_resident(self, t, tout, dt, itask=1):
        tt = C.c_double(t)
        _abi.check(_abi.lib().hrweno_mgpu_integrate_resident(self._h, C.byref(tt), float(tout), float(dt), int(itask)))
        return tt.value

    def max_wavespeed(self, install=True):
        a = C.c_double()
        _abi.check(_abi.lib().hrweno_mgpu_max_wavespeed(self._h, C.byref(a), 1 if install else 0))
        return a.value

    def set_alpha(self, alpha):
        _abi.check(_abi.lib().hrweno_mgpu_set_alpha(self._h, float(alpha)))

    @property
    def fevals(self):
        return _abi.lib().hrweno_mgpu_fevals(self._h)

    @property
    def launches(self):
        return _abi.lib().hrweno_mgpu_launches(self._h)
