"""type(weno) (src/hrweno_weno.f90:23-50) over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi

# public parameter tables of the reference module (weno.f90:9,15-21), c(j, r) as c[j, r+1]
c1 = np.array([[1.0, 1.0]])
c2 = np.array([[3.0 / 2, 1.0 / 2, -1.0 / 2], [-1.0 / 2, 1.0 / 2, 3.0 / 2]])
c3 = np.array(
    [
        [11.0 / 6, 1.0 / 3, -1.0 / 6, 1.0 / 3],
        [-7.0 / 6, 5.0 / 6, 5.0 / 6, -7.0 / 6],
        [1.0 / 3, -1.0 / 6, 1.0 / 3, 11.0 / 6],
    ]
)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class weno:
    """``weno(ncells, k=3, eps=1e-6, xedges=None)`` -- weno_init (weno.f90:54-127).

    Invalid inputs raise HrwenoError (status EINVAL) where the reference executes
    ``error stop`` after setting ierr = 1 and msg.
    """

    def __init__(self, ncells, k=3, eps=1e-6, xedges=None, mode=None):
        self.msg = ""
        self.ierr = 0
        self._h = C.c_void_p()
        xe = None
        if xedges is not None:
            xe = _f64(xedges)
            if xe.size != ncells + 1:  # weno.f90:101,106
                self.ierr, self.msg = 1, "Invalid input 'xedges': size(xedges) /= ncells + 1."
                raise _abi.HrwenoError(_abi.EINVAL, self.msg)
        st = _abi.lib().hrweno_weno_create(
            C.byref(self._h), int(ncells), int(k), float(eps), xe.ctypes.data if xe is not None else None
        )
        if st != _abi.OK:
            self.ierr = 1
            self.msg = _abi.lib().hrweno_last_error().decode()
            raise _abi.HrwenoError(st, self.msg)
        self.ncells, self.k, self.eps = int(ncells), int(k), float(eps)
        self.uniform_grid = xedges is None
        if mode is not None:  # extension: arithmetic mode of the uniform-table kernel (strict by default)
            _abi.check(_abi.lib().hrweno_weno_set_mode(self._h, int(mode)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_weno_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:  # interpreter shutdown: module globals are already gone
            pass

    @property
    def cnu(self):
        """cnu(0:k-1, -1:k-1, 1:ncells) as an array [i-1, r+1, j] (weno.f90:41)."""
        if self.uniform_grid:
            return None
        out = np.empty((self.ncells, self.k + 1, self.k))
        _abi.check(_abi.lib().hrweno_weno_get_cnu(self._h, out.ctypes.data))
        return out

    def reconstruct(self, v):
        """``call w%reconstruct(v, vl, vr)`` (weno.f90:129-219); returns (vl, vr)."""
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 1 and v.strides[0] != 8 and v.strides[0] % 8 == 0 and v.strides[0] > 0:
            # a strided section such as v(i::nc1) (example2:107)
            if v.shape[0] != self.ncells:
                raise ValueError("size(v) /= ncells")
            inc = v.strides[0] // 8
            vl, vr = np.empty(self.ncells), np.empty(self.ncells)
            base = v.__array_interface__["data"][0]
            _abi.check(
                _abi.lib().hrweno_weno_reconstruct_batch(
                    self._h, 1, base, 0, inc, vl.ctypes.data, vr.ctypes.data, self.ncells
                )
            )
            return vl, vr
        v = _f64(v)
        if v.shape[-1] != self.ncells:
            raise ValueError("size(v) /= ncells")
        vl, vr = np.empty_like(v), np.empty_like(v)
        if v.ndim == 1:
            _abi.check(_abi.lib().hrweno_weno_reconstruct(self._h, v.ctypes.data, vl.ctypes.data, vr.ctypes.data))
        else:
            rows = v.size // self.ncells
            _abi.check(
                _abi.lib().hrweno_weno_reconstruct_batch(
                    self._h, rows, v.ctypes.data, self.ncells, 1, vl.ctypes.data, vr.ctypes.data, self.ncells
                )
            )
        return vl, vr

    def reconstruct_dev(self, v_ptr, vl_ptr, vr_ptr, rows=1, ldv=None, incv=1, ldo=None, stream=None):
        """device-pointer variant, asynchronous on `stream`"""
        ldv = self.ncells if ldv is None else ldv
        ldo = self.ncells if ldo is None else ldo
        _abi.check(
            _abi.lib().hrweno_weno_reconstruct_dev(self._h, rows, v_ptr, ldv, incv, vl_ptr, vr_ptr, ldo, stream)
        )
