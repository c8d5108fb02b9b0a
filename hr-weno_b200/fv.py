"""The example `rhs` routines (example1:72-109, example2:73-129) as one fused device operator."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def make_desc(
    n,
    k=3,
    eps=1e-6,
    rows=1,
    flux_model=_abi.FLUX_BURGERS,
    flux_scheme=_abi.SCHEME_GODUNOV,
    flux_coef=(1.0, 1.0),
    alpha=1.0,
    bc=_abi.BC_COPY_NEIGHBOUR,
    width=None,
    linear=None,
    mode=_abi.MODE_STRICT,
    rank=0,
    nranks=1,
    global_n=0,
    global_offset=0,
):
    """Build a hrweno_fv_desc.  `width` = list of per-axis width arrays, or `linear` = (xmin, xmax)."""
    n = (int(n),) if np.isscalar(n) else tuple(int(x) for x in n)
    d = _abi.FvDesc()
    d.abi_version = _abi.ABI_VERSION
    d.ndim = len(n)
    d.n[0] = n[0]
    d.n[1] = n[1] if len(n) > 1 else 1
    d.rows = rows
    d.k, d.eps = k, eps
    d.flux_model, d.flux_scheme, d.bc, d.mode = flux_model, flux_scheme, bc, mode
    d.flux_coef[0], d.flux_coef[1] = flux_coef
    d.alpha = alpha
    keep = []
    if linear is not None:
        d.grid_kind = _abi.GRID_LINEAR
        d.xmin, d.xmax = linear
    else:
        d.grid_kind = _abi.GRID_WIDTH_ARRAY
        for a in range(d.ndim):
            w = np.ascontiguousarray(width[a], dtype=np.float64)
            keep.append(w)
            d.width[a] = w.ctypes.data_as(_abi.c_double_p)
    d.rank, d.nranks, d.global_n, d.global_offset = rank, nranks, global_n, global_offset
    d._keepalive = keep
    return d


class FV:
    def __init__(self, desc):
        self.desc = desc
        self._h = C.c_void_p()
        _abi.check(_abi.lib().hrweno_fv_create(C.byref(self._h), C.byref(desc)))
        self.neq = _abi.lib().hrweno_fv_neq(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _abi.lib().hrweno_fv_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:  # interpreter shutdown: module globals are already gone
            pass

    def rhs(self, t, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(v)
        _abi.check(_abi.lib().hrweno_fv_rhs(self._h, t, v.ctypes.data, out.ctypes.data))
        return out

    def rhs_dev(self, t, v_ptr, out_ptr, stream=None):
        _abi.check(_abi.lib().hrweno_fv_rhs_dev(self._h, t, v_ptr, out_ptr, stream))

    def max_wavespeed_dev(self, v_ptr, out_ptr, stream=None):
        """extension: local max |f'(v)| of the dense device vector at v_ptr into the device scalar at out_ptr (async)"""
        _abi.check(_abi.lib().hrweno_fv_max_wavespeed_dev(self._h, v_ptr, out_ptr, stream))

    def set_alpha(self, alpha):
        """extension: Lax-Friedrichs alpha used from the next stage launch on (fluxes.f90:40 takes it per call)"""
        _abi.check(_abi.lib().hrweno_fv_set_alpha(self._h, float(alpha)))

    def set_xedges(self, axis, xedges):
        """weno(ncells, k, eps, xedges) for the sweep along `axis` (weno.f90:100-112): per-cell tables cnu"""
        xe = np.ascontiguousarray(xedges, dtype=np.float64)
        # slabs: along the decomposed (slowest) axis the GLOBAL edges are passed (the tables of the slab's edge cells and of
        # the neighbour cells its interface faces need depend on edges beyond the slab)
        dec = self.desc.nranks > 1 and axis == self.desc.ndim - 1
        want = (self.desc.global_n if dec else self.desc.n[axis]) + 1
        if xe.size != want:  # weno.f90:101-108
            raise _abi.HrwenoError(_abi.EINVAL, "Invalid input 'xedges'. Valid range: size(xedges) = ncells + 1.")
        _abi.check(_abi.lib().hrweno_fv_set_xedges(self._h, axis, xe.ctypes.data))

    def set_flux_coef(self, axis, face=None, cross=None):
        """x-dependent flux f = (model(v)*cross[c])*face[f] for the faces along `axis` (example2:100-101,109-110,140,153)"""
        f = None if face is None else np.ascontiguousarray(face, dtype=np.float64)
        c = None if cross is None else np.ascontiguousarray(cross, dtype=np.float64)
        if f is not None and f.size != self.desc.n[axis] + 1:
            raise _abi.HrwenoError(_abi.EINVAL, "face coefficients: size must be ncells + 1 (indexed like edges(0:n))")
        if c is not None and (self.desc.ndim != 2 or c.size != self.desc.n[1 - axis]):
            raise _abi.HrwenoError(_abi.EINVAL, "cross coefficients: ndim == 2 and one value per cell of the other axis")
        _abi.check(
            _abi.lib().hrweno_fv_set_flux_coef(
                self._h, axis, None if f is None else f.ctypes.data, None if c is None else c.ctypes.data
            )
        )

    def set_flux_time_fn(self, g):
        """t-dependent flux f = ((model(v)*cross)*face)*g(t) (fluxes.f90:12-18): `g(t) -> float` is called on the host once
        per right-hand-side evaluation with that evaluation's time (tvdode.f90:162-166,256); None removes the factor"""
        self._tfn = _abi.TIME_FN(lambda _ctx, t: float(g(t))) if g is not None else C.cast(None, _abi.TIME_FN)
        _abi.check(_abi.lib().hrweno_fv_set_flux_time_fn(self._h, self._tfn, None))

    def export_halo(self):
        buf = C.create_string_buffer(_abi.IPC_HANDLE_BYTES)
        _abi.check(_abi.lib().hrweno_fv_export_halo(self._h, buf))
        return buf.raw

    def import_halo(self, left, right):
        _abi.check(_abi.lib().hrweno_fv_import_halo(self._h, left, right))
