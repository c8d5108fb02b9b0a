"""Independent NumPy restatement of the reference path (TEST INFRASTRUCTURE).

Written separately from oracle/hrweno_oracle.c (vectorised over cells instead of a
per-cell routine) so that the two can be compared bit for bit: every NumPy ufunc call
is one correctly rounded IEEE fp64 operation per element, no FMA contraction, which is
the arithmetic model of the reference's gfortran/x86-64 build.  Parity status is the
one stated in hrweno_oracle.h: pinned to the reference's own test tolerances only.

Citations are into /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

def _real(a):
    """kind of the computation: real32 when the data are float32 (the REAL32 build, hrweno_kinds.F90:9-17), else real64"""
    return np.float32 if getattr(a, "dtype", None) == np.float32 else np.float64


def tables(k, T=np.float64):
    """d and c(:, r) of weno.f90:12-21 as literals of kind T: 11.0_rk/6 is the quotient of T, 0.3_rk the nearest T"""
    q = lambda a, b: T(a) / T(b)  # noqa: E731
    d = {1: [T(1)], 2: [q(2, 3), q(1, 3)], 3: [T(0.3), T(0.6), T(0.1)]}[k]
    c = {
        1: [[T(1)], [T(1)]],
        2: [[q(3, 2), q(-1, 2)], [q(1, 2), q(1, 2)], [q(-1, 2), q(3, 2)]],
        3: [[q(11, 6), q(-7, 6), q(1, 3)], [q(1, 3), q(5, 6), q(-1, 6)], [q(-1, 6), q(5, 6), q(1, 3)], [q(1, 3), q(-7, 6), q(11, 6)]],
    }[k]
    return np.array(d, dtype=T), np.array(c, dtype=T)


# src/hrweno_weno.f90:12-21 -- c[k][r+1] is the column c(:, r)
D = {1: np.array([1.0]), 2: np.array([2.0 / 3, 1.0 / 3]), 3: np.array([0.3, 0.6, 0.1])}
C = {
    1: np.array([[1.0], [1.0]]),
    2: np.array([[3.0 / 2, -1.0 / 2], [1.0 / 2, 1.0 / 2], [-1.0 / 2, 3.0 / 2]]),
    3: np.array(
        [
            [11.0 / 6, -7.0 / 6, 1.0 / 3],
            [1.0 / 3, 5.0 / 6, -1.0 / 6],
            [-1.0 / 6, 5.0 / 6, 1.0 / 3],
            [1.0 / 3, -7.0 / 6, 11.0 / 6],
        ]
    ),
}


def calc_cnu(xedges, k):
    """src/hrweno_weno.f90:221-297; returns cnu[i-1, r+1, j]."""
    T = _real(np.asarray(xedges))
    xedges = np.asarray(xedges, dtype=T)
    nc = len(xedges) - 1
    ng = k + 1
    xext = {}
    for i in range(nc + 1):
        xext[i] = T(xedges[i])  # NumPy scalars of kind T: every operation below is rounded to T
    dx = xext[1] - xext[0]
    for i in range(-1, -ng - 1, -1):
        xext[i] = xext[i + 1] - dx
    dx = xext[nc] - xext[nc - 1]
    for i in range(nc + 1, nc + ng + 1):
        xext[i] = xext[i - 1] + dx
    xl = lambda m: xext[m - 1]  # noqa: E731
    xr = lambda m: xext[m]  # noqa: E731
    cnu = np.zeros((nc, k + 1, k), dtype=T)
    for i in range(1, nc + 1):
        for r in range(-1, k):
            for j in range(k):
                sum2 = T(0)
                for m in range(j + 1, k + 1):
                    prod2 = T(1)
                    for l in range(k + 1):
                        if l == m:
                            continue
                        prod2 = prod2 * (xl(i - r + m) - xl(i - r + l))
                    sum1 = T(0)
                    for l in range(k + 1):
                        if l == m:
                            continue
                        prod1 = T(1)
                        for q in range(k + 1):
                            if q == m or q == l:
                                continue
                            prod1 = prod1 * (xr(i) - xl(i - r + q))
                        sum1 = sum1 + prod1
                    sum2 = sum2 + sum1 / prod2
                cnu[i - 1, r + 1, j] = sum2 * (xr(i - r + j) - xl(i - r + j))
    return cnu


def reconstruct(v, k=3, eps=1e-6, cnu=None):
    """src/hrweno_weno.f90:129-219, vectorised over cells; returns (vl, vr).  float32 data: the REAL32 build."""
    T = _real(np.asarray(v))
    v = np.asarray(v, dtype=T)
    eps = T(eps)
    if cnu is not None:
        cnu = np.asarray(cnu, dtype=T)
    nc = v.shape[-1]
    g = k - 1
    pad = [(0, 0)] * (v.ndim - 1) + [(g, g)]
    vext = np.pad(v, pad, mode="edge")  # :171-173

    def s(off):  # vext(i+off) for all cells i
        return vext[..., g + off : g + off + nc]

    d, ctab = (D[k], C[k]) if T is np.float64 else tables(k, T)
    c1312, c14 = T(13) / T(12), T(1) / T(4)
    vrr, vlr = [], []
    for r in range(k):
        sr = np.zeros_like(v)
        sl = np.zeros_like(v)
        for j in range(k):
            if cnu is None:
                cr, cl = ctab[r + 1][j], ctab[r][j]
            else:
                cr, cl = cnu[:, r + 1, j], cnu[:, r, j]
            sr = sr + cr * s(-r + j)  # :179
            sl = sl + cl * s(-r + j)  # :180
        vrr.append(sr)
        vlr.append(sl)
    if k == 1:
        beta = [np.zeros_like(v)]
    elif k == 2:
        beta = [(s(1) - s(0)) ** 2, (s(0) - s(-1)) ** 2]
    else:
        beta = [
            c1312 * ((s(0) - 2 * s(1)) + s(2)) ** 2 + c14 * ((3 * s(0) - 4 * s(1)) + s(2)) ** 2,
            c1312 * ((s(-1) - 2 * s(0)) + s(1)) ** 2 + c14 * (s(-1) - s(1)) ** 2,
            c1312 * ((s(-2) - 2 * s(-1)) + s(0)) ** 2 + c14 * ((s(-2) - 4 * s(-1)) + 3 * s(0)) ** 2,
        ]
    den = [(eps + b) ** 2 for b in beta]
    alfa = [d[r] / den[r] for r in range(k)]
    alfat = [d[k - 1 - r] / den[r] for r in range(k)]
    sa = np.zeros_like(v)
    sat = np.zeros_like(v)
    for r in range(k):
        sa = sa + alfa[r]
        sat = sat + alfat[r]
    vr = np.zeros_like(v)
    vl = np.zeros_like(v)
    for r in range(k):
        vr = vr + (alfa[r] / sa) * vrr[r]
        vl = vl + (alfat[r] / sat) * vlr[r]
    return vl, vr


def flux_model(model, coef, v):
    if model == "burgers":
        return (v * v) / 2  # example1:120
    return coef * v  # example2:140,153


def face_flux(scheme, model, coef, alpha, vm, vp):
    fm = flux_model(model, coef, vm)
    fp = flux_model(model, coef, vp)
    if scheme == "lax_friedrichs":
        return (fm + fp - alpha * (vp - vm)) / 2  # fluxes.f90:43
    return np.where(vm <= vp, np.minimum(fm, fp), np.maximum(fm, fp))  # fluxes.f90:70-74


def grid_linear(xmin, xmax, n):
    """src/hrweno_grids.f90:76-79, 246-247 -> (edges, center, width)."""
    rx = (xmax - xmin) / n
    edges = xmin + rx * np.arange(n + 1, dtype=np.float64)
    return edges, (edges[:-1] + edges[1:]) / 2, edges[1:] - edges[:-1]


def face_flux_x(scheme, model, coef, alpha, vm, vp, cross=None, face=None):
    """x-dependent flux f = (model(v)*cross)*face (growth terms hinted at example2:140,153), absent factors skipped."""

    def f(v):
        out = flux_model(model, coef, v)
        if cross is not None:
            out = out * cross
        if face is not None:
            out = out * face
        return out

    fm, fp = f(vm), f(vp)
    if scheme == "lax_friedrichs":
        return (fm + fp - alpha * (vp - vm)) / 2
    return np.where(vm <= vp, np.minimum(fm, fp), np.maximum(fm, fp))


def _faces(vl, vr, scheme, model, coef, alpha, bc, fcoef=None, cross=None):
    """faces along the last axis: returns f[..., 0:nc+1].  fcoef[0:nc+1] per face, cross[...] per row (broadcast)."""
    if fcoef is None and cross is None:
        inner = face_flux(scheme, model, coef, alpha, vr[..., :-1], vl[..., 1:])
    else:
        inner = face_flux_x(scheme, model, coef, alpha, vr[..., :-1], vl[..., 1:],
                            None if cross is None else np.asarray(cross)[..., None],
                            None if fcoef is None else np.asarray(fcoef)[1:-1])
    if bc == "copy":
        lo, hi = inner[..., :1], inner[..., -1:]
    else:
        lo = np.zeros_like(inner[..., :1])
        hi = lo
    return np.concatenate([lo, inner, hi], axis=-1)


def rhs1d(v, width, k=3, eps=1e-6, scheme="godunov", model="burgers", coef=1.0, alpha=1.0, bc="copy", cnu=None,
          fcoef=None):
    """example/example1_burgers_1d_fv.f90:72-109 (rows of independent problems allowed).
    cnu = calc_cnu(xedges, k) for weno(nc,k,eps,xedges); fcoef[0:nc+1] = x-dependent face coefficient."""
    vl, vr = reconstruct(v, k, eps, cnu)
    f = _faces(vl, vr, scheme, model, coef, alpha, bc, fcoef)
    return -(f[..., 1:] - f[..., :-1]) / width


def rhs2d(v, w1, w2, k=3, eps=1e-6, scheme="godunov", model="linear", coef=(1.0, 1.0), alpha=1.0, bc="zero",
          cnu=(None, None), fcoef=(None, None), ccoef=(None, None)):
    """example/example2_pbe_2d_fv.f90:73-129; v[j, i] with i (x1) contiguous.
    Per axis: cnu (non-uniform tables), fcoef (per face along the axis), ccoef (per cell of the other axis)."""
    vl1, vr1 = reconstruct(v, k, eps, cnu[0])
    f1 = _faces(vl1, vr1, scheme, model, coef[0], alpha, bc, fcoef[0], ccoef[0])
    vt = np.ascontiguousarray(v.T)
    vl2, vr2 = reconstruct(vt, k, eps, cnu[1])
    f2 = _faces(vl2, vr2, scheme, model, coef[1], alpha, bc, fcoef[1], ccoef[1])
    t1 = -(f1[:, 1:] - f1[:, :-1]) / w1[None, :]
    t2 = ((f2[:, 1:] - f2[:, :-1]) / w2[None, :]).T
    return t1 - t2


def is_done(t, tout, dt):
    return (t - tout) * math.copysign(1.0, dt) > 0.0  # tvdode.f90:282


class RK:
    """src/hrweno_tvdode.f90:69-178."""

    def __init__(self, fu, order):
        self.fu, self.order, self.fevals, self.istate = fu, order, 0, 1

    def integrate(self, u, t, tout, dt, itask=1):
        T = _real(np.asarray(u))
        t, tout, dt = T(t), T(tout), T(dt)  # real(rk) :: t, tout, dt (tvdode.f90:107-110)
        if self.istate < 1 or is_done(t, tout, dt):
            return u, t
        fu = self.fu
        while True:
            if self.order == 1:
                u = u + dt * fu(t, u)
            elif self.order == 2:
                ui = u + dt * fu(t, u)
                u = (u + ui + dt * fu(t + dt, ui)) / 2
            else:
                ui = u + dt * fu(t, u)
                ui = (3 * u + ui + dt * fu(t + dt, ui)) / 4
                u = (u + 2 * ui + (2 * dt) * fu(t + dt / 2, ui)) / 3
            t = t + dt
            self.fevals += self.order
            if is_done(t, tout, dt) or itask == 2:
                break
        self.istate = 2
        return u, t


class MS:
    """src/hrweno_tvdode.f90:180-271."""

    def __init__(self, fu):
        self.fu, self.order, self.fevals, self.istate = fu, 3, 0, 1
        self.uold, self.udotold = [None] * 4, [None] * 4

    def integrate(self, u, t, tout, dt):
        T = _real(np.asarray(u))
        t, tout, dt = T(t), T(tout), T(dt)
        if self.istate < 1 or is_done(t, tout, dt):
            return u, t
        if self.istate == 1:
            start = RK(self.fu, 3)
            for c in (3, 2, 1, 0):
                self.uold[c] = u
                self.udotold[c] = self.fu(t, u)
                u, t = start.integrate(u, t, t + 2 * dt, dt, itask=2)
            self.fevals = start.fevals
            self.istate = 2
        while not is_done(t, tout, dt):
            udot = self.fu(t, u)
            ui = (25 * u + (50 * dt) * udot + 7 * self.uold[3] + (10 * dt) * self.udotold[3]) / 32
            t = t + dt
            self.fevals += 1
            self.udotold = [udot] + self.udotold[:3]
            self.uold = [u] + self.uold[:3]
            u = ui
        return u, t


def reconstruct_fast(v, k=3, eps=1e-6):
    """Algebra check of the FAST-mode device formulas (hr-weno_b200/csrc/weno_core.cuh: weno_run_k3_fast,
    weno_run_k2_fast): the same scheme as `reconstruct` (weno.f90:174-216) rewritten in differences of the cell
    averages with division-light weights.  Separately rounded NumPy operations stand in for the device FMAs, so
    this agrees with the kernel to rounding, not bit for bit; tests hold it to a few ULP of `reconstruct`.
    Magnitude guard as on the device (weno_core.cuh: weno_run): cells whose stencil holds a value >= 2^100 are
    evaluated on the stencil scaled by 2^(40 - exponent) with eps scaled by its square (kept >= 2^-240)."""
    v = np.asarray(v, dtype=np.float64)
    nc = v.shape[-1]
    if k == 1:
        return v.copy(), v.copy()
    g = k - 1
    pad = [(0, 0)] * (v.ndim - 1) + [(g, g)]
    vext = np.pad(v, pad, mode="edge")

    def raw(off):
        return vext[..., g + off : g + off + nc]

    # per-cell power-of-two scale (1 for stencils of ordinary magnitude)
    m = np.max(np.stack([np.abs(raw(o)) for o in range(-g, g + 1)]), axis=0)
    big = m >= 2.0**100
    with np.errstate(divide="ignore", invalid="ignore"):
        ex = np.where(big, np.floor(np.log2(np.where(big, m, 1.0))) - 40.0, 0.0)
    sdn, sup = np.exp2(-ex), np.exp2(ex)
    epsc = np.where(big, np.maximum(eps * sdn * sdn, 2.0**-240), eps)

    def s(off):
        return raw(off) * sdn

    def d1(off):  # v[c+off+1] - v[c+off]
        return s(off + 1) - s(off)

    if k == 2:
        den0 = (epsc + d1(0) ** 2) ** 2
        den1 = (epsc + d1(-1) ** 2) ** 2
        h = 0.5 * (d1(0) - d1(-1))
        vr = (s(0) + 0.5 * d1(0)) - (den0 * h) / (2 * den1 + den0)
        vl = (s(0) - 0.5 * d1(-1)) - (den1 * h) / (2 * den0 + den1)
        return vl * sup, vr * sup

    def d2(off):
        return d1(off) - d1(off - 1)

    def d3(off):
        return d2(off + 1) - d2(off)

    eps4 = 4.0 * epsc
    e0 = (d1(1) - 3 * d1(0)) ** 2 + ((13.0 / 3) * d2(1) ** 2 + eps4)
    e1 = (d1(-1) + d1(0)) ** 2 + ((13.0 / 3) * d2(0) ** 2 + eps4)
    e2 = (d1(-2) - 3 * d1(-1)) ** 2 + ((13.0 / 3) * d2(-1) ** 2 + eps4)
    p0, p1, p2 = (e1 * e2) ** 2, (e0 * e2) ** 2, (e0 * e1) ** 2
    # the third differences that straddle the domain edge use replicated ghosts, as vext does
    x, y = p0 * d3(0), p2 * d3(-1)
    a = 3 * (p0 + p2) + 18 * p1
    vrr1 = s(0) + (d1(-1) + 2 * d1(0)) / 6
    vlr1 = s(0) - (2 * d1(-1) + d1(0)) / 6
    vr = vrr1 - (1.5 * x + y) / (6 * p0 + a)
    vl = vlr1 + (x + 1.5 * y) / (6 * p2 + a)
    return vl * sup, vr * sup
