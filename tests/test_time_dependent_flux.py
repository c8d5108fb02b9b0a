"""t-dependent fluxes on the fused path (fluxes.f90:12-18 hands the evaluation time to the flux, the
integrators evaluate the rhs at t, t+dt, t+dt/2 -- tvdode.f90:162-166 -- and at t, :256).  Closed-set form: a separable
factor f(v, x, t) = ((model(v)*cross)*face)*g(t), g evaluated on the host once per rhs evaluation
(hrweno_fv_set_flux_time_fn).  CPU: the oracle's g(t) path against the same run spelled with per-stage face
coefficients through the generic callback integrator.  GPU: fused stage kernels against the oracle, bit for bit."""
import math

import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic


def g_of_t(t):
    return 1.0 + 0.5 * math.sin(3.0 * t) + 0.25 * t


def _steps_to(t, dt, k):
    for _ in range(k - 1):
        t = t + dt
    return t


@pytest.mark.parametrize("kind", ["rk3", "rk2", "ms"])
def test_oracle_time_factor_equals_per_stage_face_coefficient(pkg, ref, kind):
    """(model(v)*face[f]) with face[:] = g(t_stage) is the same product as model(v)*g(t_stage): the generic callback
    integrator (tvdode.f90:50-57: fu(t, u, udot)) rebuilding the coefficient at every call must reproduce the g(t) path"""
    nc = 257
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(5).standard_normal(nc)
    kw = dict(k=3, width=[g.width])
    fvt = ref.FV(pkg.fv.make_desc(nc, **kw))
    fvt.set_flux_time_fn(g_of_t)
    fvc = ref.FV(pkg.fv.make_desc(nc, **kw))
    times = []

    def rhs(t, u):
        times.append(t)
        fvc.set_flux_coef(0, np.full(nc + 1, g_of_t(t)), None)
        return fvc.rhs(t, u)

    if kind == "ms":
        a, b = ref.mstvd(fvt), ref.mstvd(rhs, neq=nc)
    else:
        a, b = ref.rktvd(fvt, int(kind[2])), ref.rktvd(rhs, int(kind[2]), neq=nc)
    ua, ub, dt = u0.copy(), u0.copy(), 2e-3
    ta = a.integrate(ua, 0.3, _steps_to(0.3, dt, 9), dt)
    tb = b.integrate(ub, 0.3, _steps_to(0.3, dt, 9), dt)
    assert ta == tb and np.array_equal(ua, ub)
    if kind == "rk3":
        assert times[:3] == [0.3, 0.3 + dt, 0.3 + dt / 2]  # tvdode.f90:162,164,166
    # and the factor matters
    c = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, **kw)), 3)
    uc = u0.copy()
    c.integrate(uc, 0.3, _steps_to(0.3, dt, 9), dt)
    assert not np.array_equal(ua, uc)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("kind", ["rk1", "rk2", "rk3", "ms"])
def test_gpu_time_factor_1d_bitwise(gpu_lib, pkg, ref, kind, mode):
    nc = 5003
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.0003, nc)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(11).standard_normal(nc)
    m = pkg._abi.MODE_FAST if mode == "fast" else pkg._abi.MODE_STRICT
    kw = dict(k=3, width=[g.width])
    fv, rfv = pkg.fv.FV(pkg.fv.make_desc(nc, mode=m, **kw)), ref.FV(pkg.fv.make_desc(nc, **kw))
    for f in (fv, rfv):
        f.set_xedges(0, g.edges)
        f.set_flux_time_fn(g_of_t)
    assert np.array_equal(fv.rhs(0.7, u0), rfv.rhs(0.7, u0))
    if kind == "ms":
        ode, rode = pkg.hrweno_tvdode.mstvd(fv, nc), ref.mstvd(rfv)
    else:
        ode, rode = pkg.hrweno_tvdode.rktvd(fv, nc, int(kind[2])), ref.rktvd(rfv, int(kind[2]))
    u, ur, t, tr, dt = u0.copy(), u0.copy(), 0.1, 0.1, 2e-4
    for nsteps in (1, 7, 12):
        t = ode.integrate(u, t, _steps_to(t, dt, nsteps), dt)
        tr = rode.integrate(ur, tr, _steps_to(tr, dt, nsteps), dt)
        assert t == tr and np.array_equal(u, ur)
    fv.set_flux_time_fn(None)
    rfv.set_flux_time_fn(None)
    assert np.array_equal(fv.rhs(0.7, u0), rfv.rhs(0.7, u0))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rk3", "ms"])
def test_gpu_time_factor_2d_growth_bitwise(gpu_lib, pkg, ref, kind):
    """example2 with the growth terms of :140,153 un-commented and a time factor on top: v*x(1)**2*g(t), v*x(1)*x(2)*g(t)"""
    n1, n2 = 130, 97
    g1, g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.01, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    u0 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(2).standard_normal((n2, n1))).reshape(-1)
    kw = dict(flux_model=1, bc=1, width=[g1.width, g2.width])
    fv, rfv = pkg.fv.FV(pkg.fv.make_desc((n1, n2), **kw)), ref.FV(pkg.fv.make_desc((n1, n2), **kw))
    for f in (fv, rfv):
        f.set_xedges(0, g1.edges)
        f.set_flux_coef(0, g1.edges**2, None)
        f.set_flux_coef(1, g2.edges, g1.center)
        f.set_flux_time_fn(g_of_t)
    if kind == "ms":
        ode, rode = pkg.hrweno_tvdode.mstvd(fv, n1 * n2), ref.mstvd(rfv)
    else:
        ode, rode = pkg.hrweno_tvdode.rktvd(fv, n1 * n2, 3), ref.rktvd(rfv, 3)
    u, ur, dt = u0.copy(), u0.copy(), 1e-4
    t = ode.integrate(u, 0.0, _steps_to(0.0, dt, 11), dt)
    tr = rode.integrate(ur, 0.0, _steps_to(0.0, dt, 11), dt)
    assert t == tr and np.array_equal(u, ur)
