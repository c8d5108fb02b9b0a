"""The oracle against every assertion of the reference's own test-drive suites, against the
independent NumPy restatement (bit for bit) and against the structural facts of SURVEY.md
section 8c.  CPU only."""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic, pulse


# ---- test/test_hrweno.f90 ---------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3])
def test_weno_uniform(ref, k):  # test_hrweno.f90:29-67, atol 1e-6 (:37)
    v = pulse(30)
    vl, vr = ref.reconstruct(v, k, 1e-6)
    assert np.max(np.abs(vl - v)) <= 1e-6
    assert np.max(np.abs(vr - v)) <= 1e-6


@pytest.mark.parametrize("k", [1, 2, 3])
def test_calc_cnu_uniform_equals_tables(ref, k):  # test_hrweno.f90:69-112, rtol 1e-5 (:86)
    nc = 30
    xe = np.array([0.0 + (3.0 - 0.0) * i / nc for i in range(nc + 1)])
    cnu = ref.calc_cnu(xe, k)
    _, c = ref.tables(k)
    for i in range(nc):
        np.testing.assert_allclose(cnu[i], c, rtol=1e-5, atol=0)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_weno_nonuniform(ref, k):  # test_hrweno.f90:114-161, atol 1e-6 (:131), default eps
    nc = 30
    xe = np.array([0.0 + (1.0 - 0.0) * i / nc for i in range(nc + 1)]) ** 3
    v = pulse(nc)
    vl, vr = ref.reconstruct(v, k, 1e-6, cnu=ref.calc_cnu(xe, k))
    assert np.max(np.abs(vl - v)) <= 1e-6
    assert np.max(np.abs(vr - v)) <= 1e-6


def test_weno_init_validation(ref):  # weno.f90:72-98
    L = ref.lib()
    assert L.hrweno_ref_weno_check(0, 3, 1e-6) != 0
    assert L.hrweno_ref_weno_check(10, 0, 1e-6) != 0
    assert L.hrweno_ref_weno_check(10, 4, 1e-6) != 0
    assert L.hrweno_ref_weno_check(10, 3, np.finfo(float).eps) != 0
    assert L.hrweno_ref_weno_check(10, 3, 1e-6) == 0


# ---- test/test_tvdode.f90 ---------------------------------------------------------------------
A = np.array([-1.0 + float(ii - 1) * 4 / 9 for ii in range(1, 11)])  # test_tvdode.f90:15


@pytest.mark.parametrize("order,rtol", [(1, 2e-2), (2, 1e-3), (3, 1e-3)])
def test_rktvd(ref, order, rtol):  # test_tvdode.f90:31-70
    t0, tout = -1.3, 1.5
    dt = (tout - t0) / 3000
    u = np.full(10, 0.1)
    ode = ref.rktvd("test_ode", order, neq=10)
    t = ode.integrate(u, t0, tout, dt)
    np.testing.assert_allclose(u, 0.1 * np.exp(A * (t - t0)), rtol=rtol)
    assert ode.fevals == 3001 * order  # 3001 steps
    assert abs(t - 1.50093333333326) < 1e-13
    assert ode.istate == 2


def test_mstvd(ref):  # test_tvdode.f90:72-104
    t0, tout, dt = -1.3, 1.5, 1e-3
    u = np.full(10, 0.1)
    ode = ref.mstvd("test_ode", neq=10)
    t = ode.integrate(u, t0, tout, dt)
    np.testing.assert_allclose(u, 0.1 * np.exp(A * (t - t0)), rtol=1e-3)
    assert ode.fevals == 12 + (2801 - 4)  # 2801 steps, 4 of them RK3 start-up
    assert ode.rhs_calls == ode.fevals + 4


def test_integrate_early_outs(ref):  # tvdode.f90:126-127
    u = np.full(10, 0.1)
    ode = ref.rktvd("test_ode", 3, neq=10)
    assert ode.integrate(u, 1.0, 0.5, 0.1) == 1.0 and ode.fevals == 0  # already past tout
    assert ode.integrate(u, 1.0, 1.0, 0.1) == 1.1 and ode.fevals == 3  # t == tout steps once (strict >)
    assert ode.integrate(u, 1.1, 5.0, 0.1, itask=2) == pytest.approx(1.2) and ode.fevals == 6
    assert ode.integrate(u, 0.0, -0.25, -0.1) == pytest.approx(-0.3)  # negative dt: sign(1,dt)


# ---- test/test_fluxes.f90 ---------------------------------------------------------------------
def test_fluxes(ref):  # test_fluxes.f90:27-71, rtol 1e-8 (:35)
    f = lambda u, x, t: u * x[0] * t  # noqa: E731  (:74-78)
    x, t, vm = [3.0], 5.0, 2.0
    for vm_, vp_ in [(vm, vm), (vm, -2 * vm), (-vm, 2 * vm)]:
        href = f(vm_, x, t)
        assert ref.godunov(f, vm_, vp_, x, t) == pytest.approx(href, rel=1e-8)
        assert ref.lax_friedrichs(f, vm_, vp_, x, t, x[0] * t) == pytest.approx(href, rel=1e-8)
    # the closed-set LINEAR model with a = x*t is the same function
    for vm_, vp_ in [(vm, vm), (vm, -2 * vm), (-vm, 2 * vm)]:
        for scheme in (0, 1):
            assert ref.face_flux(scheme, 1, 15.0, 15.0, [vm_], [vp_])[0] == pytest.approx(15.0 * vm_, rel=1e-8)


# ---- test/test_grid.f90 (linear part, which the path consumes) ---------------------------------
def test_grid_linear(ref, pkg):  # test_grid.f90:31-60
    e, c, w = ref.grid_linear(-5.0, 5.0, 100)
    assert e[0] == -5.0 and abs(e[-1] - 5.0) <= 1e-5 * 5
    assert np.array_equal(w, e[1:] - e[:-1]) and np.array_equal(c, (e[:-1] + e[1:]) / 2)
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    assert np.array_equal(g.edges, e) and np.array_equal(g.width, w) and np.array_equal(g.center, c)
    assert len(set(w.tolist())) == 4  # width is not bitwise uniform


# ---- C oracle == NumPy oracle, bit for bit -----------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3])
def test_c_vs_numpy_reconstruct(ref, npo, k):
    rng = np.random.default_rng(7 + k)
    for v in (rng.standard_normal(257), pulse(30), ex1_ic(np.linspace(-5, 5, 100))):
        a = ref.reconstruct(v, k, 1e-6)
        b = npo.reconstruct(v, k, 1e-6)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    xe = np.sort(rng.uniform(0, 1, 41))
    cnu_c, cnu_n = ref.calc_cnu(xe, k), npo.calc_cnu(xe, k)
    assert np.array_equal(cnu_c, cnu_n)
    v = rng.standard_normal(40)
    a = ref.reconstruct(v, k, 1e-6, cnu=cnu_c)
    b = npo.reconstruct(v, k, 1e-6, cnu=cnu_n)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_c_strided_equals_contiguous(ref):
    rng = np.random.default_rng(3)
    m = rng.standard_normal((50, 7))
    col = np.ascontiguousarray(m[:, 2])
    a = ref.reconstruct(col, 3, 1e-6)
    b = ref.reconstruct(m.reshape(-1)[2:], 3, 1e-6, incv=7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("scheme,model,bc", [(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 1), (0, 0, 1)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_c_vs_numpy_rhs1d(ref, npo, pkg, k, scheme, model, bc):
    rng = np.random.default_rng(11)
    nc, rows = 200, 3
    g = pkg.hrweno_grids.grid1().geometric(0.0, 4.0, 1.01, nc)
    v = ex1_ic(np.linspace(-5, 5, nc))[None, :] + 1e-3 * rng.standard_normal((rows, nc))
    d = pkg.fv.make_desc(nc, k=k, rows=rows, flux_model=model, flux_scheme=scheme, flux_coef=(1.7, 1.0),
                         alpha=1.3, bc=bc, width=[g.width])
    a = ref.FV(d).rhs(0.0, v)
    b = npo.rhs1d(v, g.width, k, 1e-6, ["godunov", "lax_friedrichs"][scheme], ["burgers", "linear"][model],
                  1.7, 1.3, ["copy", "zero"][bc])
    assert np.array_equal(a, b)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_c_vs_numpy_rhs2d(ref, npo, pkg, k):
    rng = np.random.default_rng(5)
    n1, n2 = 37, 23
    g1 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1)
    g2 = pkg.hrweno_grids.grid1().linear(0.0, 7.0, n2)
    v = ex2_ic(g1.center, g2.center) + 1e-3 * rng.standard_normal((n2, n1))
    d = pkg.fv.make_desc((n1, n2), k=k, flux_model=1, bc=1, width=[g1.width, g2.width])
    a = ref.FV(d).rhs(0.0, v)
    b = npo.rhs2d(v, g1.width, g2.width, k)
    assert np.array_equal(a, b)


def test_c_vs_numpy_integrators(ref, npo, pkg):
    nc = 100
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    u0 = ex1_ic(g.center)
    fu = lambda t, u: npo.rhs1d(u, g.width)  # noqa: E731
    for order in (1, 2, 3):
        ode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, width=[g.width])), order)
        u = u0.copy()
        t = ode.integrate(u, 0.0, 0.3, 1e-2)
        un, tn = npo.RK(fu, order).integrate(u0.copy(), 0.0, 0.3, 1e-2)
        assert t == tn and np.array_equal(u, un)
    ode = ref.mstvd(ref.FV(pkg.fv.make_desc(nc, width=[g.width])))
    u = u0.copy()
    t = ode.integrate(u, 0.0, 0.3, 1e-2)
    t = ode.integrate(u, t, 0.5, 1e-2)
    ms = npo.MS(fu)
    un, tn = ms.integrate(u0.copy(), 0.0, 0.3, 1e-2)
    un, tn = ms.integrate(un, tn, 0.5, 1e-2)
    assert t == tn and np.array_equal(u, un) and ode.fevals == ms.fevals


def test_generic_callback_equals_fv_path(ref, npo, pkg):
    nc = 64
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    fv = ref.FV(pkg.fv.make_desc(nc, width=[g.width]))
    u1, u2 = ex1_ic(g.center), ex1_ic(g.center)
    t1 = ref.rktvd(fv, 3).integrate(u1, 0.0, 0.1, 1e-2)
    t2 = ref.rktvd(lambda t, u: fv.rhs(t, u), 3, neq=nc).integrate(u2, 0.0, 0.1, 1e-2)
    assert t1 == t2 and np.array_equal(u1, u2)


# ---- structural identities the CUDA kernels rely on -------------------------
def test_candidate_sharing_identity(npo):
    """vlr(r)_i == vrr(r-1)_{i-1} bit for bit: same coefficients on the same cells."""
    k = 3
    C = npo.C[k]
    rng = np.random.default_rng(0)
    v = rng.standard_normal(64)
    for r in (1, 2):
        for i in range(4, 60):
            vlr = 0.0
            vrr = 0.0
            for j in range(k):
                vlr = vlr + C[r][j] * v[i - r + j]
                vrr = vrr + C[(r - 1) + 1][j] * v[(i - 1) - (r - 1) + j]
            assert vlr == vrr


@pytest.mark.parametrize("k", [1, 2, 3])
def test_slab_with_halo_k_is_bitwise_global(npo, k):
    """a slab updated with k halo cells per side equals the global update; k-1 is not enough"""
    rng = np.random.default_rng(1)
    nc = 96
    w = np.full(nc, 0.1)
    v = ex1_ic(np.linspace(-5, 5, nc)) + 1e-2 * rng.standard_normal(nc)
    full = npo.rhs1d(v, w, k)
    lo, hi = 32, 64
    for h, expect in ((k, True), (k - 1, False)):
        sl = slice(lo - h, hi + h)
        part = npo.rhs1d(v[sl], w[sl], k)[h : h + (hi - lo)]
        assert np.array_equal(part, full[lo:hi]) == expect or (k == 1 and not expect and h == 0)


# ---- the shipped configurations -------------------
def test_config1_example1_facts(ref, pkg):
    nc = 100
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    u = ex1_ic(g.center)
    ode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, width=[g.width])), 3)
    t, dt, times = 0.0, 1e-2, []
    for ii in range(101):
        t = ode.integrate(u, t, 12.0 * ii / 100, dt)  # example1:61-65
        times.append(t)
    assert ode.fevals == 3603 and ode.fevals // 3 == 1201
    assert times[0] == 0.01
    assert repr(times[-1]) == "12.009999999999788"
    assert -0.5 - 1e-4 <= u.min() and u.max() <= 1.0 + 1e-4  # ENO: tiny over/undershoots only


@pytest.mark.slow
def test_config2_example2_facts(ref, pkg):
    n = 250
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    u = ex2_ic(g.center, g.center).reshape(-1)
    ref.set_threads(min(8, ref.max_threads()))  # OpenMP over rows is bitwise neutral
    try:
        ode = ref.mstvd(ref.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g.width, g.width])))
        t, dt, times = 0.0, 5e-3, []
        for ii in range(101):
            t = ode.integrate(u, t, 5.0 * ii / 100, dt)  # example2:61-66
            times.append(t)
    finally:
        ref.set_threads(1)
    assert ode.fevals == 1009 and ode.rhs_calls == 1013
    assert times[0] == 0.02
    assert repr(times[-1]) == "5.0049999999999155"
    mass = float(np.sum(u.reshape(n, n) * g.width[None, :] * g.width[:, None]))
    assert abs(mass - 4.0) < 1e-12
    assert -1e-3 < u.min() < 0 and 1.0 < u.max() < 1.001
