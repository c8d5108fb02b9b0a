"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same inputs.

Bar (north star): strict mode is bit-identical to the oracle (0 ULP; -0 == +0 allowed);
fast mode is within 1e-12 normwise at every output time and a few ULP for reconstruct-only.
"""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic, normwise, pulse

pytestmark = pytest.mark.gpu

SIZES_1D = [2, 3, 5, 30, 100, 1015, 1016, 1017, 2033, 5000]


# ---- reconstruct (weno.f90:129-219) --------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3])
def test_reconstruct_reference_test_uniform(gpu_lib, pkg, ref, k):
    """test/test_hrweno.f90:29-67 through the reference-shaped API, plus bitwise vs the oracle"""
    v = pulse(30)
    w = pkg.hrweno_weno.weno(30, k, eps=1e-6)
    vl, vr = w.reconstruct(v)
    assert np.max(np.abs(vl - v)) <= 1e-6 and np.max(np.abs(vr - v)) <= 1e-6
    rl, rr = ref.reconstruct(v, k, 1e-6)
    assert np.array_equal(vl, rl) and np.array_equal(vr, rr)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nc", [1, 2, 3, 7, 257, 4096, 100003])
def test_reconstruct_bitwise(gpu_lib, pkg, ref, k, nc):
    rng = np.random.default_rng(100 * k + nc % 97)
    v = rng.standard_normal(nc)
    w = pkg.hrweno_weno.weno(nc, k, 1e-6)
    vl, vr = w.reconstruct(v)
    rl, rr = ref.reconstruct(v, k, 1e-6)
    assert np.array_equal(vl, rl) and np.array_equal(vr, rr)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_reconstruct_nonuniform(gpu_lib, pkg, ref, k):
    """test/test_hrweno.f90:69-161: cnu on a uniform grid equals the tables; cubic grid pulse"""
    nc = 30
    xe = np.array([0.0 + 3.0 * i / nc for i in range(nc + 1)])
    w = pkg.hrweno_weno.weno(nc, k, 1e-6, xedges=xe)
    tab = {1: pkg.hrweno_weno.c1, 2: pkg.hrweno_weno.c2, 3: pkg.hrweno_weno.c3}[k]
    for i in range(nc):
        np.testing.assert_allclose(w.cnu[i], tab.T, rtol=1e-5)
    assert np.array_equal(w.cnu, ref.calc_cnu(xe, k))
    xe = np.array([i / nc for i in range(nc + 1)]) ** 3
    w = pkg.hrweno_weno.weno(nc, k, xedges=xe)
    v = pulse(nc)
    vl, vr = w.reconstruct(v)
    assert np.max(np.abs(vl - v)) <= 1e-6 and np.max(np.abs(vr - v)) <= 1e-6
    rl, rr = ref.reconstruct(v, k, 1e-6, cnu=ref.calc_cnu(xe, k))
    assert np.array_equal(vl, rl) and np.array_equal(vr, rr)
    rng = np.random.default_rng(k)
    v = rng.standard_normal(nc)
    vl, vr = w.reconstruct(v)
    rl, rr = ref.reconstruct(v, k, 1e-6, cnu=ref.calc_cnu(xe, k))
    assert np.array_equal(vl, rl) and np.array_equal(vr, rr)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_reconstruct_fast_mode_within_a_few_ulp(gpu_lib, pkg, ref, k):
    """north star: reconstruction-only results at identical inputs within a few ULP (fast mode, difference form)"""
    rng = np.random.default_rng(40 + k)
    for nc in (1, 2, 5, 6, 1001, 65536):
        x = np.linspace(-5, 5, nc)
        v = np.clip(1.0 - 0.25 * (x + 4.0), -0.5, 1.0) + 1e-3 * rng.standard_normal(nc)
        vl, vr = pkg.hrweno_weno.weno(nc, k, 1e-6, mode=pkg._abi.MODE_FAST).reconstruct(v)
        rl, rr = ref.reconstruct(v, k, 1e-6)
        tol = 4 * np.finfo(float).eps * np.max(np.abs(v))
        assert np.max(np.abs(vl - rl)) <= tol and np.max(np.abs(vr - rr)) <= tol


@pytest.mark.parametrize("k", [2, 3])
def test_reconstruct_dev_unaligned_pointers_and_odd_pitches(gpu_lib, pkg, ref, k):
    """device-pointer entry: the 16-B vector path needs aligned pointers and even pitches; anything else must take the
    scalar path and give the same bits"""
    import torch

    rng = np.random.default_rng(k)
    rows, nc = 5, 1003
    w = pkg.hrweno_weno.weno(nc, k, 1e-6)
    for ldv, ldo, shift in [(nc, nc, 0), (nc + 1, nc + 1, 0), (nc + 5, nc + 3, 1), (nc + 1, nc + 1, 1)]:
        m = rng.standard_normal((rows, ldv))
        buf = torch.zeros(rows * ldv + 2, dtype=torch.float64, device="cuda")
        buf[shift:shift + rows * ldv] = torch.from_numpy(m.reshape(-1)).cuda()
        ol = torch.zeros(rows * ldo + 2, dtype=torch.float64, device="cuda")
        orr = torch.zeros(rows * ldo + 2, dtype=torch.float64, device="cuda")
        w.reconstruct_dev(buf.data_ptr() + 8 * shift, ol.data_ptr() + 8 * shift, orr.data_ptr() + 8 * shift, rows=rows, ldv=ldv, ldo=ldo)
        torch.cuda.synchronize()
        gl = ol[shift:shift + rows * ldo].cpu().numpy().reshape(rows, ldo)[:, :nc]
        gr = orr[shift:shift + rows * ldo].cpu().numpy().reshape(rows, ldo)[:, :nc]
        for j in range(rows):
            rl, rr = ref.reconstruct(m[j, :nc], k, 1e-6)
            assert np.array_equal(gl[j], rl) and np.array_equal(gr[j], rr), (ldv, ldo, shift, j)


def test_reconstruct_strided_and_batched(gpu_lib, pkg, ref):
    """the sections of example2:98,107: contiguous rows and stride-nc1 columns"""
    rng = np.random.default_rng(5)
    n1, n2 = 37, 23
    m = rng.standard_normal((n2, n1))
    w1, w2 = pkg.hrweno_weno.weno(n1, 3, 1e-6), pkg.hrweno_weno.weno(n2, 3, 1e-6)
    vl, vr = w1.reconstruct(m)  # all rows in one call
    for j in range(n2):
        rl, rr = ref.reconstruct(m[j], 3, 1e-6)
        assert np.array_equal(vl[j], rl) and np.array_equal(vr[j], rr)
    for i in (0, 5, n1 - 1):
        vl, vr = w2.reconstruct(m[:, i])  # strided view
        rl, rr = ref.reconstruct(np.ascontiguousarray(m[:, i]), 3, 1e-6)
        assert np.array_equal(vl, rl) and np.array_equal(vr, rr)
    for bad in (m[:-1, 0], m[::2, 0], m[0, :-1]):  # size(v) /= ncells: refused before any pointer reaches the C side
        with pytest.raises(ValueError):
            w2.reconstruct(bad)


# ---- fluxes (fluxes.f90) -----------------------------------------------------------------------------
def test_fluxes_reference_test(gpu_lib, pkg):
    """test/test_fluxes.f90:27-71 (rtol 1e-8) for the pointwise helpers and the device face kernel"""
    f = lambda u, x, t: u * x[0] * t  # noqa: E731
    x, t, vm = [3.0], 5.0, 2.0
    for vm_, vp_ in [(vm, vm), (vm, -2 * vm), (-vm, 2 * vm)]:
        href = f(vm_, x, t)
        assert pkg.hrweno_fluxes.godunov(f, vm_, vp_, x, t) == pytest.approx(href, rel=1e-8)
        assert pkg.hrweno_fluxes.lax_friedrichs(f, vm_, vp_, x, t, x[0] * t) == pytest.approx(href, rel=1e-8)
        for scheme in (0, 1):
            h = pkg.hrweno_fluxes.flux_faces(scheme, 1, [vm_], [vp_], coef=15.0, alpha=15.0)
            assert h[0] == pytest.approx(href, rel=1e-8)


@pytest.mark.parametrize("scheme", [0, 1])
@pytest.mark.parametrize("model", [0, 1])
def test_flux_faces_bitwise(gpu_lib, pkg, ref, scheme, model):
    rng = np.random.default_rng(9)
    vm, vp = rng.standard_normal(1000), rng.standard_normal(1000)
    vp[:100] = vm[:100]
    h = pkg.hrweno_fluxes.flux_faces(scheme, model, vm, vp, coef=1.7, alpha=1.3)
    assert np.array_equal(h, ref.face_flux(scheme, model, 1.7, 1.3, vm, vp))


# ---- fused rhs, 1D (example1:72-109) --------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nc", SIZES_1D)
def test_rhs1d_bitwise_sizes(gpu_lib, pkg, ref, k, nc):
    rng = np.random.default_rng(nc)
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    v = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)
    d = pkg.fv.make_desc(nc, k=k, width=[g.width])
    assert np.array_equal(pkg.fv.FV(d).rhs(0.0, v), ref.FV(d).rhs(0.0, v))


@pytest.mark.parametrize("scheme,model,bc", [(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 1), (0, 0, 1), (1, 1, 0)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_rhs1d_bitwise_variants_batched(gpu_lib, pkg, ref, k, scheme, model, bc):
    rng = np.random.default_rng(11)
    nc, rows = 1500, 7
    g = pkg.hrweno_grids.grid1().geometric(0.0, 4.0, 1.001, nc)
    v = ex1_ic(np.linspace(-5, 5, nc))[None, :] + 1e-3 * rng.standard_normal((rows, nc))
    d = pkg.fv.make_desc(nc, k=k, rows=rows, flux_model=model, flux_scheme=scheme, flux_coef=(1.7, 1.0),
                         alpha=1.3, bc=bc, width=[g.width])
    assert np.array_equal(pkg.fv.FV(d).rhs(0.0, v), ref.FV(d).rhs(0.0, v))


def test_rhs1d_linear_grid_analytic_width_is_bitwise(gpu_lib, pkg, ref):
    """GRID_LINEAR recomputes grid1%linear's width on the device (grids.f90:76-79,247)"""
    rng = np.random.default_rng(2)
    for nc, (a, b) in [(100, (-5.0, 5.0)), (4097, (0.0, 10.0)), (1000, (-1.3, 7.7))]:
        g = pkg.hrweno_grids.grid1().linear(a, b, nc)
        v = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)
        got = pkg.fv.FV(pkg.fv.make_desc(nc, linear=(a, b))).rhs(0.0, v)
        want = ref.FV(pkg.fv.make_desc(nc, width=[g.width])).rhs(0.0, v)
        assert np.array_equal(got, want)


# ---- integrators (tvdode.f90) ---------------------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_rktvd_fused_bitwise(gpu_lib, pkg, ref, k, order):
    nc = 300
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    kw = dict(n=nc, k=k, width=[g.width])
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), nc, order)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), order)
    u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
    for tout in (0.0, 0.05, 0.05, 0.2):  # includes t == tout (steps once) and an already-done call
        t = ode.integrate(u, t, tout, 5e-3)
        tr = rode.integrate(ur, tr, tout, 5e-3)
        assert t == tr and np.array_equal(u, ur)
    t, tr = ode.integrate(u, t, 9.0, 5e-3, itask=2), rode.integrate(ur, tr, 9.0, 5e-3, itask=2)
    assert t == tr and np.array_equal(u, ur)
    assert ode.fevals == rode.fevals and ode.istate == rode.istate == 2


def test_config1_example1_every_output_time(gpu_lib, pkg, ref):
    """configs[0]: example1 as shipped (Godunov, example1:99) -- and with Lax-Friedrichs alpha=1 as
    BASELINE.json words it (example1:98).  Bitwise at all 101 output times."""
    nc = 100
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    for scheme in (0, 1):
        kw = dict(n=nc, k=3, eps=1e-6, flux_scheme=scheme, alpha=1.0, width=[g.width])
        ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), nc, 3)
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), 3)
        u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
        for ii in range(101):
            tout = 12.0 * ii / 100
            t = ode.integrate(u, t, tout, 1e-2)
            tr = rode.integrate(ur, tr, tout, 1e-2)
            assert t == tr and np.array_equal(u, ur), f"output {ii}: normwise {normwise(u, ur):.3e}"
        assert ode.fevals == 3603 and repr(t) == "12.009999999999788"


def test_mstvd_fused_1d_bitwise(gpu_lib, pkg, ref):
    nc = 500
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    kw = dict(n=nc, k=3, width=[g.width])
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), nc)
    rode = ref.mstvd(ref.FV(pkg.fv.make_desc(**kw)))
    u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
    for tout in (0.0, 0.1, 0.1, 0.25, 0.6):
        t = ode.integrate(u, t, tout, 4e-3)
        tr = rode.integrate(ur, tr, tout, 4e-3)
        assert t == tr and np.array_equal(u, ur)
    assert ode.fevals == rode.fevals


@pytest.mark.parametrize("k", [1, 2, 3])
def test_fast_mode_within_north_star_tolerance(gpu_lib, pkg, ref, k):
    """fast mode: 1e-12 normwise at every output time (tolerance from BASELINE.json north_star)"""
    nc = 100
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    kw = dict(n=nc, k=k, eps=1e-6, width=[g.width])
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)), nc, 3)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), 3)
    u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
    worst = 0.0
    for ii in range(101):
        tout = 12.0 * ii / 100
        t = ode.integrate(u, t, tout, 1e-2)
        tr = rode.integrate(ur, tr, tout, 1e-2)
        assert t == tr
        worst = max(worst, normwise(u, ur))
    assert worst <= 1e-12, worst


def test_generic_callback_integrators_device_resident(gpu_lib, pkg, ref):
    """rktvd(fu, neq, order) / mstvd(fu, neq) with a user rhs working on device memory (tvdode.f90:50-57)"""
    nc = 200
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    kw = dict(n=nc, k=3, width=[g.width])
    fv = pkg.fv.FV(pkg.fv.make_desc(**kw))
    calls = []

    def fu(t, neq, u_ptr, udot_ptr, stream):
        calls.append(t)
        fv.rhs_dev(t, u_ptr, udot_ptr, stream)

    for make, rmake in [
        (lambda: pkg.hrweno_tvdode.rktvd(fu, nc, 3), lambda: ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), 3)),
        (lambda: pkg.hrweno_tvdode.mstvd(fu, nc), lambda: ref.mstvd(ref.FV(pkg.fv.make_desc(**kw)))),
    ]:
        ode, rode = make(), rmake()
        u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
        for tout in (0.0, 0.1, 0.3):
            t = ode.integrate(u, t, tout, 5e-3)
            tr = rode.integrate(ur, tr, tout, 5e-3)
            assert t == tr and np.array_equal(u, ur)
        assert ode.fevals == rode.fevals
    assert calls[:3] == [0.0, 0.005, 0.0025]  # stage times t, t+dt, t+dt/2 (tvdode.f90:162-166)


def test_reference_tvdode_tests_through_gpu_path(gpu_lib, pkg, ref):
    """test/test_tvdode.f90 with the linear test ODE evaluated by a device rhs (a*u as LINEAR flux is not
    applicable, so the callback scales on the device through the combine path): use the oracle numbers."""
    a = np.array([-1.0 + float(ii - 1) * 4 / 9 for ii in range(1, 11)])
    torch = pytest.importorskip("torch")
    a_dev = torch.tensor(a, dtype=torch.float64, device="cuda")

    def fu(t, neq, u_ptr, udot_ptr, stream):
        # udot = a*u enqueued on the integrator's stream (torch only as a device-array convenience here)
        n = int(neq)
        with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
            _as_tensor(torch, udot_ptr, n).copy_(a_dev * _as_tensor(torch, u_ptr, n))

    t0, tout = -1.3, 1.5
    for order, rtol, dt in [(1, 2e-2, (tout - t0) / 3000), (2, 1e-3, (tout - t0) / 3000), (3, 1e-3, (tout - t0) / 3000)]:
        ode = pkg.hrweno_tvdode.rktvd(fu, 10, order)
        u = np.full(10, 0.1)
        t = ode.integrate(u, t0, tout, dt)
        np.testing.assert_allclose(u, 0.1 * np.exp(a * (t - t0)), rtol=rtol)
        ur = np.full(10, 0.1)
        tr = ref.rktvd("test_ode", order, neq=10).integrate(ur, t0, tout, dt)
        assert t == tr and np.array_equal(u, ur)


def _as_tensor(torch, ptr, n):
    """wrap a raw device pointer as a torch tensor (test helper)"""

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_Holder(), device="cuda")


# ---- fused rhs, 2D dimension-split (example2:73-129) -----------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("n1,n2", [(5, 7), (2, 2), (37, 23), (64, 32), (65, 33), (130, 70), (250, 250)])
def test_rhs2d_bitwise_sizes(gpu_lib, pkg, ref, k, n1, n2):
    rng = np.random.default_rng(n1 * 1000 + n2)
    g1 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1)
    g2 = pkg.hrweno_grids.grid1().linear(0.0, 7.0, n2)
    v = ex2_ic(g1.center, g2.center) + 1e-3 * rng.standard_normal((n2, n1))
    d = pkg.fv.make_desc((n1, n2), k=k, flux_model=1, bc=1, width=[g1.width, g2.width])
    assert np.array_equal(pkg.fv.FV(d).rhs(0.0, v), ref.FV(d).rhs(0.0, v))


@pytest.mark.parametrize("scheme,model,bc", [(0, 0, 0), (1, 0, 0), (1, 1, 1), (0, 0, 1), (1, 1, 0)])
def test_rhs2d_bitwise_variants(gpu_lib, pkg, ref, scheme, model, bc):
    rng = np.random.default_rng(17)
    n1, n2 = 150, 90
    g1 = pkg.hrweno_grids.grid1().geometric(0.0, 4.0, 1.01, n1)
    g2 = pkg.hrweno_grids.grid1().log(0.5, 9.0, n2)
    v = rng.standard_normal((n2, n1))
    d = pkg.fv.make_desc((n1, n2), k=3, flux_model=model, flux_scheme=scheme, flux_coef=(1.7, -0.6), alpha=1.3,
                         bc=bc, width=[g1.width, g2.width])
    assert np.array_equal(pkg.fv.FV(d).rhs(0.0, v), ref.FV(d).rhs(0.0, v))


@pytest.mark.parametrize("order", [1, 2, 3])
def test_rktvd_fused_2d_bitwise(gpu_lib, pkg, ref, order):
    n1, n2 = 70, 45
    g1 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1)
    g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    kw = dict(n=(n1, n2), k=3, flux_model=1, bc=1, width=[g1.width, g2.width])
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), n1 * n2, order)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), order)
    u = ex2_ic(g1.center, g2.center).reshape(-1)
    ur = u.copy()
    t, tr = 0.0, 0.0
    for tout in (0.0, 0.1, 0.3):
        t = ode.integrate(u, t, tout, 2e-2)
        tr = rode.integrate(ur, tr, tout, 2e-2)
        assert t == tr and np.array_equal(u, ur)


def test_config2_example2_every_output_time(gpu_lib, pkg, ref):
    """configs[1]: example2 as shipped (250x250, k=3, mstvd, dt=5e-3, 101 outputs to t=5): bitwise at every output"""
    n = 250
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    kw = dict(n=(n, n), k=3, eps=1e-6, flux_model=1, bc=1, width=[g.width, g.width])
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), n * n)
    ref.set_threads(min(16, ref.max_threads()))
    try:
        rode = ref.mstvd(ref.FV(pkg.fv.make_desc(**kw)))
        u = ex2_ic(g.center, g.center).reshape(-1)
        ur = u.copy()
        t, tr = 0.0, 0.0
        for ii in range(101):
            tout = 5.0 * ii / 100
            t = ode.integrate(u, t, tout, 5e-3)
            tr = rode.integrate(ur, tr, tout, 5e-3)
            assert t == tr and np.array_equal(u, ur), f"output {ii}: normwise {normwise(u, ur):.3e}"
    finally:
        ref.set_threads(1)
    assert ode.fevals == 1009 and repr(t) == "5.0049999999999155"
    assert abs(float(np.sum(u.reshape(n, n) * g.width[None, :] * g.width[:, None])) - 4.0) < 1e-12


def test_config2_fast_mode_within_tolerance(gpu_lib, pkg, ref):
    """fast mode on example2: 1e-12 normwise at every output time (SURVEY 7.4-1 measured 3.5e-13 drift)"""
    n = 250
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    kw = dict(n=(n, n), k=3, eps=1e-6, flux_model=1, bc=1, width=[g.width, g.width])
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)), n * n)
    ref.set_threads(min(16, ref.max_threads()))
    try:
        rode = ref.mstvd(ref.FV(pkg.fv.make_desc(**kw)))
        u = ex2_ic(g.center, g.center).reshape(-1)
        ur = u.copy()
        t, tr, worst = 0.0, 0.0, 0.0
        for ii in range(101):
            tout = 5.0 * ii / 100
            t = ode.integrate(u, t, tout, 5e-3)
            tr = rode.integrate(ur, tr, tout, 5e-3)
            assert t == tr
            worst = max(worst, normwise(u, ur))
    finally:
        ref.set_threads(1)
    assert worst <= 1e-12, worst


def test_fast_mode_rhs_within_a_few_ulp(gpu_lib, pkg, ref):
    """one rhs evaluation in fast mode against the oracle: a few ULP of the quantities it differences (north star:
    'a few ULP' per reconstruction; the divergence divides O(ulp) flux differences by the cell width)"""
    rng = np.random.default_rng(4)
    nc = 4000
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    v = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)
    for k in (1, 2, 3):
        got = pkg.fv.FV(pkg.fv.make_desc(nc, k=k, width=[g.width], mode=pkg._abi.MODE_FAST)).rhs(0.0, v)
        want = ref.FV(pkg.fv.make_desc(nc, k=k, width=[g.width])).rhs(0.0, v)
        scale = 0.5 * np.max(v * v) / g.width.min()  # magnitude of f/width: the rhs is a difference of two such terms
        assert np.max(np.abs(got - want)) <= 16 * np.finfo(float).eps * scale
    n1, n2 = 130, 70
    g1, g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1), pkg.hrweno_grids.grid1().linear(0.0, 7.0, n2)
    v2 = ex2_ic(g1.center, g2.center) + 1e-3 * rng.standard_normal((n2, n1))
    kw = dict(n=(n1, n2), k=3, flux_model=1, bc=1, width=[g1.width, g2.width])
    got = pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)).rhs(0.0, v2)
    want = ref.FV(pkg.fv.make_desc(**kw)).rhs(0.0, v2)
    assert np.max(np.abs(got - want)) <= 16 * np.finfo(float).eps * np.max(np.abs(v2)) / min(g1.width.min(), g2.width.min())


@pytest.mark.parametrize("order", [1, 2, 3])
def test_host_integrate_chunk_pipeline_bitwise(gpu_lib, pkg, ref, order, monkeypatch):
    """the host-pointer integrate of a large 1D row runs as a time-skewed chunk pipeline (ode.cu: every stage shifts the
    chunk boundaries one tile to the left); forced here at a small size with four tiles per chunk, so that the skew of the
    longer calls moves whole chunks out of the domain: results must equal the oracle's bit for bit, call after call"""
    monkeypatch.setenv("HRWENO_PIPE_CHUNK_TILES", "4")
    nc = 41 * 1016 + 333  # 11 chunks, the last one partial
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    rng = np.random.default_rng(order)
    u0 = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)
    for kw in (dict(n=nc, k=3, linear=(-5.0, 5.0)), dict(n=nc, k=2, flux_scheme=1, alpha=1.2, width=[g.width])):
        rkw = dict(kw)
        rkw.pop("linear", None)
        rkw["width"] = [g.width]
        ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), nc, order)
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**rkw)), order)
        launches0 = ode.launches
        u, ur, t, tr = u0.copy(), u0.copy(), 0.0, 0.0
        dt = 0.2 * 10.0 / nc
        for tout in (0.0, 7 * dt, 7 * dt, 20 * dt):
            t = ode.integrate(u, t, tout, dt)
            tr = rode.integrate(ur, tr, tout, dt)
            assert t == tr and np.array_equal(u, ur)
        t, tr = ode.integrate(u, t, 1e9, dt, itask=2), rode.integrate(ur, tr, 1e9, dt, itask=2)
        assert t == tr and np.array_equal(u, ur) and ode.fevals == rode.fevals
        assert ode.launches - launches0 > 2 * order * 22  # more than one launch per stage: the pipeline was taken


@pytest.mark.parametrize("k", [2, 3])
def test_host_integrate_chunk_pipeline_fast_mode_equals_resident(gpu_lib, pkg, k, monkeypatch):
    """fast mode: the chunk pipeline must reproduce the plain H2D / stages / D2H sequence bit for bit (same kernels, same
    per-cell arithmetic, only the launch ranges differ)"""
    nc = 37 * 1008 + 123
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(k).standard_normal(nc)
    dt = 0.2 * 10.0 / nc
    out = []
    for chunk in ("0", "3"):  # 0 = pipeline off
        monkeypatch.setenv("HRWENO_PIPE_CHUNK_TILES", chunk)
        ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(n=nc, k=k, linear=(-5.0, 5.0), mode=pkg._abi.MODE_FAST)), nc, 3)
        u, t = u0.copy(), 0.0
        for tout in (5 * dt, 30 * dt):
            t = ode.integrate(u, t, tout, dt)
        out.append((t, u, ode.launches))
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])
    assert out[1][2] > 2 * out[0][2]


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [2, 3])
def test_rktvd_fused_tiled_path_bitwise(gpu_lib, pkg, ref, k, order):
    """more than 1024 cells: the tiled persistent kernel (the <=1024-cell case takes the single-launch kernel)"""
    nc = 2500
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    rng = np.random.default_rng(k * 10 + order)
    u0 = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)
    kw = dict(n=nc, k=k, width=[g.width])
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(**kw)), nc, order)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), order)
    u, ur, t, tr = u0.copy(), u0.copy(), 0.0, 0.0
    for tout in (0.0, 0.01, 0.03):
        t = ode.integrate(u, t, tout, 5e-4)
        tr = rode.integrate(ur, tr, tout, 5e-4)
        assert t == tr and np.array_equal(u, ur)
    assert ode.fevals == rode.fevals


# ---- adaptive Lax-Friedrichs alpha (extension, K5) ---------------------------------------------------------
@pytest.mark.parametrize("nc", [2, 3, 7, 4096, 100003])
def test_max_wavespeed_matches_numpy(gpu_lib, pkg, nc):
    """local max |f'(v)|: Burgers -> max|v| (odd/even sizes, unaligned pointer, NaN cells ignored); linear -> |a|"""
    import torch

    rng = np.random.default_rng(nc)
    v = rng.standard_normal(nc + 1)
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=3, flux_scheme=1, alpha=1.0, width=[g.width]))
    buf = torch.from_numpy(v).cuda()
    out = torch.full((1,), -1.0, dtype=torch.float64, device="cuda")
    for shift in (0, 1):
        fv.max_wavespeed_dev(buf.data_ptr() + 8 * shift, out.data_ptr())
        torch.cuda.synchronize()
        assert float(out.item()) == np.max(np.abs(v[shift:shift + nc]))
    fvl = pkg.fv.FV(pkg.fv.make_desc(nc, k=3, flux_model=1, flux_coef=(-2.5, 1.0), width=[g.width]))
    fvl.max_wavespeed_dev(buf.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    assert float(out.item()) == 2.5


def test_set_alpha_takes_effect_and_matches_oracle(gpu_lib, pkg, ref):
    """alpha installed with set_alpha (from the device reduction) gives the oracle's Lax-Friedrichs rhs for that alpha"""
    import torch

    nc = 3000
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    v = ex1_ic(g.center) + 1e-3 * np.random.default_rng(9).standard_normal(nc)
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=3, flux_scheme=1, alpha=0.25, width=[g.width]))
    vd = torch.from_numpy(v).cuda()
    alpha = pkg.slab.global_max_wavespeed(fv, vd, 1)
    assert alpha == np.max(np.abs(v))
    got = fv.rhs(0.0, v)
    want = ref.FV(pkg.fv.make_desc(nc, k=3, flux_scheme=1, alpha=alpha, width=[g.width])).rhs(0.0, v)
    assert np.array_equal(got, want)
    with pytest.raises(pkg._abi.HrwenoError):
        fv.set_alpha(-1.0)


# ---- the reference's host integrand, unchanged (tvdode.f90:50-57) --------------------------------------------
A_TVD = np.array([-1.0 + float(ii - 1) * 4 / 9 for ii in range(1, 11)])  # test_tvdode.f90:15


@pytest.mark.parametrize("order,rtol", [(1, 2e-2), (2, 1e-3), (3, 1e-3)])
def test_rktvd_host_integrand_reference_test(gpu_lib, pkg, order, rtol):
    """test/test_tvdode.f90:31-70 as written there: u' = a*u with the HOST integrand `fu(t, u, udot)`, nu = 10,
    t0 = -1.3, tout = 1.5, dt = (tout-t0)/3000, through rktvd(fu, nu, order); the reference's tolerances, and bitwise
    against the NumPy driver of the same steps (3001 steps)"""
    from oracle import np_oracle

    def fu(t, u, udot):
        udot[:] = A_TVD * u

    t0, tout = -1.3, 1.5
    dt = (tout - t0) / 3000
    ode = pkg.hrweno_tvdode.rktvd(fu, 10, order, host=True)
    u = np.full(10, 0.1)
    t = ode.integrate(u, t0, tout, dt)
    np.testing.assert_allclose(u, 0.1 * np.exp(A_TVD * (t - t0)), rtol=rtol)
    ur, tr = np_oracle.RK(lambda t_, u_: A_TVD * u_, order).integrate(np.full(10, 0.1), t0, tout, dt)
    assert t == tr and np.array_equal(u, ur) and ode.fevals == 3001 * order


def test_mstvd_host_integrand_reference_test(gpu_lib, pkg):
    """test/test_tvdode.f90:72-104: mstvd(fu, nu) with the host integrand, dt = 1e-3, rtol 1e-3; bitwise vs the NumPy driver"""
    from oracle import np_oracle

    def fu(t, u, udot):
        udot[:] = A_TVD * u

    ode = pkg.hrweno_tvdode.mstvd(fu, 10, host=True)
    u = np.full(10, 0.1)
    t = ode.integrate(u, -1.3, 1.5, 1e-3)
    np.testing.assert_allclose(u, 0.1 * np.exp(A_TVD * (t + 1.3)), rtol=1e-3)
    ur, tr = np_oracle.MS(lambda t_, u_: A_TVD * u_).integrate(np.full(10, 0.1), -1.3, 1.5, 1e-3)
    assert t == tr and np.array_equal(u, ur)


def test_example1_with_unchanged_host_rhs_equals_fused(gpu_lib, pkg, ref):
    """example1 as a reference user writes it: a HOST rhs calling w%reconstruct and godunov per face (example1:72-109),
    handed to rktvd(rhs, nc, 3).  Must equal the fused path (and so the oracle) bit for bit."""
    nc = 100
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    w = pkg.hrweno_weno.weno(nc, 3, 1e-6)
    flux = lambda v, x, t: (v ** 2) / 2  # noqa: E731  example1:120

    def rhs(t, v, vdot):
        vl, vr = w.reconstruct(np.ascontiguousarray(v))
        fed = np.empty(nc + 1)
        for i in range(1, nc):
            fed[i] = pkg.hrweno_fluxes.godunov(flux, vr[i - 1], vl[i], [g.right[i - 1]], t)
        fed[0], fed[nc] = fed[1], fed[nc - 1]
        vdot[:] = -(fed[1:] - fed[:-1]) / g.width

    ode = pkg.hrweno_tvdode.rktvd(rhs, nc, 3, host=True)
    fused = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=3, eps=1e-6, width=[g.width])), nc, 3)
    u, uf, t, tf = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
    for tout in (0.05, 0.12):
        t, tf = ode.integrate(u, t, tout, 1e-2), fused.integrate(uf, tf, tout, 1e-2)
        assert t == tf and np.array_equal(u, uf)
