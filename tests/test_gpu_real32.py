"""REAL32 build of the path on the GPU (hrweno_*_f32, csrc/real32.cu; src/hrweno_kinds.F90:9-17) against the REAL32 build of
the oracle: bit-identical (the general kernels instantiated for float evaluate the reference's operation order in
separately rounded float operations).  Bar restated for fp32: 0 ULP against the fp32 oracle; against the fp64 run only
single-precision rounding (checked in tests/test_oracle_real32.py)."""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def ref32(pkg):
    from oracle import ref32

    return ref32


def _grid32(a, b, n):
    rx = (F(b) - F(a)) / F(n)
    e = (F(a) + rx * np.arange(n + 1, dtype=F)).astype(F)
    return e, ((e[:-1] + e[1:]) / F(2)).astype(F), (e[1:] - e[:-1]).astype(F)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 5, 257, 100003])
def test_real32_reconstruct_bitwise(gpu_lib, pkg, ref32, k, n):
    rng = np.random.default_rng(n + k)
    v = rng.standard_normal(n).astype(F)
    vl, vr = pkg.real32.weno(n, k, 1e-6).reconstruct(v)
    rl, rr = ref32.reconstruct(v, k, 1e-6)
    assert vl.dtype == F and np.array_equal(vl, rl) and np.array_equal(vr, rr)
    if n >= 5:
        xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.1, 2.0, n))]).astype(F)
        vl, vr = pkg.real32.weno(n, k, 1e-6, xedges=xe).reconstruct(v)
        rl, rr = ref32.reconstruct(v, k, 1e-6, cnu=ref32.calc_cnu(xe, k))
        assert np.array_equal(vl, rl) and np.array_equal(vr, rr)


@pytest.mark.parametrize("kind", ["rk1", "rk2", "rk3", "ms"])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_real32_1d_integrators_bitwise(gpu_lib, pkg, ref32, kind, k):
    n = 5003
    e, c, w = _grid32(-5.0, 5.0, n)
    u0 = (ex1_ic(c.astype(np.float64)) + 1e-3 * np.random.default_rng(k).standard_normal(n)).astype(F)
    for kw in (dict(width=[w]), dict(linear=(-5.0, 5.0)), dict(width=[w], flux_scheme=1, alpha=1.2), dict(width=[w], flux_model=1, flux_coef=(0.7, 1.0), bc=1)):
        fv, rfv = pkg.real32.FV(pkg.real32.make_desc(n, k=k, **kw)), ref32.FV(pkg.real32.make_desc(n, k=k, **kw))
        assert np.array_equal(fv.rhs(0.0, u0), rfv.rhs(0.0, u0))
        if kind == "ms":
            ode, rode = pkg.real32.mstvd(fv), ref32.mstvd(rfv)
        else:
            ode, rode = pkg.real32.rktvd(fv, int(kind[2])), ref32.rktvd(rfv, int(kind[2]))
        u, ur, t, tr = u0.copy(), u0.copy(), 0.0, 0.0
        dt = 0.2 * 10.0 / n
        for tout in (0.0, 7 * dt, 20 * dt):
            t = ode.integrate(u, t, tout, dt)
            tr = rode.integrate(ur, tr, tout, dt)
            assert t == tr and np.array_equal(u, ur), (kind, k, list(kw))
        assert ode.fevals == rode.fevals


def test_real32_rows_and_2d_bitwise(gpu_lib, pkg, ref32):
    rows, nc = 7, 1000
    e, c, w = _grid32(-5.0, 5.0, nc)
    r0 = (ex1_ic(c.astype(np.float64))[None, :] * np.linspace(0.5, 1.5, rows)[:, None]).astype(F).reshape(-1)
    kw = dict(k=3, rows=rows, width=[w])
    ode, rode = pkg.real32.rktvd(pkg.real32.FV(pkg.real32.make_desc(nc, **kw)), 3), ref32.rktvd(ref32.FV(pkg.real32.make_desc(nc, **kw)), 3)
    u, ur = r0.copy(), r0.copy()
    assert ode.integrate(u, 0.0, 0.01, 1e-3) == rode.integrate(ur, 0.0, 0.01, 1e-3) and np.array_equal(u, ur)
    n1, n2 = 130, 97
    e1, c1, w1 = _grid32(0.0, 10.0, n1)
    e2, c2, w2 = _grid32(0.0, 10.0, n2)
    v0 = (ex2_ic(c1.astype(np.float64), c2.astype(np.float64)) + 1e-3 * np.random.default_rng(2).standard_normal((n2, n1))).astype(F).reshape(-1)
    g = lambda t: 1.0 + 0.5 * np.sin(3.0 * float(t))  # noqa: E731
    for general in (False, True):
        for kind in ("ms", "rk3"):
            kw = dict(flux_model=1, bc=1, width=[w1, w2])
            fv, rfv = pkg.real32.FV(pkg.real32.make_desc((n1, n2), **kw)), ref32.FV(pkg.real32.make_desc((n1, n2), **kw))
            if general:  # geometric-type tables along x1, growth terms (example2:140,153) and a time factor
                xe = (e1 * (F(1) + F(0.01) * e1)).astype(F)
                for f in (fv, rfv):
                    f.set_xedges(0, xe)
                    f.set_flux_coef(0, (e1 * e1).astype(F), None)
                    f.set_flux_coef(1, e2, c1)
                    f.set_flux_time_fn(g)
            assert np.array_equal(fv.rhs(0.3, v0), rfv.rhs(0.3, v0))
            ode, rode = (pkg.real32.mstvd(fv), ref32.mstvd(rfv)) if kind == "ms" else (pkg.real32.rktvd(fv, 3), ref32.rktvd(rfv, 3))
            u, ur = v0.copy(), v0.copy()
            dt = 1e-4 if general else 5e-3  # the growth terms (coefficients up to 100) need the smaller step to stay bounded
            t, tr = ode.integrate(u, 0.0, 7.5 * dt, dt), rode.integrate(ur, 0.0, 7.5 * dt, dt)
            assert np.all(np.isfinite(ur)) and t == tr and np.array_equal(u, ur), (general, kind)


def test_real32_example1_every_output_time(gpu_lib, pkg, ref32):
    """example1 as shipped, evaluated in real32: 101 integrate calls; the float32 accumulation t = t + dt decides the steps"""
    nc = 100
    e, c, w = _grid32(-5.0, 5.0, nc)
    u0 = ex1_ic(c.astype(np.float64)).astype(F)
    ode = pkg.real32.rktvd(pkg.real32.FV(pkg.real32.make_desc(nc, k=3, width=[w])), 3)
    rode = ref32.rktvd(ref32.FV(pkg.real32.make_desc(nc, k=3, width=[w])), 3)
    u, ur, t, tr = u0.copy(), u0.copy(), 0.0, 0.0
    for ii in range(101):
        tout = float(F(12.0) * F(ii) / F(100))  # time_end*ii/num_time_points in real(rk) (example1:62)
        t, tr = ode.integrate(u, t, tout, 1e-2), rode.integrate(ur, tr, tout, 1e-2)
        assert t == tr and np.array_equal(u, ur), ii
    assert ode.fevals == rode.fevals


def test_real32_validation(gpu_lib, pkg):
    with pytest.raises(pkg.HrwenoError) as ei:
        pkg.real32.weno(10, 3, 1e-8)  # eps <= epsilon(1.0_real32) = 1.19e-7 (weno.f90:92-96 in real32)
    assert ei.value.status == pkg._abi.EINVAL
    with pytest.raises(pkg.HrwenoError):
        pkg.real32.weno(0, 3, 1e-6)
    d = pkg.real32.make_desc(100, linear=(-5.0, 5.0))
    d.nranks = 2
    with pytest.raises(pkg.HrwenoError):
        pkg.real32.FV(d)
