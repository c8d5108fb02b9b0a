"""REAL32 (src/hrweno_kinds.F90:9-17: rk = real32 is a compile-time switch of the whole reference).  The fp32 oracle is
oracle/hrweno_oracle.c compiled with every `double` read as `float` and single-precision literals
(oracle/libhrweno_oracle_f32.so).  Its first pin is the reference's source executed as its REAL32 build
(tests/test_reference_source_exec_real32.py, fixtures ref_exec_f32_*.npz).  This file is the second one, at sizes and on
inputs the fixtures do not hold: an independently written restatement, oracle/np_oracle.py evaluated on float32 arrays
(NumPy rounds every float32 operation to float32), must agree with it bit for bit -- tables, cnu, reconstruct, both
example right-hand sides, rktvd and mstvd including the step counts the float32 time accumulation produces."""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic

F = np.float32


@pytest.fixture(scope="module")
def ref32(pkg):
    from oracle import ref32

    return ref32


def _grid32(a, b, n):
    """grid1%linear evaluated in real32 (grids.f90:76-79,246-247)"""
    rx = (F(b) - F(a)) / F(n)
    e = (F(a) + rx * np.arange(n + 1, dtype=F)).astype(F)
    return e, ((e[:-1] + e[1:]) / F(2)).astype(F), (e[1:] - e[:-1]).astype(F)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_real32_reconstruct_and_cnu_two_restatements_agree(ref32, npo, k):
    rng = np.random.default_rng(k)
    n = 257
    for v in (rng.standard_normal(n).astype(F), np.where(np.arange(n) < n // 2, F(1), F(0)).astype(F), np.full(n, 0.7, dtype=F)):
        a, b = ref32.reconstruct(v, k, 1e-6), npo.reconstruct(v, k, 1e-6)
        assert b[0].dtype == F and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.1, 2.0, n))]).astype(F)
    ca, cb = ref32.calc_cnu(xe, k), npo.calc_cnu(xe, k)
    assert cb.dtype == F and np.array_equal(ca, cb)
    v = rng.standard_normal(n).astype(F)
    a, b = ref32.reconstruct(v, k, 1e-6, cnu=ca), npo.reconstruct(v, k, 1e-6, cnu=cb)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the reference's own unit test (test_hrweno.f90:104-108) in real32: cnu on a uniform grid equals c1/c2/c3
    xu = (F(3) * np.arange(31, dtype=F) / F(30)).astype(F)
    cu = ref32.calc_cnu(xu, k)
    d, c = npo.tables(k, F)
    assert np.allclose(cu, c[None, :, :], rtol=1e-4)


@pytest.mark.parametrize("kind", ["rk1", "rk2", "rk3", "ms"])
def test_real32_example1_type_run_two_restatements_agree(pkg, ref32, npo, kind):
    n = 200
    e, c, w = _grid32(-5.0, 5.0, n)
    u0 = ex1_ic(c.astype(np.float64)).astype(F)
    for scheme, sname in ((0, "godunov"), (1, "lax_friedrichs")):
        d = pkg.real32.make_desc(n, k=3, flux_scheme=scheme, alpha=1.25, width=[w])
        fv = ref32.FV(d)
        ode = ref32.mstvd(fv) if kind == "ms" else ref32.rktvd(fv, int(kind[2]))
        rhs = lambda t, u: npo.rhs1d(u, w, 3, 1e-6, scheme=sname, alpha=1.25)  # noqa: E731
        o2 = npo.MS(rhs) if kind == "ms" else npo.RK(rhs, int(kind[2]))
        u, u2, t, t2 = u0.copy(), u0.copy(), 0.0, 0.0
        for tout in (0.0, 0.02, 0.05):
            t = ode.integrate(u, t, tout, 1e-3)
            u2, t2 = o2.integrate(u2, t2, tout, 1e-3)
            assert u2.dtype == F and t == float(t2) and np.array_equal(u, u2)
        assert ode.fevals == o2.fevals


def test_real32_example2_type_rhs_and_growth_two_restatements_agree(pkg, ref32, npo):
    n1, n2 = 60, 47
    e1, c1, w1 = _grid32(0.0, 10.0, n1)
    e2, c2, w2 = _grid32(0.0, 10.0, n2)
    v = (ex2_ic(c1.astype(np.float64), c2.astype(np.float64)) + 1e-3 * np.random.default_rng(1).standard_normal((n2, n1))).astype(F)
    fv = ref32.FV(pkg.real32.make_desc((n1, n2), flux_model=1, bc=1, width=[w1, w2]))
    a = fv.rhs(0.0, v.reshape(-1)).reshape(n2, n1)
    b = npo.rhs2d(v, w1, w2, 3, 1e-6)
    assert b.dtype == F and np.array_equal(a, b)
    # growth terms of example2:140,153 on the same grids with per-cell tables
    fv.set_xedges(0, e1)
    fv.set_flux_coef(0, (e1 * e1).astype(F), None)
    fv.set_flux_coef(1, e2, c1)
    a = fv.rhs(0.0, v.reshape(-1)).reshape(n2, n1)
    b = npo.rhs2d(v, w1, w2, 3, 1e-6, cnu=(npo.calc_cnu(e1, 3), None), fcoef=((e1 * e1).astype(F), e2), ccoef=(None, c1))
    assert np.array_equal(a, b)


def test_real32_is_a_different_arithmetic_close_to_real64(pkg, ref, ref32):
    """sanity of the kind switch: same run in real64 differs, by single-precision rounding only"""
    n = 200
    e, c, w = _grid32(-5.0, 5.0, n)
    u0 = ex1_ic(c.astype(np.float64)).astype(F)
    o32 = ref32.rktvd(ref32.FV(pkg.real32.make_desc(n, k=3, width=[w])), 3)
    o64 = ref.rktvd(ref.FV(pkg.fv.make_desc(n, k=3, width=[w.astype(np.float64)])), 3)
    u32, u64 = u0.copy(), u0.astype(np.float64)
    o32.integrate(u32, 0.0, 0.1, 1e-3)
    o64.integrate(u64, 0.0, 0.1, 1e-3)
    err = np.max(np.abs(u32 - u64))
    assert 0 < err < 5e-5
