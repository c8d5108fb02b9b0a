"""The REAL32 oracle against the reference's OWN SOURCE TEXT executed as its REAL32 build (src/hrweno_kinds.F90:9-10:
-DREAL32 makes rk = real32 for the whole library).

tests/golden/ref_exec_f32_*.npz come from `tests/golden/make_ref_exec_golden.py --real32`: tools/f90exec/f90py.py
translates the same source lines as for the binary64 fixtures (src/hrweno_{weno,fluxes,tvdode,grids}.f90, the two
example programs with their main programs) and executes them with every real entity a binary32 value -- literals,
scalars, arrays, `sum()` accumulators; a binary64 value reaching a store raises, so a silent promotion cannot hide.
The C oracle's REAL32 build (oracle/libhrweno_oracle_f32.so, the checker of the hrweno_*_f32 entry points on the GPU)
must reproduce them BIT FOR BIT: reconstruct (tables and cnu), calc_cnu, both numerical fluxes, rktvd 1-3 / mstvd with
the step counts the float32 time accumulation gives, grid1%linear, example1 as shipped over its 101 outputs, the
k x order sweep, the Lax-Friedrichs variant, example2 at 40x40 and example2 on geometric grids with growth terms."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("HRWENO_REFERENCE", "/root/reference")
F = np.float32


def gold(name):
    path = os.path.join(GOLD, f"ref_exec_f32_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


@pytest.fixture(scope="module")
def ref32(pkg):
    from oracle import ref32

    return ref32


@pytest.mark.parametrize("k", [1, 2, 3])
def test_real32_oracle_reconstruct_equals_reference_source(ref32, k):
    g = gold("reconstruct")
    for gname in ("none", "uniform", "cubic"):
        cnu = None
        if gname != "none":
            cnu = ref32.calc_cnu(g["xe_" + gname], k)
            assert cnu.dtype == F and np.array_equal(cnu, g[f"cnu_{gname}_k{k}"]), (gname, k)
        for vname in ("pulse", "rand"):
            vl, vr = ref32.reconstruct(g["v_" + vname], k, 1e-6, cnu=cnu)
            assert g[f"vl_{gname}_{vname}_k{k}"].dtype == F
            assert np.array_equal(vl, g[f"vl_{gname}_{vname}_k{k}"]) and np.array_equal(vr, g[f"vr_{gname}_{vname}_k{k}"]), (gname, vname, k)


def test_real32_oracle_fluxes_equal_reference_source(ref32):
    g = gold("fluxes")
    burgers = lambda u, x, t: (u * u) / F(2)  # example1:111-122 in real32  # noqa: E731
    for a, b, god, lf in zip(g["vm"], g["vp"], g["godunov"], g["lax_friedrichs"]):
        assert ref32.godunov(burgers, a, b, [3.0], 5.0) == god
        assert ref32.lax_friedrichs(burgers, a, b, [3.0], 5.0, g["alpha"]) == lf
    # the built-in flux models the fused kernels use (hrweno_ref_face_flux) are the same numbers
    assert np.array_equal(ref32.face_flux(0, 0, 1.0, 0.0, g["vm"], g["vp"]), g["godunov"])
    assert np.array_equal(ref32.face_flux(1, 0, 1.0, float(g["alpha"]), g["vm"], g["vp"]), g["lax_friedrichs"])


def test_real32_oracle_tvdode_equals_reference_source(ref32):
    g = gold("tvdode")
    a = g["a"]
    fu = lambda t, u: a * u  # noqa: E731
    for order in (1, 2, 3):
        ode = ref32.rktvd(fu, order, neq=10)
        u, t = np.ones(10, dtype=F), 0.0
        for tout in (0.0, 0.1, 0.1, 0.35):
            t = ode.integrate(u, t, tout, 1e-2)
        t = ode.integrate(u, t, 99.0, 1e-2, itask=2)
        assert np.array_equal(u, g[f"rk{order}_u"]) and F(t) == g[f"rk{order}_t"] and ode.fevals == int(g[f"rk{order}_fevals"]), order
    ode = ref32.mstvd(fu, neq=10)
    u, t = np.ones(10, dtype=F), 0.0
    for tout in (0.0, 0.1, 0.1, 0.35):
        t = ode.integrate(u, t, tout, 1e-2)
    assert np.array_equal(u, g["ms_u"]) and F(t) == g["ms_t"] and ode.fevals == int(g["ms_fevals"])


def _grid32(a, b, n):
    """grid1%linear evaluated in real32 (grids.f90:76-79,246-247), as tests/test_oracle_real32.py and bench.py write it"""
    rx = (F(b) - F(a)) / F(n)
    e = (F(a) + rx * np.arange(n + 1, dtype=F)).astype(F)
    return e, ((e[:-1] + e[1:]) / F(2)).astype(F), (e[1:] - e[:-1]).astype(F)


def test_real32_linear_grid_mirror_equals_reference_source():
    g = gold("grids")
    e, c, w = _grid32(-5.0, 5.0, 100)
    assert np.array_equal(e, g["linear"]) and np.array_equal(c, g["linear_center"]) and np.array_equal(w, g["linear_width"])


def test_real32_grid1_mirror_equals_reference_source(pkg):
    """hrweno_b200.real32.grid1 (the host-side set-up a REAL32 caller uses) against grid1 executed from source in real32:
    linear / geometric / bilinear bit for bit, log to 2 ulp of float32 (libm's logf / expf are not the path)"""
    g = gold("grids")
    G = pkg.real32.grid1
    lin = G().linear(-5.0, 5.0, 100)
    assert lin.edges.dtype == F and lin.center.dtype == F and lin.width.dtype == F
    assert np.array_equal(lin.edges, g["linear"]) and np.array_equal(lin.center, g["linear_center"]) and np.array_equal(lin.width, g["linear_width"])
    geo = G().geometric(1e1, 1e3, 1.1, 100)
    assert np.array_equal(geo.edges, g["geometric"]) and np.array_equal(geo.width, g["geometric_width"])
    assert np.array_equal(G().geometric(0.0, 10.0, 1.02, 24).edges, g["geometric_24"])
    bil = G().bilinear(0.0, 1e1, 1e3, [124, 365])
    assert np.array_equal(bil.edges, g["bilinear"]) and np.array_equal(bil.center, g["bilinear_center"])
    lg = G().log(1e-1, 1e3, 1000)
    assert lg.edges.dtype == F and np.max(np.abs(lg.edges / g["log"] - F(1))) <= 2 * np.finfo(F).eps
    # the example2 + growth fixture's grids are these
    g2 = gold("example2_growth")
    assert np.array_equal(G().geometric(0.0, 10.0, 1.02, 24).edges, g2["edges1"]) and np.array_equal(G().geometric(0.0, 10.0, 1.03, 18).edges, g2["edges2"])


def _example1(pkg, ref32, g_times, u0, width, k, order, scheme, snaps, final=None):
    """example1's driver loop (example1:55-65) in real32: time_out = time_end*ii/num_time_points in binary32.  `ref32` is
    the module that provides FV / rktvd / mstvd: the oracle here, the CUDA path (pkg.real32) in the GPU tests"""
    fv = ref32.FV(pkg.real32.make_desc(100, k=k, eps=1e-6, width=[width], flux_scheme=scheme, alpha=1.0))
    ode = ref32.rktvd(fv, order)
    u, t = u0.copy(), 0.0
    for ii in range(len(g_times)):
        t = ode.integrate(u, t, F(12.0) * F(ii) / F(100), 1e-2)
        assert F(t) == g_times[ii], f"output {ii}: t = {t!r}, reference source {g_times[ii]!r}"
        if ii in snaps:
            assert np.array_equal(u, snaps[ii]), f"output {ii}: max diff {np.max(np.abs(u - snaps[ii])):.3e}"
    if final is not None:
        assert np.array_equal(u, final)
    return ode


def test_real32_oracle_example1_equals_reference_source(pkg, ref32):
    g = gold("example1")
    assert np.array_equal(g["width"], _grid32(-5.0, 5.0, 100)[2])
    snaps = {ii: g[f"u_{ii}"] for ii in (0, 1, 10, 50, 100)}
    ode = _example1(pkg, ref32, g["times"], g["ic"], g["width"], 3, 3, 0, snaps)
    assert ode.fevals == int(g["fevals"])
    # the binary32 clock differs from the binary64 one (3603 evaluations): a different number of steps reaches t = 12
    assert int(g["fevals"]) != 3603 or g["times"][-1] != F(12.009999999999788)


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_real32_oracle_example1_k_order_sweep_equals_reference_source(pkg, ref32, k, order):
    g = gold("example1_sweep")
    w = _grid32(-5.0, 5.0, 100)[2]
    ode = _example1(pkg, ref32, g[f"t_k{k}_o{order}"], g["ic"], w, k, order, 0, {}, final=g[f"u_k{k}_o{order}"])
    assert ode.fevals == int(g[f"fevals_k{k}_o{order}"])


def test_real32_oracle_example1_lax_friedrichs_equals_reference_source(pkg, ref32):
    g = gold("example1_lf")
    snaps = {ii: g[f"u_{ii}"] for ii in (0, 10, 20)}
    ode = _example1(pkg, ref32, g["times"], g["ic"], g["width"], 3, 3, 1, snaps)
    assert ode.fevals == int(g["fevals"])


def g_of_t32(t):
    """the time factor of the REAL32 *_tfactor fixtures, 1.0_rk + 0.25_rk*t evaluated in binary32"""
    return float(F(1.0) + F(0.25) * F(t))


@pytest.mark.parametrize("order", [1, 2, 3])
def test_real32_oracle_time_dependent_flux_example1_equals_reference_source(pkg, ref32, order):
    """example1 executed in real32 with flux = (v**2)/2*(1 + t/4): stage times and the separable factor of the REAL32 oracle"""
    _example1_tfactor32(pkg, ref32, order)


def _example1_tfactor32(pkg, ref32, order):
    g = gold("example1_tfactor")
    fv = ref32.FV(pkg.real32.make_desc(100, k=3, eps=1e-6, width=[_grid32(-5.0, 5.0, 100)[2]]))
    fv.set_flux_time_fn(g_of_t32)
    ode = ref32.rktvd(fv, order)
    u, t = gold("example1")["ic"].copy(), 0.0
    for ii in range(21):
        t = ode.integrate(u, t, F(12.0) * F(ii) / F(100), 1e-2)
        assert F(t) == g[f"times_o{order}"][ii]
        if ii in (0, 10, 20):
            assert np.array_equal(u, g[f"u_{ii}_o{order}"]), f"order {order} output {ii}"
    assert ode.fevals == int(g[f"fevals_o{order}"])


def test_real32_oracle_time_dependent_growth_example2_equals_reference_source(pkg, ref32):
    _example2(pkg, ref32, gold("example2_growth_tfactor"), 24, 18, 2.5e-4, 0.5, growth=True, time_fn=g_of_t32)
    assert not np.array_equal(gold("example2_growth_tfactor")["u_20"], gold("example2_growth")["u_20"])


def _example2(pkg, ref32, g, n1, n2, dt, time_end, growth=False, time_fn=None):
    """example2's driver (example2:57-66) in real32, from the state its own `ic` gives (example2:48-52)"""
    e1, e2 = g["edges1"], g["edges2"]
    w1, w2, c1 = e1[1:] - e1[:-1], e2[1:] - e2[:-1], (e1[:-1] + e1[1:]) / F(2)
    fv = ref32.FV(pkg.real32.make_desc((n1, n2), k=3, eps=1e-6, flux_model=1, bc=1, width=[w1, w2]))
    if growth:
        fv.set_xedges(0, e1)
        fv.set_xedges(1, e2)
        fv.set_flux_coef(0, e1 * e1, None)  # flux1 = v*x(1)**2,   x = [right1(i), center2(j)]   example2:100-101,140
        fv.set_flux_coef(1, e2, c1)         # flux2 = v*x(1)*x(2), x = [center1(i), right2(j)]  example2:109-110,153
    if time_fn:
        fv.set_flux_time_fn(time_fn)
    ode = ref32.mstvd(fv)
    u, t = g["ic"].copy(), 0.0
    for ii in range(len(g["times"])):
        t = ode.integrate(u, t, F(time_end) * F(ii) / F(100), dt)
        assert F(t) == g["times"][ii], f"output {ii}: t = {t!r}, reference source {g['times'][ii]!r}"
        if f"u_{ii}" in g:
            assert np.array_equal(u, g[f"u_{ii}"]), f"output {ii}: max diff {np.max(np.abs(u - g[f'u_{ii}'])):.3e}"
    assert ode.fevals == int(g["fevals"])


def test_real32_oracle_example2_40x40_equals_reference_source(pkg, ref32):
    _example2(pkg, ref32, gold("example2_40"), 40, 40, 5e-3, 5.0)


def test_real32_oracle_example2_growth_on_geometric_grids_equals_reference_source(pkg, ref32):
    _example2(pkg, ref32, gold("example2_growth"), 24, 18, 2.5e-4, 0.5, growth=True)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is not on this machine")
def test_real32_fixtures_are_reproducible_and_oracle_equals_live_execution(ref32):
    """re-executes the source in real32 live: the committed reconstruct / fluxes / tvdode fixtures come out again, and
    on fresh random inputs (sizes the fixtures do not hold) the oracle equals the executed source"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_ref_exec_golden as G
    from f90py import FArr, callm, real_kind

    with real_kind(4):
        ns = G.load("example1_burgers_1d_fv.f90")
        for name, gen in (("reconstruct", G.gen_reconstruct), ("fluxes", G.gen_fluxes), ("tvdode", G.gen_tvdode)):
            live, g = gen(ns), gold(name)
            assert set(live) == set(g.files)
            for key in g.files:
                assert np.array_equal(np.asarray(live[key]), g[key]), (name, key)
        rng = np.random.default_rng(77)
        for nc, k in ((7, 1), (19, 2), (64, 3), (5, 3)):
            xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.2, 1.5, nc))]).astype(F)
            v = (rng.standard_normal(nc) * 10.0 ** rng.integers(-3, 4)).astype(F)
            for edges in (None, xe):
                w = ns["weno"](nc, k, F(1e-6)) if edges is None else ns["weno"](nc, k, F(1e-6), FArr(edges, (0,)))
                vl, vr = np.zeros(nc, dtype=F), np.zeros(nc, dtype=F)
                callm(w, "reconstruct", v, vl, vr)
                cnu = None if edges is None else ref32.calc_cnu(edges, k)
                a = ref32.reconstruct(v, k, 1e-6, cnu=cnu)
                assert np.array_equal(a[0], vl) and np.array_equal(a[1], vr), (nc, k, edges is None)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is not on this machine")
def test_real32_oracle_equals_executed_example_programs_on_small_grids(pkg, ref32):
    """both example programs executed from source in real32 at grid sizes the fixtures do not hold (down to 2 cells,
    n1 /= n2, every WENO order and RK order, with and without geometric grids + xedges + growth terms): the REAL32 oracle's
    fused rhs + integrators, bit for bit, including t and fevals"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_ref_exec_golden as m
    from f90py import real_kind

    with real_kind(4):
        for nc, k, order in [(2, 3, 3), (3, 2, 2), (4, 3, 1), (5, 1, 3), (7, 3, 3), (17, 2, 3), (33, 3, 2)]:
            live = m.run_example1(npts=4, snaps=(4,), k=k, order=order, nc=nc)
            assert np.array_equal(_grid32(-5.0, 5.0, nc)[2], live["width"])
            ode = ref32.rktvd(ref32.FV(pkg.real32.make_desc(nc, k=k, eps=1e-6, width=[live["width"]])), order)
            u, t = live["ic"].copy(), 0.0
            for ii in range(5):
                t = ode.integrate(u, t, F(12.0) * F(ii) / F(100), 1e-2)
                assert F(t) == live["times"][ii], (nc, k, order, ii)
            assert np.array_equal(u, live["u_4"]) and ode.fevals == live["fevals"], (nc, k, order)
        for n1, n2, growth in [(2, 2, False), (3, 5, False), (7, 4, False), (5, 3, True), (9, 11, True)]:
            kw = dict(grids="geometric", nonuniform=True, dt=2.5e-4, time_end=0.5, growth=True) if growth else {}
            live = m.run_example2(n1, 2, (0, 2), n2=n2, **kw)
            _example2(pkg, ref32, live, n1, n2, kw.get("dt", 5e-3), kw.get("time_end", 5.0), growth=growth)
