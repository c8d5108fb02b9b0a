"""The oracle against the reference's OWN SOURCE TEXT, executed mechanically.

tests/golden/ref_exec_*.npz were produced by tests/golden/make_ref_exec_golden.py: tools/f90exec/f90py.py translates
the procedures of /root/reference/src/*.f90 and the two example programs (main programs included; only their formatted
`output` routine is replaced by a recording hook) statement by statement into Python and runs
them on IEEE binary64 (no algorithm restated by hand; see the docstrings there for the arithmetic model and its limits --
it is the source's operation order, not a gfortran binary).  Here:

  * CPU (always): the C oracle -- the checker of every GPU parity test -- reproduces those fixtures BIT FOR BIT:
    reconstruct (uniform tables and cnu), calc_cnu, both numerical fluxes, rktvd 1-3 / mstvd incl. itask = 2, fevals and
    the strict is_done, grid1 (linear / bilinear / geometric; log to 2 ulp: libm), example1 as shipped at outputs
    0, 1, 50, 100 (+ all 101 output times), example2 at 40x40 over all 101 outputs and as shipped (250x250) at its first
    two outputs, and example2's program on geometric grids with `xedges` and the growth terms of its comments (K7);
  * CPU, only where /root/reference exists: a part of the fixtures is re-generated live and must be identical;
  * GPU (tests/test_zzz_gpu_reference_source.py): the CUDA path through the C ABI against the same fixtures directly.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, ex2_ic

GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("HRWENO_REFERENCE", "/root/reference")


def gold(name):
    path = os.path.join(GOLD, f"ref_exec_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


# ---- reconstruct / cnu ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3])
def test_oracle_reconstruct_equals_reference_source(ref, npo, k):
    g = gold("reconstruct")
    for gname in ("none", "uniform", "cubic"):
        cnu = None
        if gname != "none":
            cnu = ref.calc_cnu(g["xe_" + gname], k)
            assert np.array_equal(cnu, g[f"cnu_{gname}_k{k}"]), f"calc_cnu {gname} k={k}"
            assert np.array_equal(npo.calc_cnu(g["xe_" + gname], k), g[f"cnu_{gname}_k{k}"])
        for vname in ("pulse", "rand"):
            vl, vr = ref.reconstruct(g["v_" + vname], k, 1e-6, cnu=cnu)
            assert np.array_equal(vl, g[f"vl_{gname}_{vname}_k{k}"]) and np.array_equal(vr, g[f"vr_{gname}_{vname}_k{k}"]), (gname, vname)
            nl, nr = npo.reconstruct(g["v_" + vname], k, 1e-6, cnu)
            assert np.array_equal(nl, vl) and np.array_equal(nr, vr)


def test_oracle_fluxes_equal_reference_source(ref):
    g = gold("fluxes")
    assert np.array_equal(ref.face_flux(0, 0, 1.0, 0.0, g["vm"], g["vp"]), g["godunov"])
    assert np.array_equal(ref.face_flux(1, 0, 1.0, float(g["alpha"]), g["vm"], g["vp"]), g["lax_friedrichs"])
    burgers = lambda v, x, t: (v * v) / 2  # noqa: E731
    assert ref.godunov(burgers, float(g["vm"][50]), float(g["vp"][50]), [3.0], 5.0) == g["godunov"][50]
    assert ref.lax_friedrichs(burgers, float(g["vm"][50]), float(g["vp"][50]), [3.0], 5.0, 1.3) == g["lax_friedrichs"][50]


def test_oracle_tvdode_equals_reference_source(ref):
    g = gold("tvdode")
    for order in (1, 2, 3):
        ode = ref.rktvd("test_ode", order, neq=10)
        u, t = np.ones(10), 0.0
        for tout in (0.0, 0.1, 0.1, 0.35):
            t = ode.integrate(u, t, tout, 1e-2)
        t = ode.integrate(u, t, 99.0, 1e-2, itask=2)
        assert t == float(g[f"rk{order}_t"]) and ode.fevals == int(g[f"rk{order}_fevals"])
        assert np.array_equal(u, g[f"rk{order}_u"]), f"rktvd order {order}"
    ode = ref.mstvd("test_ode", neq=10)
    u, t = np.ones(10), 0.0
    for tout in (0.0, 0.1, 0.1, 0.35):
        t = ode.integrate(u, t, tout, 1e-2)
    assert t == float(g["ms_t"]) and ode.fevals == int(g["ms_fevals"]) and np.array_equal(u, g["ms_u"])


def test_grid_mirrors_equal_reference_source(pkg, ref):
    g = gold("grids")
    G = pkg.hrweno_grids.grid1
    assert np.array_equal(G().linear(-5.0, 5.0, 100).edges, g["linear"])
    assert np.array_equal(ref.grid_linear(-5.0, 5.0, 100)[0], g["linear"])
    assert np.array_equal(G().geometric(1e1, 1e3, 1.1, 100).edges, g["geometric"])
    assert np.array_equal(G().bilinear(0.0, 1e1, 1e3, (124, 365)).edges, g["bilinear"])
    assert np.max(np.abs(G().log(1e-1, 1e3, 1000).edges / g["log"] - 1.0)) <= 2 * np.finfo(float).eps


# ---- the example programs --------------------------------------------------------------------------------------------
def _example1(pkg, make_ode, g, upto=100):
    """example1's driver loop (example1:55-65) on `make_ode(desc)`; checks every stored output"""
    assert np.array_equal(pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100).width, g["width"])
    ode = make_ode(pkg.fv.make_desc(100, k=3, eps=1e-6, width=[g["width"]]))
    u, t = np.clip(1.0 + (-1.5 / 6.0) * (g["center"] + 4.0), -0.5, 1.0), 0.0
    for ii in range(upto + 1):
        t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
        assert t == g["times"][ii], f"output {ii}: t = {t!r}, reference source {g['times'][ii]!r}"
        if f"u_{ii}" in g:
            assert np.array_equal(u, g[f"u_{ii}"]), f"output {ii}: max diff {np.max(np.abs(u - g[f'u_{ii}'])):.3e}"
    return ode


def test_oracle_example1_equals_reference_source(pkg, ref):
    g = gold("example1")
    ode = _example1(pkg, lambda d: ref.rktvd(ref.FV(d), 3), g)
    assert ode.fevals == int(g["fevals"]) == 3603 and repr(float(g["times"][-1])) == "12.009999999999788"


def g_of_t(t):
    """the time factor of the *_tfactor fixtures (make_ref_exec_golden.py: TFACTOR1, GROWTH_T)"""
    return 1.0 + 0.25 * t


def _example1_variant(pkg, make_ode, g_times, g_u, k, order, scheme=0, upto=10, snaps=None):
    grid = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    ode = make_ode(pkg.fv.make_desc(100, k=k, eps=1e-6, width=[grid.width], flux_scheme=scheme, alpha=1.0), order)
    u, t = np.clip(1.0 + (-1.5 / 6.0) * (grid.center + 4.0), -0.5, 1.0), 0.0
    for ii in range(upto + 1):
        t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
        assert t == g_times[ii]
        if snaps and ii in snaps:
            assert np.array_equal(u, snaps[ii]), f"output {ii}: max diff {np.max(np.abs(u - snaps[ii])):.3e}"
    if g_u is not None:
        assert np.array_equal(u, g_u), f"k={k} order={order}: max diff {np.max(np.abs(u - g_u)):.3e}"
    return ode


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_oracle_example1_k_order_sweep_equals_reference_source(pkg, ref, k, order):
    """WENO k = 1..3 x rktvd order 1..3 (the sweep of BASELINE.json's ensemble config) on example1's problem"""
    g = gold("example1_sweep")
    ode = _example1_variant(pkg, lambda d, o: ref.rktvd(ref.FV(d), o), g[f"t_k{k}_o{order}"], g[f"u_k{k}_o{order}"], k, order)
    assert ode.fevals == int(g[f"fevals_k{k}_o{order}"])


def test_oracle_example1_lax_friedrichs_equals_reference_source(pkg, ref):
    """BASELINE.json's first config names Lax-Friedrichs: example1 with its stale line 98 swapped in (alpha = 1)"""
    g = gold("example1_lf")
    snaps = {ii: g[f"u_{ii}"] for ii in (0, 50, 100)}
    ode = _example1_variant(pkg, lambda d, o: ref.rktvd(ref.FV(d), o), g["times"], None, 3, 3, scheme=1, upto=100, snaps=snaps)
    assert ode.fevals == int(g["fevals"]) == 3603


def _example2(pkg, make_ode, g, n1, n2, dt, time_end, growth=False, mod=None, time_fn=None):
    e1, e2 = g["edges1"], g["edges2"]
    w1, w2, c1, c2 = e1[1:] - e1[:-1], e2[1:] - e2[:-1], (e1[:-1] + e1[1:]) / 2, (e2[:-1] + e2[1:]) / 2
    fv = mod.FV(pkg.fv.make_desc((n1, n2), k=3, eps=1e-6, flux_model=1, bc=1, width=[w1, w2]))
    if growth:
        fv.set_xedges(0, e1)
        fv.set_xedges(1, e2)
        fv.set_flux_coef(0, e1 * e1, None)  # flux1 = v*x(1)**2,   x = [right1(i), center2(j)]   example2:100-101,140
        fv.set_flux_coef(1, e2, c1)         # flux2 = v*x(1)*x(2), x = [center1(i), right2(j)]  example2:109-110,153
    if time_fn:
        fv.set_flux_time_fn(time_fn)        # ... *g(t): the flux functions' third argument, fluxes.f90:12-18
    ode = make_ode(fv)
    u, t = ex2_ic(c1, c2).reshape(-1), 0.0
    for ii in range(len(g["times"])):
        t = ode.integrate(u, t, time_end * ii / 100, dt)
        assert t == g["times"][ii], f"output {ii}: t = {t!r}, reference source {g['times'][ii]!r}"
        if f"u_{ii}" in g:
            assert np.array_equal(u, g[f"u_{ii}"]), f"output {ii}: max diff {np.max(np.abs(u - g[f'u_{ii}'])):.3e}"
    assert ode.fevals == int(g["fevals"])


def test_oracle_example2_40x40_equals_reference_source(pkg, ref):
    g = gold("example2_40")
    _example2(pkg, ref.mstvd, g, 40, 40, 5e-3, 5.0, mod=ref)
    assert repr(float(g["times"][-1])) == "5.0049999999999155" and int(g["fevals"]) == 1009


def test_oracle_example2_as_shipped_first_outputs_equal_reference_source(pkg, ref):
    g = gold("example2_250_first2")
    assert np.array_equal(g["edges1"], pkg.hrweno_grids.grid1().linear(0.0, 10.0, 250).edges)
    ref.set_threads(min(8, ref.max_threads()))
    try:
        _example2(pkg, ref.mstvd, g, 250, 250, 5e-3, 5.0, mod=ref)
    finally:
        ref.set_threads(1)


def test_oracle_example2_growth_on_geometric_grids_equals_reference_source(pkg, ref):
    """the general path's semantics (per-axis xedges, f = (v*cross)*face) against example2's source with the comment
    markers in front of its growth terms removed"""
    g = gold("example2_growth")
    assert np.array_equal(g["edges1"], pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.02, 24).edges)
    assert np.array_equal(g["edges2"], pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.03, 18).edges)
    _example2(pkg, ref.mstvd, g, 24, 18, 2.5e-4, 0.5, growth=True, mod=ref)


def _example1_tfactor(pkg, FV, make_ode, order):
    """example1 with flux = (v**2)/2*g(t): the closed-set form model(v)*g(t) through set_flux_time_fn"""
    g = gold("example1_tfactor")
    grid = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    fv = FV(pkg.fv.make_desc(100, k=3, eps=1e-6, width=[grid.width]))
    fv.set_flux_time_fn(g_of_t)
    ode = make_ode(fv, order)
    u, t = np.clip(1.0 + (-1.5 / 6.0) * (grid.center + 4.0), -0.5, 1.0), 0.0
    for ii in range(21):
        t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
        assert t == g[f"times_o{order}"][ii]
        if ii in (0, 10, 20):
            assert np.array_equal(u, g[f"u_{ii}_o{order}"]), f"order {order} output {ii}: max diff {np.max(np.abs(u - g[f'u_{ii}_o{order}'])):.3e}"
    assert ode.fevals == int(g[f"fevals_o{order}"])
    return u


@pytest.mark.parametrize("order", [1, 2, 3])
def test_oracle_time_dependent_flux_example1_equals_reference_source(pkg, ref, order):
    """t-dependent fluxes: the reference's integrators hand t, t + dt and t + dt/2 to the rhs (tvdode.f90:162-166);
    example1 executed from source with its flux multiplied by g(t) = 1 + t/4 pins the stage times at which the oracle (and
    through it the CUDA path) evaluates the separable factor of hrweno_fv_set_flux_time_fn"""
    u = _example1_tfactor(pkg, ref.FV, lambda fv, o: ref.rktvd(fv, o), order)
    assert not np.array_equal(u, gold("example1")["u_0"])
    if order == 3:  # the factor matters: without it the state at output 20 is example1's own
        plain = ref.rktvd(ref.FV(pkg.fv.make_desc(100, k=3, eps=1e-6, width=[pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100).width])), 3)
        v, t = np.clip(1.0 + (-1.5 / 6.0) * (pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100).center + 4.0), -0.5, 1.0), 0.0
        for ii in range(21):
            t = plain.integrate(v, t, 12.0 * ii / 100, 1e-2)
        assert not np.array_equal(u, v)


def test_oracle_time_dependent_growth_example2_equals_reference_source(pkg, ref):
    """example2 on geometric grids with v*x(1)**2*g(t) and v*x(1)*x(2)*g(t), multi-step integrator (rhs at t, tvdode.f90:256;
    start-up by four RK3 single steps, :236-247)"""
    _example2(pkg, ref.mstvd, gold("example2_growth_tfactor"), 24, 18, 2.5e-4, 0.5, growth=True, mod=ref, time_fn=g_of_t)
    assert not np.array_equal(gold("example2_growth_tfactor")["u_20"], gold("example2_growth")["u_20"])


# ---- live re-execution where the reference tree is present ---------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is not on this machine")
def test_fixtures_are_reproducible_from_the_reference_tree():
    sys.path.insert(0, GOLD)
    import make_ref_exec_golden as m

    ns = m.load("example1_burgers_1d_fv.f90")
    for name, fn in (("reconstruct", m.gen_reconstruct), ("fluxes", m.gen_fluxes), ("tvdode", m.gen_tvdode), ("grids", m.gen_grids)):
        g, live = gold(name), fn(ns)
        for key, v in live.items():
            assert np.array_equal(np.asarray(v), g[key]), (name, key)
    g, live = gold("example1"), m.run_example1(npts=3, snaps=(0, 1))
    assert np.array_equal(live["u_0"], g["u_0"]) and np.array_equal(live["u_1"], g["u_1"])
    assert np.array_equal(live["times"], g["times"][:4])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is not on this machine")
def test_oracle_equals_executed_reference_source_on_random_inputs(ref):
    """beyond the frozen fixtures: random sizes (including ncells smaller than the stencil), orders, smoothing factors,
    value scales and non-uniform grids -- reconstruct and calc_cnu of the executed source vs the C oracle, bit for bit"""
    sys.path.insert(0, GOLD)
    import make_ref_exec_golden as m
    from f90py import FArr, callm

    ns = m.load()
    rng = np.random.default_rng(777)
    for trial in range(120):
        nc = int(rng.choice([1, 2, 3, 4, 5, 7, 16, 33]))
        k = int(rng.integers(1, 4))
        eps = float(10.0 ** rng.uniform(-12, -2))
        scale = float(10.0 ** rng.uniform(-6, 6))
        v = scale * rng.standard_normal(nc)
        if trial % 3 == 0:
            v[rng.integers(0, nc)] = 0.0
        nonuniform = trial % 2 == 1 and nc >= 2
        if nonuniform:
            xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.1, 3.0, nc))])
            w = ns["weno"](nc, k, eps, FArr(xe, (0,)))
            cnu = ref.calc_cnu(xe, k)
            assert np.array_equal(cnu, np.transpose(w.cnu.a, (2, 1, 0))), (trial, nc, k)
        else:
            w, cnu = ns["weno"](nc, k, eps), None
        vl, vr = np.zeros(nc), np.zeros(nc)
        callm(w, "reconstruct", v, vl, vr)
        rl, rr = ref.reconstruct(v, k, eps, cnu=cnu)
        assert np.array_equal(vl, rl) and np.array_equal(vr, rr), (trial, nc, k, eps, nonuniform)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is not on this machine")
def test_oracle_equals_executed_example1_program_on_small_and_odd_grids(pkg, ref):
    """example1's program (grid1%linear, ic, rhs, godunov, rktvd) executed from source at grid sizes down to 2 cells,
    every WENO order and RK order: the oracle's fused rhs + integrators, bit for bit, including t and fevals"""
    sys.path.insert(0, GOLD)
    import make_ref_exec_golden as m

    for nc, k, order in [(2, 3, 3), (3, 2, 2), (4, 3, 1), (5, 1, 3), (7, 3, 3), (17, 2, 3), (33, 3, 2)]:
        live = m.run_example1(npts=4, snaps=(4,), k=k, order=order, nc=nc)
        grid = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
        assert np.array_equal(grid.width, live["width"]) and np.array_equal(grid.center, live["center"])
        ode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, k=k, eps=1e-6, width=[grid.width])), order)
        u, t = np.clip(1.0 + (-1.5 / 6.0) * (grid.center + 4.0), -0.5, 1.0), 0.0
        for ii in range(5):
            t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
            assert t == live["times"][ii]
        assert np.array_equal(u, live["u_4"]) and ode.fevals == live["fevals"], (nc, k, order)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is not on this machine")
def test_oracle_equals_executed_example2_program_on_small_rectangular_grids(pkg, ref):
    """example2's program (2D split rhs with strided column sections, zero-flux walls, mstvd start-up + multistep) executed
    from source on small n1 /= n2 grids, with and without geometric grids + xedges + growth terms"""
    sys.path.insert(0, GOLD)
    import make_ref_exec_golden as m

    for n1, n2, growth in [(2, 2, False), (3, 5, False), (7, 4, False), (12, 9, False), (5, 3, True), (9, 11, True)]:
        kw = dict(grids="geometric", nonuniform=True, dt=2.5e-4, time_end=0.5, growth=True) if growth else {}
        live = m.run_example2(n1, 2, (0, 2), n2=n2, **kw)
        _example2(pkg, ref.mstvd, live, n1, n2, kw.get("dt", 5e-3), kw.get("time_end", 5.0), growth=growth, mod=ref)
