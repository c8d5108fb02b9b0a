"""GPU parity of the FAST-mode TILED kernels -- the instantiations bench.py times -- against the CPU oracle.

The <= 1024-cell fast tests of test_gpu_parity.py run the single-CTA kernel (fv1d_small.cu); everything here is
large enough for the persistent tiled kernels: fv1d_stage_kernel<K, {EULER, RK2_FINAL, RK3_S2, RK3_S3, MS}, Fast, ...>
with runs of 8 cells and the folded stage combinations (fv1d.cuh), its half-size tile for batched rows, and
fv2d_stage_kernel<..., Fast, UPW = 0|1, ...>.

Bar (BASELINE.json north_star): max|u_gpu - u_oracle| / max|u_oracle| <= 1e-12 at EVERY output time (the oracle follows
weno.f90:174-216, example1:97-107, example2:97-127, tvdode.f90:160-171,257 in the reference's operation order).
The measured drifts are appended to gpurun_out/fast_drift.txt when that directory exists (profiles/ keeps a copy).
"""
import os

import numpy as np
import pytest

from conftest import ROOT, ex1_ic, ex2_ic, normwise

pytestmark = pytest.mark.gpu

TOL = 1e-12  # north star: 1e-12 relative (normwise) in fp64 at every output time


def _log(line):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "fast_drift.txt"), "a") as f:
            f.write(line + "\n")


@pytest.fixture(scope="module")
def threads(ref):
    ref.set_threads(min(16, ref.max_threads()))
    yield
    ref.set_threads(1)


# ---- 1D, one row: the kernel of the headline benchmark ---------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nc", [2500, 100003, 1 << 20])
def test_fast_tiled_rktvd_1d_vs_oracle(gpu_lib, pkg, ref, threads, nc, k, order):
    """cfg3's arithmetic (Burgers + Godunov, linear grid through the width dictionary, ramp + noise, dt = 0.2 dx... the
    CFL of bench.py is 0.1): >= 60 steps in 6 integrate calls, every output compared"""
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(nc + 10 * k + order).standard_normal(nc)
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=k, eps=1e-6, linear=(-5.0, 5.0), mode=pkg._abi.MODE_FAST)), nc, order)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, k=k, eps=1e-6, width=[g.width])), order)
    u, ur, t, tr, worst = u0.copy(), u0.copy(), 0.0, 0.0, 0.0
    dt = 0.2 * 10.0 / nc
    nsteps = 240 if nc < 10000 else (120 if nc < 500000 else 66)
    for i in range(1, 7):
        tout = (i * nsteps // 6 - 0.5) * dt
        t = ode.integrate(u, t, tout, dt)
        tr = rode.integrate(ur, tr, tout, dt)
        assert t == tr
        worst = max(worst, normwise(u, ur))
    assert ode.fevals == rode.fevals == order * nsteps
    _log(f"1D rktvd{order} k={k} nc={nc} steps={nsteps}: {worst:.3e}")
    assert worst <= TOL, worst


@pytest.mark.parametrize("scheme,model", [(1, 0), (0, 1), (1, 1)])
def test_fast_tiled_generic_flux_1d_vs_oracle(gpu_lib, pkg, ref, threads, scheme, model):
    """the FK_GENERIC instantiations (Lax-Friedrichs, linear flux) and a streamed width array (geometric grid: more than
    256 distinct widths)"""
    nc = 60001
    for grid in ("linear", "geometric"):
        gg = pkg.hrweno_grids.grid1()
        g = gg.linear(-5.0, 5.0, nc) if grid == "linear" else gg.geometric(-5.0, 5.0, 1.00001, nc)
        u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(7).standard_normal(nc)
        kw = dict(k=3, eps=1e-6, flux_scheme=scheme, flux_model=model, flux_coef=(0.8, 1.0), alpha=1.1, bc=model, width=[g.width])
        ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nc, mode=pkg._abi.MODE_FAST, **kw)), nc, 3)
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, **kw)), 3)
        u, ur, t, tr, worst = u0.copy(), u0.copy(), 0.0, 0.0, 0.0
        dt = 0.2 * float(g.width.min())
        for i in range(1, 5):
            t = ode.integrate(u, t, (15 * i - 0.5) * dt, dt)
            tr = rode.integrate(ur, tr, (15 * i - 0.5) * dt, dt)
            assert t == tr
            worst = max(worst, normwise(u, ur))
        _log(f"1D rktvd3 k=3 nc={nc} {grid} scheme={scheme} model={model} steps=60: {worst:.3e}")
        assert worst <= TOL, worst


@pytest.mark.parametrize("k", [2, 3])
def test_fast_tiled_mstvd_1d_vs_oracle(gpu_lib, pkg, ref, threads, k):
    """fv1d_stage_kernel<K, C_MS, Fast>: the multistep combination with its two pointwise operands (tvdode.f90:257)"""
    nc = 100003
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(k).standard_normal(nc)
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=k, eps=1e-6, linear=(-5.0, 5.0), mode=pkg._abi.MODE_FAST)), nc)
    rode = ref.mstvd(ref.FV(pkg.fv.make_desc(nc, k=k, eps=1e-6, width=[g.width])))
    u, ur, t, tr, worst = u0.copy(), u0.copy(), 0.0, 0.0, 0.0
    dt = 0.1 * 10.0 / nc
    for i in range(1, 7):
        t = ode.integrate(u, t, (20 * i - 0.5) * dt, dt)
        tr = rode.integrate(ur, tr, (20 * i - 0.5) * dt, dt)
        assert t == tr
        worst = max(worst, normwise(u, ur))
    assert ode.fevals == rode.fevals
    _log(f"1D mstvd k={k} nc={nc} steps=120: {worst:.3e}")
    assert worst <= TOL, worst


# ---- 1D, batched rows (cfg5 scaled down): the half-size tile -----------------------------------------------------
def _ensemble_ic(rows, nc, seed=2024):
    """per-row ramp parameters va, vb, xa, xb from default_rng(2024)"""
    rng = np.random.default_rng(seed)
    va, vb = rng.uniform(0.5, 1.5, rows), rng.uniform(-1.0, 0.0, rows)
    xa, xb = rng.uniform(-4.5, -3.0, rows), rng.uniform(1.0, 3.0, rows)
    return va, vb, xa, xb


@pytest.mark.parametrize("mode", ["fast", "strict"])
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_ensemble_rows_half_tile_vs_oracle(gpu_lib, pkg, ref, threads, k, order, mode):
    """rows x 4096 cells: 4096-cell rows take the half-size tile (fv.cu: fv1d_half_tile), 9 tiles per row with a ragged
    last one; every row is an independent problem (copy-neighbour boundaries at both ends of each row)"""
    rows, nc = 37, 4096
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    va, vb, xa, xb = _ensemble_ic(rows, nc)
    x = g.center[None, :]
    u0 = np.clip(va[:, None] + (vb - va)[:, None] / (xb - xa)[:, None] * (x - xa[:, None]), vb[:, None], va[:, None])
    u0 = np.ascontiguousarray(u0)
    m = pkg._abi.MODE_FAST if mode == "fast" else pkg._abi.MODE_STRICT
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=k, rows=rows, linear=(-5.0, 5.0), mode=m)), nc * rows, order)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, k=k, rows=rows, width=[g.width])), order)
    u, ur, t, tr, worst = u0.reshape(-1).copy(), u0.reshape(-1).copy(), 0.0, 0.0, 0.0
    dt = 0.1 * 10.0 / nc
    for i in range(1, 5):
        t = ode.integrate(u, t, (20 * i - 0.5) * dt, dt)
        tr = rode.integrate(ur, tr, (20 * i - 0.5) * dt, dt)
        assert t == tr
        worst = max(worst, normwise(u, ur))
    _log(f"ensemble {rows}x{nc} rktvd{order} k={k} {mode} steps=80: {worst:.3e}")
    if mode == "strict":
        assert np.array_equal(u, ur)
    assert worst <= TOL, worst


# ---- 2D --------------------------------------------------------------------------------------------------------
CASES_2D = [
    # name, flux_model, flux_scheme, coef, bc
    ("upwind-linear-godunov", 1, 0, (1.0, 1.0), 1),     # example2's configuration: the UPW instantiation
    ("linear-godunov-negative", 1, 0, (1.0, -0.7), 1),  # a < 0 along x2: not upwind-specialisable
    ("burgers-godunov", 0, 0, (1.0, 1.0), 0),
    ("burgers-lax-friedrichs", 0, 1, (1.0, 1.0), 1),
]


@pytest.mark.parametrize("integrator", ["mstvd", "rktvd3"])
@pytest.mark.parametrize("case", CASES_2D, ids=[c[0] for c in CASES_2D])
def test_fast_tiled_2d_vs_oracle(gpu_lib, pkg, ref, threads, case, integrator):
    """fv2d_stage_kernel<3, ., Fast, UPW, 64, 32, 512> on a ragged grid of more than 512^2 cells, 60 steps"""
    name, model, scheme, coef, bc = case
    n1, n2 = 530, 517
    g1 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1)
    g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    rng = np.random.default_rng(n1 + model + 2 * scheme)
    u0 = (ex2_ic(g1.center, g2.center) + 1e-3 * rng.standard_normal((n2, n1))).reshape(-1)
    kw = dict(n=(n1, n2), k=3, eps=1e-6, flux_model=model, flux_scheme=scheme, flux_coef=coef, alpha=1.2, bc=bc, width=[g1.width, g2.width])
    fv, rfv = pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)), ref.FV(pkg.fv.make_desc(**kw))
    if integrator == "mstvd":
        ode, rode = pkg.hrweno_tvdode.mstvd(fv, n1 * n2), ref.mstvd(rfv)
    else:
        ode, rode = pkg.hrweno_tvdode.rktvd(fv, n1 * n2, 3), ref.rktvd(rfv, 3)
    u, ur, t, tr, worst = u0.copy(), u0.copy(), 0.0, 0.0, 0.0
    dt = 0.125 * 10.0 / n1
    for i in range(1, 5):
        t = ode.integrate(u, t, (15 * i - 0.5) * dt, dt)
        tr = rode.integrate(ur, tr, (15 * i - 0.5) * dt, dt)
        assert t == tr
        worst = max(worst, normwise(u, ur))
    _log(f"2D {n1}x{n2} {name} {integrator} steps=60: {worst:.3e}")
    assert worst <= TOL, worst


# ---- drift as a function of the number of steps and of the grid (example2, the configuration with the least margin) ----
def _example2_drift(pkg, ref, n, dt, t_end, nout):
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    kw = dict(n=(n, n), k=3, eps=1e-6, flux_model=1, bc=1, width=[g.width, g.width])
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)), n * n)
    rode = ref.mstvd(ref.FV(pkg.fv.make_desc(**kw)))
    u = ex2_ic(g.center, g.center).reshape(-1)
    ur, t, tr, curve = u.copy(), 0.0, 0.0, []
    for ii in range(1, nout + 1):
        tout = t_end * ii / nout
        t = ode.integrate(u, t, tout, dt)
        tr = rode.integrate(ur, tr, tout, dt)
        assert t == tr
        curve.append((ode.fevals, normwise(u, ur)))
    return curve


def test_fast_drift_vs_steps_example2(gpu_lib, pkg, ref, threads):
    """example2 with 4x the steps (dt/4 to the same end time: 4001 steps): the drift of the fast arithmetic must stay
    under the bar along the whole curve, not just at the shipped step count"""
    curve = _example2_drift(pkg, ref, 250, 5e-3 / 4, 5.0, 20)
    _log("example2 250^2 dt/4 (fevals, drift): " + " ".join(f"({f},{d:.2e})" for f, d in curve))
    assert max(d for _, d in curve) <= TOL, curve


def test_fast_drift_vs_grid_example2(gpu_lib, pkg, ref, threads):
    """example2's problem on 1000^2 cells at its CFL number (dt = 1.25e-3), 1000 steps"""
    curve = _example2_drift(pkg, ref, 1000, 1.25e-3, 1.25, 10)
    _log("example2 1000^2 (fevals, drift): " + " ".join(f"({f},{d:.2e})" for f, d in curve))
    assert max(d for _, d in curve) <= TOL, curve


# ---- 2D general operators (per-cell tables, growth fluxes, time factor) in FAST mode ---------------------------------------
@pytest.mark.parametrize("n,integrator", [(640, "ms"), (640, "rk3"), (300, "ms")])
def test_fast_general_2d_vs_oracle(gpu_lib, pkg, ref, threads, n, integrator):
    """non-uniform grids + growth terms (example2:140,153) + time factor on the tile kernel with division-light weights
    (fv2d.cu GEN, weno_core.cuh: weno_cell_nonuniform fast branch): <= 1e-12 normwise of the oracle at every output;
    640^2 takes the 32x32 tiles, 300^2 the small ones"""
    import math

    g1 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.003, n)
    g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.002, n)
    u0 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(8).standard_normal((n, n))).reshape(-1)
    kw = dict(n=(n, n), k=3, flux_model=1, bc=1, width=[g1.width, g2.width])
    fv, rfv = pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)), ref.FV(pkg.fv.make_desc(**kw))
    gt = lambda t: 1.0 + 0.3 * math.sin(2.0 * t)  # noqa: E731
    for f in (fv, rfv):
        f.set_xedges(0, g1.edges)
        f.set_xedges(1, g2.edges)
        f.set_flux_coef(0, 0.01 * g1.edges**2, None)
        f.set_flux_coef(1, 0.1 * g2.edges, 0.1 * g1.center)
        f.set_flux_time_fn(gt)
    if integrator == "ms":
        ode, rode = pkg.hrweno_tvdode.mstvd(fv, n * n), ref.mstvd(rfv)
    else:
        ode, rode = pkg.hrweno_tvdode.rktvd(fv, n * n, 3), ref.rktvd(rfv, 3)
    u, ur, t, tr, dt = u0.copy(), u0.copy(), 0.0, 0.0, 2e-3
    worst = 0.0
    for nsteps in (1, 9, 20):
        tt = t
        for _ in range(nsteps - 1):
            tt = tt + dt
        t, tr = ode.integrate(u, t, tt, dt), rode.integrate(ur, tr, tt, dt)
        assert t == tr
        worst = max(worst, normwise(u, ur))
    _log(f"general 2D fast {n}^2 {integrator} steps=30: {worst:.3e}")
    assert 0.0 < worst <= 1e-12  # > 0: the fast arithmetic really ran
