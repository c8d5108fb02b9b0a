#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): every hand-rolled synchronisation of the stage
kernels at small sizes -- mbarrier + TMA bulk / tensor loads, the exchange barrier, PDL, the single-CTA kernel, the
general kernel, pack/unpack, the chunk pipeline, and (with >= 2 GPUs) the peer-memory halo flags."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex1_ic, ex2_ic  # noqa: E402

pkg = graft.load_package()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
FAST, STRICT = pkg._abi.MODE_FAST, pkg._abi.MODE_STRICT


def steps_to(t, dt, k):
    for _ in range(k - 1):
        t = t + dt
    return t


if which in ("all", "1d"):
    for mode in (FAST, STRICT):
        nc = 3 * 1260 + 17
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
        u = ex1_ic(g.center) + 1e-3 * np.random.default_rng(1).standard_normal(nc)
        dt = 0.2 * 10.0 / nc
        for k in (2, 3):
            ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=k, linear=(-5.0, 5.0), mode=mode)), nc, 3)
            ode.integrate(u.copy(), 0.0, steps_to(0.0, dt, 2), dt)
        ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=3, linear=(-5.0, 5.0), mode=mode)), nc)
        ode.integrate(u.copy(), 0.0, steps_to(0.0, dt, 6), dt)
        # rows (half tiles), generic flux, width array
        rows = 3
        fv = pkg.fv.FV(pkg.fv.make_desc(4096, k=3, rows=rows, flux_scheme=1, alpha=1.1, linear=(-5.0, 5.0), mode=mode))
        pkg.hrweno_tvdode.rktvd(fv, 4096 * rows, 2).integrate(np.tile(ex1_ic(pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 4096).center), rows), 0.0, 0.0, 1e-4)
    print("1d ok", flush=True)

if which in ("all", "small"):
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    for mode in (FAST, STRICT):
        ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(100, width=[g.width], mode=mode)), 100, 3)
        ode.integrate(ex1_ic(g.center), 0.0, 0.05, 1e-2)
    print("small ok", flush=True)

if which in ("all", "2d"):
    n1, n2 = 150, 70
    g1, g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    u0 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(3).standard_normal((n2, n1))).reshape(-1)
    for mode in (FAST, STRICT):
        for kw in (dict(flux_model=1, bc=1), dict(flux_model=0, flux_scheme=1, alpha=1.2, bc=0)):
            fv = pkg.fv.FV(pkg.fv.make_desc((n1, n2), width=[g1.width, g2.width], mode=mode, **kw))
            pkg.hrweno_tvdode.mstvd(fv, n1 * n2).integrate(u0.copy(), 0.0, steps_to(0.0, 5e-3, 6), 5e-3)
            pkg.hrweno_tvdode.rktvd(fv, n1 * n2, 3).integrate(u0.copy(), 0.0, 0.0, 5e-3)
    print("2d ok", flush=True)

if which in ("all", "general"):
    nc = 700
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.002, nc)
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=3, width=[g.width]))
    fv.set_xedges(0, g.edges)
    pkg.hrweno_tvdode.rktvd(fv, nc, 3).integrate(ex1_ic(g.center), 0.0, 0.0, 1e-3)
    n1, n2 = 70, 50
    g1, g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.01, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    fv = pkg.fv.FV(pkg.fv.make_desc((n1, n2), flux_model=1, bc=1, width=[g1.width, g2.width]))
    fv.set_xedges(0, g1.edges)
    fv.set_flux_coef(0, g1.edges**2, None)
    fv.set_flux_coef(1, g2.edges, g1.center)
    pkg.hrweno_tvdode.mstvd(fv, n1 * n2).integrate(ex2_ic(g1.center, g2.center).reshape(-1), 0.0, steps_to(0.0, 1e-4, 5), 1e-4)
    w = pkg.hrweno_weno.weno(501, 3, 1e-6, mode=FAST)
    w.reconstruct(np.random.default_rng(0).standard_normal(501))
    print("general ok", flush=True)

if which in ("all", "pipeline"):
    os.environ["HRWENO_PIPE_CHUNK_TILES"] = "2"
    nc = 9 * 1260 + 5
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nc, k=3, linear=(-5.0, 5.0), mode=FAST)), nc, 3)
    ode.integrate(ex1_ic(g.center), 0.0, steps_to(0.0, 1e-4, 2), 1e-4)
    print("pipeline ok", flush=True)

if which in ("all", "mgpu") and pkg.lib().hrweno_device_count() >= 2:
    os.environ["HRWENO_PIPE_CHUNK_TILES"] = "29"
    nglob = 2 * 140000 + 77
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
    u = ex1_ic(g.center)
    m = pkg.mgpu.MultiGPU(pkg.fv.make_desc(nglob, k=3, linear=(-5.0, 5.0), mode=FAST), 2).rktvd(3)
    dt = 0.2 * 10.0 / nglob
    m.integrate(u, 0.0, steps_to(0.0, dt, 2), dt)        # host path: wide halos + per-slab pipeline
    m.upload(u)
    m.integrate_resident(0.0, steps_to(0.0, dt, 2), dt)  # resident path: in-kernel halo flags per stage
    m.max_wavespeed()
    n1, n2 = 130, 64
    g1, g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    m2 = pkg.mgpu.MultiGPU(pkg.fv.make_desc((n1, n2), flux_model=1, bc=1, width=[g1.width, g2.width]), 2).mstvd()
    m2.integrate(ex2_ic(g1.center, g2.center).reshape(-1).copy(), 0.0, steps_to(0.0, 5e-3, 6), 5e-3)
    print("mgpu ok", flush=True)
