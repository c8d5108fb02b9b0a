#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (not the bench.py line).

  cfg1: example1 as shipped (100 cells): latency per integrate call / per step
  cfg2: example2 as shipped (250x250 mstvd): time for the whole run
  cfg4: 2D linear advection WENO5+mstvd, 16384 x 16384 (or --n2d)
  cfg5: batched ensemble 65536 rows x 4096 cells, k in {1,2,3} x rktvd order in {1,2,3}
  general (only with --only general): the general stage kernel K7 (non-uniform grids, growth fluxes), 1D 2^22 cells
           rktvd3 and 2D 4096 x 4096 mstvd
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex1_ic, ex2_ic  # noqa: E402

import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n2d", type=int, default=16384)
ap.add_argument("--rows", type=int, default=65536)
ap.add_argument("--mode", default="strict")
ap.add_argument("--only", default="")
args = ap.parse_args()
pkg = graft.load_package()
MODE = pkg._abi.MODE_STRICT if args.mode == "strict" else pkg._abi.MODE_FAST
PEAK = 6538.9
stream = torch.cuda.current_stream().cuda_stream


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def steps_to(t, dt, k):
    tt = t
    for _ in range(k - 1):
        tt = tt + dt
    return tt


if args.only in ("", "cfg1"):
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(100, width=[g.width], mode=MODE)), 100, 3)
    u, t = ex1_ic(g.center), 0.0
    t = ode.integrate(u, t, 0.0, 1e-2)  # first call: also loads the kernels (lazy module loading), untimed
    w0 = time.perf_counter()
    for ii in range(1, 101):
        t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
    el = time.perf_counter() - w0
    print(f"cfg1 example1 ({args.mode}): 100 integrate calls, 1200 steps in {el*1e3:.1f} ms -> {el/1200*1e6:.1f} us/step, {el/100*1e6:.0f} us per integrate call (host u, 12 steps)")

if args.only in ("", "cfg2"):
    n = 250
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g.width, g.width], mode=MODE)), n * n)
    u, t = ex2_ic(g.center, g.center).reshape(-1), 0.0
    t = ode.integrate(u, t, 0.0, 5e-3)  # first call (4 RK3 start-up steps): also loads the kernels, untimed
    w0 = time.perf_counter()
    for ii in range(1, 101):
        t = ode.integrate(u, t, 5.0 * ii / 100, 5e-3)
    el = time.perf_counter() - w0
    print(f"cfg2 example2 ({args.mode}): 100 integrate calls, 997 multistep steps in {el*1e3:.1f} ms -> {el/997*1e6:.1f} us/step ({62500*997/el:.3e} cell-steps/s)")

if args.only in ("", "cfg4"):
    n = args.n2d
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    fv = pkg.fv.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g.width, g.width], mode=MODE))
    ode = pkg.hrweno_tvdode.mstvd(fv, n * n)
    c = g.center
    rng = np.random.default_rng(12345)
    u = ex2_ic(c, c)
    u += 1e-3 * rng.standard_normal(u.shape)
    ud = torch.from_numpy(u.reshape(-1)).cuda()
    dt = 0.125 * 10.0 / n
    t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt, 6), dt, 1, stream)  # start-up (4 RK3) + 2 MS steps
    K = 20
    el = timed(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt, K), dt, 1, stream))
    cells = n * n
    gbs = cells * 40.0 * K / el / 1e9
    print(f"cfg4 2D {n}x{n} WENO5+mstvd ({args.mode}): {K} steps in {el*1e3:.1f} ms -> {cells*K/el:.3e} cell-steps/s, {gbs:.0f} GB/s algorithmic (40 B/cell-step) = {gbs/PEAK:.3f} of measured HBM peak")
    del ode, fv, ud

if args.only in ("", "recon"):
    # reconstruct-only (K1r): w%reconstruct on device arrays, 2^27 cells, 24 algorithmic B/cell (R v, W vl, W vr)
    n = 1 << 27
    v = torch.from_numpy(np.random.default_rng(1).standard_normal(n)).cuda()
    vl, vr = torch.empty_like(v), torch.empty_like(v)
    for k in (1, 2, 3):
        w = pkg.hrweno_weno.weno(n, k, 1e-6, mode=MODE)
        w.reconstruct_dev(v.data_ptr(), vl.data_ptr(), vr.data_ptr(), stream=stream)
        el = timed(lambda: [w.reconstruct_dev(v.data_ptr(), vl.data_ptr(), vr.data_ptr(), stream=stream) for _ in range(10)]) / 10
        gbs = n * 24.0 / el / 1e9
        print(f"reconstruct-only 2^27 cells k={k} ({args.mode}): {n/el:.3e} cells/s, {gbs:.0f} GB/s algorithmic (24 B/cell) = {gbs/PEAK:.3f} of peak")
    del v, vl, vr

if args.only in ("", "cfg5"):
    rows, nc = args.rows, 4096
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    rng = np.random.default_rng(2024)
    va, vb = rng.uniform(0.5, 1.5, rows), rng.uniform(-1.0, 0.0, rows)
    xa, xb = rng.uniform(-4.5, -3.0, rows), rng.uniform(1.0, 3.0, rows)
    x = g.center[None, :]
    u0 = np.clip(va[:, None] + (vb - va)[:, None] / (xb - xa)[:, None] * (x - xa[:, None]), np.minimum(va, vb)[:, None], np.maximum(va, vb)[:, None])
    ud0 = torch.from_numpy(u0.reshape(-1)).cuda()
    dt = 0.1 * 10.0 / nc
    bytes_step = {1: 16.0, 2: 40.0, 3: 64.0}
    for k in (1, 2, 3):
        for order in (1, 2, 3):
            fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=k, rows=rows, width=[g.width], mode=MODE))
            ode = pkg.hrweno_tvdode.rktvd(fv, rows * nc, order)
            ud = ud0.clone()
            t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt, 3), dt, 1, stream)
            K = 10
            el = timed(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt, K), dt, 1, stream))
            cells = rows * nc
            gbs = cells * bytes_step[order] * K / el / 1e9
            print(f"cfg5 ensemble {rows}x{nc} k={k} rktvd{order} ({args.mode}): {cells*order*K/el:.3e} cell-stages/s, {gbs:.0f} GB/s algorithmic = {gbs/PEAK:.3f} of peak")
            del ode, fv, ud

if args.only == "general":
    # K7 (csrc/fvgen.cu): 1D geometric grid, per-cell tables streamed from HBM (96 B/cell at k=3 on top of the stage bytes)
    nc = 1 << 22
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.0 + 1e-7, nc)
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=3, width=[g.width]))
    fv.set_xedges(0, g.edges)
    ode = pkg.hrweno_tvdode.rktvd(fv, nc, 3)
    ud = torch.from_numpy(ex1_ic(g.center) + 1e-3 * np.random.default_rng(12345).standard_normal(nc)).cuda()
    dt = 0.1 * 10.0 / nc
    t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt, 3), dt, 1, stream)
    K = 20
    el = timed(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt, K), dt, 1, stream))
    gbs = nc * (64.0 + 3 * (96.0 + 8.0)) * K / el / 1e9  # 64 B/cell-step of state + per stage 96 B cnu + 8 B width
    print(f"general 1D 2^22 cells geometric grid WENO5+rktvd3 (K7, reference order): {nc*3*K/el:.3e} cell-stages/s, {gbs:.0f} GB/s algorithmic (64 + 3*104 B/cell-step) = {gbs/PEAK:.3f} of measured HBM peak")
    del ode, fv, ud
    # 2D geometric x geometric with the growth terms of example2:140,153: tile kernel in GEN mode, strict (reference order,
    # bit-identical) and fast (division-light weights, 1e-12)
    n = 4096
    g1 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.0005, n)
    g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.0003, n)
    u = ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(12345).standard_normal((n, n))
    for mname, m in (("reference order (strict)", pkg._abi.MODE_STRICT), ("fast", pkg._abi.MODE_FAST)):
        fv = pkg.fv.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g1.width, g2.width], mode=m))
        fv.set_xedges(0, g1.edges)
        fv.set_xedges(1, g2.edges)
        fv.set_flux_coef(0, g1.edges**2, None)
        fv.set_flux_coef(1, g2.edges, g1.center)
        ode = pkg.hrweno_tvdode.mstvd(fv, n * n)
        ud = torch.from_numpy(u.reshape(-1)).cuda()
        dt = 1e-6
        t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt, 6), dt, 1, stream)
        el = timed(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt, K), dt, 1, stream))
        gbs = n * n * 40.0 * K / el / 1e9
        print(f"general 2D {n}x{n} geometric grids + growth fluxes WENO5+mstvd (tile kernel GEN, {mname}): {n*n*K/el:.3e} cell-steps/s, {gbs:.0f} GB/s algorithmic (40 B/cell-step) = {gbs/PEAK:.3f} of measured HBM peak")
        del ode, fv, ud
