#!/usr/bin/env python
"""Generates tests/golden/ref_exec_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE TEXT.

The reference is Fortran and no Fortran compiler exists in this image, so the reference binary cannot be run.  What can
be done is what this script does: tools/f90exec/f90py.py translates the procedures of /root/reference/src/*.f90 and of
the two example programs statement by statement into Python (no algorithm is restated by hand -- operation order,
`sum()`, sections, bounds, control flow come from the source lines) and runs them on IEEE binary64.  The outputs are
committed as fixtures because /root/reference does not travel to the GPU box; tests/test_reference_source_exec.py holds
the C oracle (and through it the CUDA path) to these fixtures bit for bit and, when /root/reference is present,
re-executes a part of them live.

What this pins and what it does not: the oracle's operation order against the reference's source as an IEEE evaluation
with no contraction or re-association (gfortran on x86-64 without -ffast-math / -march=native); NOT a gfortran binary's
output (libm's exp/log in grid1%log, and any compiler-specific transformation, remain outside).

    python tests/golden/make_ref_exec_golden.py [--quick]      # --quick skips the two long runs (example2 fixtures)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools", "f90exec"))
import f90py  # noqa: E402
from f90py import FArr, Ref, callm  # noqa: E402

REF = os.environ.get("HRWENO_REFERENCE", "/root/reference")
SRC = ("src/hrweno_weno.f90", "src/hrweno_fluxes.f90", "src/hrweno_tvdode.f90", "src/hrweno_grids.f90")


def load(example=None, patch=None):
    """translate the four library modules (+ one example program's procedures); `patch` = [(old, new)] applied to the
    example's text first (used only to remove the comment marker in front of the growth terms of example2:140,153)"""
    P = f90py.Program()
    for f in SRC:
        P.add_source(os.path.join(REF, f))
    if example:
        path = os.path.join(REF, "example", example)
        if patch:
            text = open(path).read()
            for old, new in patch:
                assert text.count(old) == 1, old
                text = text.replace(old, new)
            path = os.path.join("/tmp", "f90exec_" + example)
            open(path, "w").write(text)
        P.add_source(path, skip=("output", "timer"))
    return P.build()


def pulse(nc=30):  # test/test_hrweno.f90:45-46
    v = np.zeros(nc)
    v[nc // 3 - 1 : 2 * nc // 3] = 1.0
    return v


def gen_reconstruct(ns):
    """weno_init + weno_calc_cnu + weno_reconstruct (weno.f90:54-297) on the inputs of test/test_hrweno.f90 and random data"""
    out = {}
    rng = np.random.default_rng(20260102)
    nc = 30
    out["v_pulse"], out["v_rand"] = pulse(nc), rng.standard_normal(nc)
    out["xe_uniform"] = np.array([0.0 + 3.0 * i / nc for i in range(nc + 1)])   # test_hrweno.f90:87-93 style
    out["xe_cubic"] = np.array([i / nc for i in range(nc + 1)]) ** 3            # test_hrweno.f90:128-135 style
    for k in (1, 2, 3):
        for gname in ("none", "uniform", "cubic"):
            w = ns["weno"](nc, k, 1e-6) if gname == "none" else ns["weno"](nc, k, 1e-6, FArr(out["xe_" + gname], (0,)))
            if gname != "none":
                out[f"cnu_{gname}_k{k}"] = np.ascontiguousarray(np.transpose(w.cnu.a, (2, 1, 0)))  # [i-1, r+1, j]
            for vname in ("pulse", "rand"):
                vl, vr = np.zeros(nc), np.zeros(nc)
                callm(w, "reconstruct", out["v_" + vname], vl, vr)
                out[f"vl_{gname}_{vname}_k{k}"], out[f"vr_{gname}_{vname}_k{k}"] = vl, vr
    return out


def gen_fluxes(ns):
    """lax_friedrichs / godunov (fluxes.f90:22-76) with Burgers' flux of example1:111-122"""
    rng = np.random.default_rng(20260103)
    vm, vp = rng.standard_normal(200), rng.standard_normal(200)
    vm[:20] = vp[:20]  # h(a,a) = h(a) cases (test_fluxes.f90:41-47)
    x = FArr.from_list([3.0])
    god = np.array([ns["godunov"](ns["flux"], float(a), float(b), x, 5.0) for a, b in zip(vm, vp)])
    lf = np.array([ns["lax_friedrichs"](ns["flux"], float(a), float(b), x, 5.0, 1.3) for a, b in zip(vm, vp)])
    return dict(vm=vm, vp=vp, godunov=god, lax_friedrichs=lf, alpha=1.3)


def gen_tvdode(ns):
    """rktvd orders 1-3 and mstvd (tvdode.f90:69-271) on the linear test ODE of test/test_tvdode.f90:107-111"""
    a = np.array([-1.0 + float(ii - 1) * 4 / (10 - 1) for ii in range(1, 11)])  # test_tvdode.f90:15

    def fu(t, u, udot):  # udot = a*u
        FArr.wrap(udot).assign(FArr(a) * FArr.wrap(u))

    out = {"a": a}
    for order in (1, 2, 3):
        ode = ns["rktvd"](fu, 10, order)
        u, t = FArr(np.ones(10)), Ref(0.0)
        for tout in (0.0, 0.1, 0.1, 0.35):
            callm(ode, "integrate", u, t, tout, 1e-2)
        callm(ode, "integrate", u, t, 99.0, 1e-2, itask=2)  # single step
        out[f"rk{order}_u"], out[f"rk{order}_t"], out[f"rk{order}_fevals"] = u.a.copy(), t.v, ode.fevals
    ode = ns["mstvd"](fu, 10)
    u, t = FArr(np.ones(10)), Ref(0.0)
    for tout in (0.0, 0.1, 0.1, 0.35):
        callm(ode, "integrate", u, t, tout, 1e-2)
    out["ms_u"], out["ms_t"], out["ms_fevals"] = u.a.copy(), t.v, ode.fevals
    return out


def gen_grids(ns):
    out = {}
    g = ns["new_grid1"]()
    callm(g, "linear", -5.0, 5.0, 100)
    out["linear"] = g.edges.a.copy()
    callm(g, "geometric", 1e1, 1e3, 1.1, 100)  # test_grid.f90:94-126
    out["geometric"] = g.edges.a.copy()
    callm(g, "bilinear", 0.0, 1e1, 1e3, FArr(np.array([124, 365], dtype=np.int64)))  # test_grid.f90:128-160
    out["bilinear"] = g.edges.a.copy()
    callm(g, "log", 1e-1, 1e3, 1000)
    out["log"] = g.edges.a.copy()
    return out


def run_example1(ns, npts=100, snaps=(0, 1, 50, 100), k=3, order=3, nc=100):
    """program example1 (example1:31-65): the driver loop is these few lines, everything it calls is translated source
    (nc, k and order are the literals of lines 34, 44 and 53)"""
    gx = ns["new_grid1"]()
    callm(gx, "linear", -5.0, 5.0, nc)                    # example1:41
    ns["nc"], ns["gx"] = nc, gx
    ns["myweno"] = ns["weno"](nc, k, 1e-6)                 # example1:44
    u = ns["ic"](gx.center)                               # example1:50
    ode = ns["rktvd"](ns["rhs"], nc, order)                # example1:53
    time_end, dt, t = 12.0, 1e-2, Ref(0.0)                 # example1:56-58
    out, times = {}, []
    for ii in range(npts + 1):                             # example1:61-65
        time_out = time_end * ii / 100
        callm(ode, "integrate", u, t, time_out, dt)
        times.append(t.v)
        if ii in snaps:
            out[f"u_{ii}"] = u.a.copy()
    out.update(times=np.array(times), fevals=ode.fevals, edges=gx.edges.a.copy(), width=gx.width.a.copy(), center=gx.center.a.copy())
    return out


def setup_example2(ns, n1, n2, grids="linear", nonuniform=False):
    nc = FArr(np.array([n1, n2], dtype=np.int64))
    gx, myweno = FArr(np.empty(2, dtype=object)), FArr(np.empty(2, dtype=object))
    for i, n in ((1, n1), (2, n2)):
        g = ns["new_grid1"]()
        if grids == "linear":
            callm(g, "linear", 0.0, 10.0, n)               # example2:38-39
        else:
            callm(g, "geometric", 0.0, 10.0, (1.02, 1.03)[i - 1], n)
        gx[i] = g
        myweno[i] = ns["weno"](n, 3, 1e-6, g.edges) if nonuniform else ns["weno"](n, 3, 1e-6)  # example2:42-43
    ns["nc"], ns["gx"], ns["myweno"] = nc, gx, myweno
    u = FArr(np.zeros(n1 * n2))
    for j in range(1, n2 + 1):                             # example2:49-51
        for i in range(1, n1 + 1):
            u[(j - 1) * n1 + i] = ns["ic"](FArr.from_list([gx[1].center[i], gx[2].center[j]]))
    return gx, u


def run_example2(ns, n, npts, snaps, dt=5e-3, time_end=5.0, grids="linear", nonuniform=False, n2=None):
    """program example2 (example2:25-69)"""
    n2 = n if n2 is None else n2
    gx, u = setup_example2(ns, n, n2, grids, nonuniform)
    ode = ns["mstvd"](ns["rhs"], n * n2)                   # example2:54
    t, out, times = Ref(0.0), {}, []
    for ii in range(npts + 1):                             # example2:61-66
        time_out = time_end * ii / 100
        callm(ode, "integrate", u, t, time_out, dt)
        times.append(t.v)
        if ii in snaps:
            out[f"u_{ii}"] = u.a.copy()
    out.update(times=np.array(times), fevals=ode.fevals, edges1=gx[1].edges.a.copy(), edges2=gx[2].edges.a.copy())
    return out


# example1:98 is a stale commented alternative to line 99: `lax_friedrichs(flux, vr(i), vl(i+1), gx%r(i), t, alpha = 1.0_rk)`
# (`gx%r` no longer exists in grid1).  BASELINE.json's first config names Lax-Friedrichs, so that line is swapped in with
# its coordinate argument written as line 99 writes it.
LAX_FRIEDRICHS = [("fedges(i) = godunov(flux, vr(i), vl(i + 1), [gx%right(i)], t)",
                   "fedges(i) = lax_friedrichs(flux, vr(i), vl(i + 1), [gx%right(i)], t, alpha=1.0_rk)")]
GROWTH = [("flux1 = v !*x(1)**2", "flux1 = v*x(1)**2"), ("flux2 = v !*x(1)*x(2)", "flux2 = v*x(1)*x(2)")]


def main():
    quick = "--quick" in sys.argv
    t0 = time.time()
    ns1 = load("example1_burgers_1d_fv.f90")
    np.savez(os.path.join(HERE, "ref_exec_reconstruct.npz"), **gen_reconstruct(ns1))
    np.savez(os.path.join(HERE, "ref_exec_fluxes.npz"), **gen_fluxes(ns1))
    np.savez(os.path.join(HERE, "ref_exec_tvdode.npz"), **gen_tvdode(ns1))
    np.savez(os.path.join(HERE, "ref_exec_grids.npz"), **gen_grids(ns1))
    print(f"reconstruct / fluxes / tvdode / grids done ({time.time() - t0:.0f} s)", flush=True)
    np.savez(os.path.join(HERE, "ref_exec_example1.npz"), **run_example1(ns1))
    print(f"example1 as shipped, 101 outputs done ({time.time() - t0:.0f} s)", flush=True)
    sweep = {}
    for k in (1, 2, 3):  # the k x order sweep of BASELINE.json's ensemble config, on example1's problem (outputs 0..10)
        for order in (1, 2, 3):
            r = run_example1(load("example1_burgers_1d_fv.f90"), npts=10, snaps=(10,), k=k, order=order)
            sweep[f"u_k{k}_o{order}"], sweep[f"t_k{k}_o{order}"], sweep[f"fevals_k{k}_o{order}"] = r["u_10"], r["times"], r["fevals"]
    np.savez(os.path.join(HERE, "ref_exec_example1_sweep.npz"), **sweep)
    np.savez(os.path.join(HERE, "ref_exec_example1_lf.npz"),
             **run_example1(load("example1_burgers_1d_fv.f90", patch=LAX_FRIEDRICHS), npts=100, snaps=(0, 50, 100)))
    print(f"example1 k x order sweep and Lax-Friedrichs variant done ({time.time() - t0:.0f} s)", flush=True)
    # example2's program with geometric grids, xedges and the growth terms its own comments hold (markers removed)
    nsg = load("example2_pbe_2d_fv.f90", patch=GROWTH)
    np.savez(os.path.join(HERE, "ref_exec_example2_growth.npz"),
             **run_example2(nsg, 24, 20, (0, 10, 20), dt=2.5e-4, time_end=0.5, grids="geometric", nonuniform=True, n2=18))
    print(f"example2 + growth on geometric 24x18 done ({time.time() - t0:.0f} s)", flush=True)
    if quick:
        return
    ns2 = load("example2_pbe_2d_fv.f90")
    np.savez(os.path.join(HERE, "ref_exec_example2_40.npz"), **run_example2(ns2, 40, 100, (0, 1, 50, 100)))
    print(f"example2 at 40x40, 101 outputs done ({time.time() - t0:.0f} s)", flush=True)
    ns2 = load("example2_pbe_2d_fv.f90")
    out = run_example2(ns2, 250, 1, (0, 1))
    np.savez_compressed(os.path.join(HERE, "ref_exec_example2_250_first2.npz"), **out)
    print(f"example2 as shipped (250x250), outputs 0 and 1 done ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
