#!/usr/bin/env python
"""Generates tests/golden/ref_exec_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE TEXT.

The reference is Fortran and no Fortran compiler exists in this image, so the reference binary cannot be run.  What can
be done is what this script does: tools/f90exec/f90py.py translates the procedures of /root/reference/src/*.f90 and of
the two example programs -- their main programs included -- statement by statement into Python (no algorithm or driver
loop is restated by hand: operation order, `sum()`, sections, bounds, control flow come from the source lines; only the
formatted-output routine `output` is replaced by a hook that records `u` and `time`) and runs them on IEEE binary64.
Variants patch literals of the source text (grid size, k, order, dt), listed where they are applied.  The outputs are
committed as fixtures because /root/reference does not travel to the GPU box; tests/test_reference_source_exec.py holds
the C oracle (and through it the CUDA path) to these fixtures bit for bit and, when /root/reference is present,
re-executes a part of them live.

What this pins and what it does not: the oracle's operation order against the reference's source as an IEEE evaluation
with no contraction or re-association (gfortran on x86-64 without -ffast-math / -march=native); NOT a gfortran binary's
output (libm's exp/log in grid1%log, and any compiler-specific transformation, remain outside).

    python tests/golden/make_ref_exec_golden.py [--quick]      # --quick skips the two long runs (example2 fixtures)
    python tests/golden/make_ref_exec_golden.py --real32       # the REAL32 build of the same source -> ref_exec_f32_*.npz
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
OUT = os.environ.get("HRWENO_GOLDEN_OUT", HERE)  # where the fixtures are written (default: next to this script)
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


def R(x):
    """a harness-side real scalar of the kind being run (np.float32 in a REAL32 run, a Python float otherwise)"""
    return f90py.rl(x)


def A(x):
    """a harness-side real array of the kind being run"""
    return np.asarray(x, dtype=f90py.rdtype())


def pulse(nc=30):  # test/test_hrweno.f90:45-46
    v = np.zeros(nc, dtype=f90py.rdtype())
    v[nc // 3 - 1 : 2 * nc // 3] = 1.0
    return v


def gen_reconstruct(ns):
    """weno_init + weno_calc_cnu + weno_reconstruct (weno.f90:54-297) on the inputs of test/test_hrweno.f90 and random data"""
    out = {}
    rng = np.random.default_rng(20260102)
    nc = 30
    out["v_pulse"], out["v_rand"] = pulse(nc), A(rng.standard_normal(nc))
    out["xe_uniform"] = A([0.0 + 3.0 * i / nc for i in range(nc + 1)])   # test_hrweno.f90:87-93 style
    out["xe_cubic"] = A(np.array([i / nc for i in range(nc + 1)]) ** 3)  # test_hrweno.f90:128-135 style
    for k in (1, 2, 3):
        for gname in ("none", "uniform", "cubic"):
            w = ns["weno"](nc, k, R(1e-6)) if gname == "none" else ns["weno"](nc, k, R(1e-6), FArr(out["xe_" + gname], (0,)))
            if gname != "none":
                out[f"cnu_{gname}_k{k}"] = np.ascontiguousarray(np.transpose(w.cnu.a, (2, 1, 0)))  # [i-1, r+1, j]
            for vname in ("pulse", "rand"):
                vl, vr = A(np.zeros(nc)), A(np.zeros(nc))
                callm(w, "reconstruct", out["v_" + vname], vl, vr)
                out[f"vl_{gname}_{vname}_k{k}"], out[f"vr_{gname}_{vname}_k{k}"] = vl, vr
    return out


def gen_fluxes(ns):
    """lax_friedrichs / godunov (fluxes.f90:22-76) with Burgers' flux of example1:111-122"""
    rng = np.random.default_rng(20260103)
    vm, vp = A(rng.standard_normal(200)), A(rng.standard_normal(200))
    vm[:20] = vp[:20]  # h(a,a) = h(a) cases (test_fluxes.f90:41-47)
    x = FArr.from_list([3.0])
    god = A([ns["godunov"](ns["flux"], R(a), R(b), x, R(5.0)) for a, b in zip(vm, vp)])
    lf = A([ns["lax_friedrichs"](ns["flux"], R(a), R(b), x, R(5.0), R(1.3)) for a, b in zip(vm, vp)])
    return dict(vm=vm, vp=vp, godunov=god, lax_friedrichs=lf, alpha=R(1.3))


def gen_tvdode(ns):
    """rktvd orders 1-3 and mstvd (tvdode.f90:69-271) on the linear test ODE of test/test_tvdode.f90:107-111"""
    a = A([-1.0 + float(ii - 1) * 4 / (10 - 1) for ii in range(1, 11)])  # test_tvdode.f90:15

    def fu(t, u, udot):  # udot = a*u
        FArr.wrap(udot).assign(FArr(a) * FArr.wrap(u))

    out = {"a": a}
    for order in (1, 2, 3):
        ode = ns["rktvd"](fu, 10, order)
        u, t = FArr(A(np.ones(10))), Ref(R(0.0))
        for tout in (0.0, 0.1, 0.1, 0.35):
            callm(ode, "integrate", u, t, R(tout), R(1e-2))
        callm(ode, "integrate", u, t, R(99.0), R(1e-2), itask=2)  # single step
        out[f"rk{order}_u"], out[f"rk{order}_t"], out[f"rk{order}_fevals"] = u.a.copy(), t.v, ode.fevals
    ode = ns["mstvd"](fu, 10)
    u, t = FArr(A(np.ones(10))), Ref(R(0.0))
    for tout in (0.0, 0.1, 0.1, 0.35):
        callm(ode, "integrate", u, t, R(tout), R(1e-2))
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


class _Stop(Exception):
    """raised by the output hook to end a program after the outputs that are wanted"""


def run_program(ns, main, npts, snaps, state=("u", "time")):
    """run the example's OWN main program (translated like everything else).  The programs call `output(2)` after every
    `integrate` (example1:64, example2:65); `output` itself is formatted file I/O and is replaced by a hook that records
    the state -- nothing else of the program is touched.  Stops after output number `npts`."""
    rec = {"times": [], "n": 0}

    def output(action):
        if val_int(action) != 2:
            return
        ii = rec["n"]
        rec["times"].append(ns[state[1]])
        if ii in snaps:
            rec[f"u_{ii}"] = ns[state[0]].a.copy()
        rec["n"] += 1
        if ii >= npts:
            raise _Stop

    ns["output"] = output
    try:
        ns[main]()
    except _Stop:
        pass
    out = {k: v for k, v in rec.items() if k.startswith("u_")}
    out["times"], out["fevals"] = np.array(rec["times"]), ns["ode"].fevals
    return out


def val_int(x):
    return int(f90py.val(x))


def ex1_patch(k=3, order=3, nc=100):
    """the three literals of example1 that the sweeps vary (lines 34, 44, 53)"""
    p = []
    if nc != 100:
        p.append(("integer, parameter :: nc = 100", f"integer, parameter :: nc = {nc}"))
    if k != 3:
        p.append(("myweno = weno(ncells=nc, k=3, eps=1e-6_rk)", f"myweno = weno(ncells=nc, k={k}, eps=1e-6_rk)"))
    if order != 3:
        p.append(("ode = rktvd(rhs, nc, order=3)", f"ode = rktvd(rhs, nc, order={order})"))
    return p


def run_example1(npts=100, snaps=(0, 1, 50, 100), k=3, order=3, nc=100, extra_patch=()):
    """program example1_burgers_1d_fv, all of it, from the source"""
    ns = load("example1_burgers_1d_fv.f90", patch=ex1_patch(k, order, nc) + list(extra_patch))
    out = run_program(ns, "main_example1_burgers_1d_fv", npts, snaps)
    gx = ns["gx"]
    out.update(edges=gx.edges.a.copy(), width=gx.width.a.copy(), center=gx.center.a.copy())
    if f90py.rkind() == 4:  # the initial state as the program's own `ic` gives it (example1:49; output 0 is one step later)
        out["ic"] = ns["ic"](gx.center).a.copy()
    return out


def run_example2(n, npts, snaps, dt=5e-3, time_end=5.0, grids="linear", nonuniform=False, n2=None, growth=False, growth_patch=None):
    """program example_pbe_2d_fv from the source; `n`, `dt`, `time_end` patch its literals (lines 28, 57-58); grids =
    "geometric" swaps the two `linear` calls for `geometric` ones and hands the edges to `weno(...)` (what a user of
    non-uniform grids writes, weno.f90:100-112); growth removes the comment markers of lines 140, 153"""
    n2 = n if n2 is None else n2
    p = []
    if (n, n2) != (250, 250):
        p.append(("nc(2) = [250, 250]", f"nc(2) = [{n}, {n2}]"))
    if dt != 5e-3:
        p.append(("dt = 5e-3_rk", f"dt = {dt!r}_rk"))
    if time_end != 5.0:
        p.append(("time_end = 5.0_rk", f"time_end = {time_end!r}_rk"))
    if grids == "geometric":
        for i, ratio in ((1, 1.02), (2, 1.03)):
            p.append((f"call gx({i})%linear(xmin=0.0_rk, xmax=10.0_rk, ncells=nc({i}))",
                      f"call gx({i})%geometric(xmin=0.0_rk, xmax=10.0_rk, ratio={ratio!r}_rk, ncells=nc({i}))"))
    if nonuniform:
        for i in (1, 2):
            p.append((f"myweno({i}) = weno(ncells=nc({i}), k=k, eps=1e-6_rk)",
                      f"myweno({i}) = weno(ncells=nc({i}), k=k, eps=1e-6_rk, xedges=gx({i})%edges)"))
    if growth:
        p += growth_patch or GROWTH
    ns = load("example2_pbe_2d_fv.f90", patch=p)
    out = run_program(ns, "main_example_pbe_2d_fv", npts, snaps)
    out.update(edges1=ns["gx"][1].edges.a.copy(), edges2=ns["gx"][2].edges.a.copy())
    if f90py.rkind() == 4:  # example2:48-52
        c1, c2 = ns["gx"][1].center, ns["gx"][2].center
        out["ic"] = A([ns["ic"](FArr.from_list([c1[ii], c2[jj]])) for jj in range(1, n2 + 1) for ii in range(1, n + 1)])
    return out


# example1:98 is a stale commented alternative to line 99: `lax_friedrichs(flux, vr(i), vl(i+1), gx%r(i), t, alpha = 1.0_rk)`
# (`gx%r` no longer exists in grid1).  BASELINE.json's first config names Lax-Friedrichs, so that line is swapped in with
# its coordinate argument written as line 99 writes it.
LAX_FRIEDRICHS = [("fedges(i) = godunov(flux, vr(i), vl(i + 1), [gx%right(i)], t)",
                   "fedges(i) = lax_friedrichs(flux, vr(i), vl(i + 1), [gx%right(i)], t, alpha=1.0_rk)")]
GROWTH = [("flux1 = v !*x(1)**2", "flux1 = v*x(1)**2"), ("flux2 = v !*x(1)*x(2)", "flux2 = v*x(1)*x(2)")]
# t-dependent fluxes: the flux functions receive the evaluation time from the integrators through `rhs` (fluxes.f90:12-18,
# example1:99, example2:100-110) and ignore it as shipped; these patches multiply the shipped expressions by
# g(t) = 1 + t/4, which makes the stage times of tvdode.f90:162-166 (t, t + dt, t + dt/2) and :256 (t) visible in the result
TFACTOR1 = [("flux = (v**2)/2", "flux = (v**2)/2*(1.0_rk + 0.25_rk*t)")]
GROWTH_T = [("flux1 = v !*x(1)**2", "flux1 = v*x(1)**2*(1.0_rk + 0.25_rk*t)"), ("flux2 = v !*x(1)*x(2)", "flux2 = v*x(1)*x(2)*(1.0_rk + 0.25_rk*t)")]


def tfactor_fixtures(prefix="ref_exec_"):
    """ref_exec_example1_tfactor.npz / ref_exec_example2_growth_tfactor.npz: both programs with g(t) = 1 + t/4 on their fluxes
    (prefix "ref_exec_f32_" inside `with f90py.real_kind(4)`: the REAL32 build of the same patched sources)"""
    out = {}
    for order in (1, 2, 3):
        r = run_example1(npts=20, snaps=(0, 10, 20), order=order, extra_patch=TFACTOR1)
        for key in ("u_0", "u_10", "u_20", "times", "fevals"):
            out[f"{key}_o{order}"] = r[key]
    np.savez(os.path.join(OUT, prefix + "example1_tfactor.npz"), **out)
    np.savez(os.path.join(OUT, prefix + "example2_growth_tfactor.npz"),
             **run_example2(24, 20, (0, 10, 20), dt=2.5e-4, time_end=0.5, grids="geometric", nonuniform=True, n2=18, growth=True,
                            growth_patch=GROWTH_T))


def main32():
    """the REAL32 build of the reference (src/hrweno_kinds.F90:9-10, -DREAL32: rk = real32): the same source lines
    translated and executed with every real entity a binary32 value (f90py.real_kind(4)) -> ref_exec_f32_*.npz.
    Smaller than the binary64 set (NumPy float32 scalars are slow): it pins the REAL32 oracle
    (oracle/libhrweno_oracle_f32.so), which is the checker of the hrweno_*_f32 entry points."""
    t0 = time.time()
    with f90py.real_kind(4):
        ns1 = load("example1_burgers_1d_fv.f90")
        np.savez(os.path.join(OUT, "ref_exec_f32_reconstruct.npz"), **gen_reconstruct(ns1))
        np.savez(os.path.join(OUT, "ref_exec_f32_fluxes.npz"), **gen_fluxes(ns1))
        np.savez(os.path.join(OUT, "ref_exec_f32_tvdode.npz"), **gen_tvdode(ns1))
        g = ns1["new_grid1"]()
        callm(g, "linear", R(-5.0), R(5.0), 100)
        grids = {"linear": g.edges.a.copy(), "linear_center": g.center.a.copy(), "linear_width": g.width.a.copy()}
        callm(g, "geometric", R(1e1), R(1e3), R(1.1), 100)
        grids.update(geometric=g.edges.a.copy(), geometric_width=g.width.a.copy())
        callm(g, "geometric", R(0.0), R(10.0), R(1.02), 24)  # the grids of the example2 + growth fixture
        grids["geometric_24"] = g.edges.a.copy()
        callm(g, "bilinear", R(0.0), R(1e1), R(1e3), FArr(np.array([124, 365], dtype=np.int64)))  # test_grid.f90:128-160
        grids.update(bilinear=g.edges.a.copy(), bilinear_center=g.center.a.copy())
        callm(g, "log", R(1e-1), R(1e3), 1000)
        grids["log"] = g.edges.a.copy()
        np.savez(os.path.join(OUT, "ref_exec_f32_grids.npz"), **grids)
        if "--only-grids" in sys.argv:
            return
        print(f"real32: reconstruct / fluxes / tvdode / grids done ({time.time() - t0:.0f} s)", flush=True)
        np.savez(os.path.join(OUT, "ref_exec_f32_example1.npz"), **run_example1(npts=100, snaps=(0, 1, 10, 50, 100)))
        print(f"real32: example1 as shipped, 101 outputs done ({time.time() - t0:.0f} s)", flush=True)
        sweep = {}
        for k in (1, 2, 3):
            for order in (1, 2, 3):
                r = run_example1(npts=10, snaps=(0, 10), k=k, order=order)
                sweep["ic"] = r["ic"]
                sweep[f"u_k{k}_o{order}"], sweep[f"t_k{k}_o{order}"], sweep[f"fevals_k{k}_o{order}"] = r["u_10"], r["times"], r["fevals"]
        np.savez(os.path.join(OUT, "ref_exec_f32_example1_sweep.npz"), **sweep)
        np.savez(os.path.join(OUT, "ref_exec_f32_example1_lf.npz"),
                 **run_example1(npts=20, snaps=(0, 10, 20), extra_patch=LAX_FRIEDRICHS))
        print(f"real32: example1 k x order sweep and Lax-Friedrichs variant done ({time.time() - t0:.0f} s)", flush=True)
        np.savez(os.path.join(OUT, "ref_exec_f32_example2_growth.npz"),
                 **run_example2(24, 20, (0, 10, 20), dt=2.5e-4, time_end=0.5, grids="geometric", nonuniform=True, n2=18, growth=True))
        print(f"real32: example2 + growth on geometric 24x18 done ({time.time() - t0:.0f} s)", flush=True)
        tfactor_fixtures("ref_exec_f32_")
        print(f"real32: t-dependent fluxes done ({time.time() - t0:.0f} s)", flush=True)
        np.savez(os.path.join(OUT, "ref_exec_f32_example2_40.npz"), **run_example2(40, 10, (0, 1, 5, 10)))
        print(f"real32: example2 at 40x40, outputs 0..10 done ({time.time() - t0:.0f} s)", flush=True)


def main():
    if "--real32" in sys.argv:
        return main32()
    if "--only-tfactor" in sys.argv:
        return tfactor_fixtures()
    if "--only-tfactor32" in sys.argv:
        with f90py.real_kind(4):
            return tfactor_fixtures("ref_exec_f32_")
    quick = "--quick" in sys.argv
    t0 = time.time()
    ns1 = load("example1_burgers_1d_fv.f90")
    np.savez(os.path.join(OUT, "ref_exec_reconstruct.npz"), **gen_reconstruct(ns1))
    np.savez(os.path.join(OUT, "ref_exec_fluxes.npz"), **gen_fluxes(ns1))
    np.savez(os.path.join(OUT, "ref_exec_tvdode.npz"), **gen_tvdode(ns1))
    np.savez(os.path.join(OUT, "ref_exec_grids.npz"), **gen_grids(ns1))
    print(f"reconstruct / fluxes / tvdode / grids done ({time.time() - t0:.0f} s)", flush=True)
    np.savez(os.path.join(OUT, "ref_exec_example1.npz"), **run_example1())
    print(f"example1 as shipped, 101 outputs done ({time.time() - t0:.0f} s)", flush=True)
    sweep = {}
    for k in (1, 2, 3):  # the k x order sweep of BASELINE.json's ensemble config, on example1's problem (outputs 0..10)
        for order in (1, 2, 3):
            r = run_example1(npts=10, snaps=(10,), k=k, order=order)
            sweep[f"u_k{k}_o{order}"], sweep[f"t_k{k}_o{order}"], sweep[f"fevals_k{k}_o{order}"] = r["u_10"], r["times"], r["fevals"]
    np.savez(os.path.join(OUT, "ref_exec_example1_sweep.npz"), **sweep)
    np.savez(os.path.join(OUT, "ref_exec_example1_lf.npz"),
             **run_example1(npts=100, snaps=(0, 50, 100), extra_patch=LAX_FRIEDRICHS))
    print(f"example1 k x order sweep and Lax-Friedrichs variant done ({time.time() - t0:.0f} s)", flush=True)
    # example2's program with geometric grids, xedges and the growth terms its own comments hold (markers removed)
    np.savez(os.path.join(OUT, "ref_exec_example2_growth.npz"),
             **run_example2(24, 20, (0, 10, 20), dt=2.5e-4, time_end=0.5, grids="geometric", nonuniform=True, n2=18, growth=True))
    print(f"example2 + growth on geometric 24x18 done ({time.time() - t0:.0f} s)", flush=True)
    tfactor_fixtures()
    print(f"t-dependent fluxes (example1 x rktvd 1-3, example2 + growth x mstvd) done ({time.time() - t0:.0f} s)", flush=True)
    if quick:
        return
    np.savez(os.path.join(OUT, "ref_exec_example2_40.npz"), **run_example2(40, 100, (0, 1, 50, 100)))
    print(f"example2 at 40x40, 101 outputs done ({time.time() - t0:.0f} s)", flush=True)
    out = run_example2(250, 1, (0, 1))
    np.savez_compressed(os.path.join(OUT, "ref_exec_example2_250_first2.npz"), **out)
    print(f"example2 as shipped (250x250), outputs 0 and 1 done ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
