#!/usr/bin/env python
"""Worst normwise drift max|u_gpu - u_oracle| / max|u_oracle| over all output times of the shipped
configurations, for both modes.  Run on the GPU box: python tools/drift_report.py > profiles/<name>.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex1_ic, ex2_ic, normwise  # noqa: E402

pkg, ref = graft.load_package(), graft.load_oracle()
ref.set_threads(min(16, ref.max_threads()))
print("config                                   mode    worst normwise drift over 101 output times")
for k in (1, 2, 3):
    for scheme in (0, 1):
        for mode in (0, 1):
            g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
            kw = dict(n=100, k=k, eps=1e-6, flux_scheme=scheme, alpha=1.0, width=[g.width])
            ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(mode=mode, **kw)), 100, 3)
            rode = ref.rktvd(ref.FV(pkg.fv.make_desc(**kw)), 3)
            u, ur, t, tr, worst = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0, 0.0
            for ii in range(101):
                t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
                tr = rode.integrate(ur, tr, 12.0 * ii / 100, 1e-2)
                worst = max(worst, normwise(u, ur))
            print(f"example1 k={k} {'godunov' if scheme == 0 else 'lax-friedrichs':15s} rktvd3   {'strict' if mode == 0 else 'fast  '}  {worst:.3e}")
n = 250
g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
for mode in (0, 1):
    kw = dict(n=(n, n), k=3, eps=1e-6, flux_model=1, bc=1, width=[g.width, g.width])
    ode = pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc(mode=mode, **kw)), n * n)
    rode = ref.mstvd(ref.FV(pkg.fv.make_desc(**kw)))
    u = ex2_ic(g.center, g.center).reshape(-1)
    ur, t, tr, worst = u.copy(), 0.0, 0.0, 0.0
    for ii in range(101):
        t = ode.integrate(u, t, 5.0 * ii / 100, 5e-3)
        tr = rode.integrate(ur, tr, 5.0 * ii / 100, 5e-3)
        worst = max(worst, normwise(u, ur))
    print(f"example2 250x250 k=3 godunov mstvd        {'strict' if mode == 0 else 'fast  '}  {worst:.3e}")
