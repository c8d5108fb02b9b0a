#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/hrweno_oracle.c).

These are ORACLE-generated regression anchors: the reference is Fortran, cannot be run in this image and ships no
golden vectors of its own.  They freeze the oracle's answers so that a later change to the
oracle or to the CUDA path that alters a single bit is caught on CPU and on GPU.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex1_ic, ex2_ic, pulse  # noqa: E402

pkg, ref = graft.load_package(), graft.load_oracle()

# reconstruct: the pulse of test_hrweno.f90:45-46 and a seeded random vector, k = 1..3
out = {}
rng = np.random.default_rng(20260101)
vr_ = rng.standard_normal(64)
for k in (1, 2, 3):
    for name, v in (("pulse", pulse(30)), ("rand", vr_)):
        vl, vr = ref.reconstruct(v, k, 1e-6)
        out[f"{name}_k{k}_vl"], out[f"{name}_k{k}_vr"] = vl, vr
out["rand_v"] = vr_
np.savez(os.path.join(HERE, "reconstruct.npz"), **out)

# example1 as shipped: u at outputs 0, 50, 100
g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
ode = ref.rktvd(ref.FV(pkg.fv.make_desc(100, width=[g.width])), 3)
u, t, snaps, times = ex1_ic(g.center), 0.0, {}, []
for ii in range(101):
    t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
    times.append(t)
    if ii in (0, 50, 100):
        snaps[f"u_{ii}"] = u.copy()
np.savez(os.path.join(HERE, "example1.npz"), times=np.array(times), **snaps)

# example2 as shipped: rows 20..29 and 60..69 of the final state + per-output mass and extrema
n = 250
g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
ref.set_threads(min(8, ref.max_threads()))
ode = ref.mstvd(ref.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g.width, g.width])))
u, t, times, mass, umin, umax = ex2_ic(g.center, g.center).reshape(-1), 0.0, [], [], [], []
for ii in range(101):
    t = ode.integrate(u, t, 5.0 * ii / 100, 5e-3)
    times.append(t)
    U = u.reshape(n, n)
    mass.append(float(np.sum(U * g.width[None, :] * g.width[:, None])))
    umin.append(U.min())
    umax.append(U.max())
U = u.reshape(n, n)
np.savez(os.path.join(HERE, "example2.npz"), times=np.array(times), mass=np.array(mass), umin=np.array(umin), umax=np.array(umax),
         rows_20_29=U[20:30].copy(), rows_60_69=U[60:70].copy(), row_sums=U.sum(axis=1))
print("golden fixtures written")
