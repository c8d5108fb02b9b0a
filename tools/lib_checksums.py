#!/usr/bin/env python
"""Result checksums of the 1D fused path over a grid of cases, for comparing two builds of the library that must agree
BIT FOR BIT (a change of the load/store path, not of the arithmetic): run once per library and diff the outputs.
    HRWENO_B200_LIB=a.so python tools/lib_checksums.py > a.txt; HRWENO_B200_LIB=b.so python tools/lib_checksums.py > b.txt; diff a.txt b.txt
No torch: the host-pointer entry points only."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
A = pkg._abi
rng = np.random.default_rng(7)


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def widths(kind, nc):
    g = pkg.hrweno_grids.grid1()
    return g.linear(-5.0, 5.0, nc).width if kind == "lin" else g.geometric(-5.0, 5.0, 1.0 + 3.0 / nc, nc).width


SHAPES = [(1, 2500), (1, 100003), (37, 4096), (5, 1000), (3, 13)]
FLUX = [("burgers-godunov", A.FLUX_BURGERS, A.SCHEME_GODUNOV), ("burgers-lf", A.FLUX_BURGERS, A.SCHEME_LAX_FRIEDRICHS),
        ("linear-godunov", 1, A.SCHEME_GODUNOV)]
n_cases = 0
for mode_name, mode in (("strict", A.MODE_STRICT), ("fast", A.MODE_FAST)):
    for k in (1, 2, 3):
        for rows, nc in SHAPES:
            x = np.linspace(-5, 5, nc)
            u0 = (np.clip(1.0 - 0.25 * (x + 4.0), -0.5, 1.0)[None, :] * rng.uniform(0.5, 1.5, rows)[:, None] + 0.01 * rng.standard_normal((rows, nc))).reshape(-1)
            for wk in ("lin", "geo"):
                w = widths(wk, nc)
                for fname, model, scheme in FLUX:
                    if wk == "geo" and fname != "burgers-godunov":
                        continue
                    tag = f"{mode_name} k={k} {rows}x{nc} {wk} {fname}"
                    try:
                        d = pkg.fv.make_desc(nc, k=k, rows=rows, width=[w], mode=mode, flux_model=model, flux_scheme=scheme, alpha=1.3, flux_coef=(0.7, 1.0))
                        fv = pkg.fv.FV(d)
                        dt = 0.2 * float(w.min())
                        print(f"{tag} rhs {h(fv.rhs(0.0, u0))}")
                        for integ in ("rk1", "rk2", "rk3", "ms"):
                            ode = pkg.hrweno_tvdode.mstvd(fv, rows * nc) if integ == "ms" else pkg.hrweno_tvdode.rktvd(fv, rows * nc, int(integ[2]))
                            u = u0.copy()
                            t = ode.integrate(u, 0.0, 5.5 * dt, dt)
                            print(f"{tag} {integ} t={t!r} {h(u)}")
                            n_cases += 1
                            del ode
                        del fv
                    except Exception as e:  # the same refusal from both builds is agreement too
                        print(f"{tag} ERROR {type(e).__name__}: {str(e)[:80]}")
print(f"cases {n_cases}")
