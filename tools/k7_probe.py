#!/usr/bin/env python
"""A few launches of the general stage kernel K7 (csrc/fvgen.cu) for an ncu capture.

usage (under gpurun): ncu --set full --clock-control none --import-source on -k regex:fvgen -s 3 -c 1 \\
                          -o gpurun_out/k7_1d python tools/k7_probe.py --dim 1
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex1_ic, ex2_ic  # noqa: E402

import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=1)
ap.add_argument("--log2n", type=int, default=22)
ap.add_argument("--n2d", type=int, default=4096)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
pkg = graft.load_package()
stream = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(12345)
if args.dim == 1:
    nc = 1 << args.log2n
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.0 + 1e-7, nc)
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=3, width=[g.width]))
    fv.set_xedges(0, g.edges)
    ode = pkg.hrweno_tvdode.rktvd(fv, nc, 3)
    ud = torch.from_numpy(ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)).cuda()
    dt = 0.1 * 10.0 / nc
else:
    n = args.n2d
    g1 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.0005, n)
    g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.0003, n)
    fv = pkg.fv.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g1.width, g2.width]))
    fv.set_xedges(0, g1.edges)
    fv.set_xedges(1, g2.edges)
    fv.set_flux_coef(0, g1.edges**2, None)
    fv.set_flux_coef(1, g2.edges, g1.center)
    ode = pkg.hrweno_tvdode.rktvd(fv, n * n, 3)
    ud = torch.from_numpy((ex2_ic(g1.center, g2.center) + 1e-3 * rng.standard_normal((n, n))).reshape(-1)).cuda()
    dt = 1e-6
t = 0.0
for _ in range(args.steps):
    t = ode.integrate_dev(ud.data_ptr(), t, t, dt, 2, stream)  # itask = 2: one step per call
torch.cuda.synchronize()
print("k7 probe done", t)
