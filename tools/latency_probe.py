#!/usr/bin/env python
"""Where the time of the small shipped configurations goes: device-resident integrate vs host-pointer integrate."""
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

pkg = graft.load_package()
stream = torch.cuda.current_stream().cuda_stream
for name in ("cfg1", "cfg2"):
    if name == "cfg1":
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
        mk = lambda: pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(100, width=[g.width])), 100, 3)
        u0, dt, tend, nsteps = ex1_ic(g.center), 1e-2, 12.0, 1200
    else:
        g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, 250)
        mk = lambda: pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc((250, 250), flux_model=1, bc=1, width=[g.width, g.width])), 62500)
        u0, dt, tend, nsteps = ex2_ic(g.center, g.center).reshape(-1), 5e-3, 5.0, 1000
    for kind in ("host pageable", "host pinned", "device"):
        ode = mk()
        if kind == "device":
            ud = torch.from_numpy(u0.copy()).cuda()
            step = lambda t, tout: ode.integrate_dev(ud.data_ptr(), t, tout, dt, 1, stream)
        else:
            uh = torch.from_numpy(u0.copy())
            if kind == "host pinned":
                uh = uh.pin_memory()
            un = uh.numpy()
            step = lambda t, tout: ode.integrate(un, t, tout, dt)
        t = step(0.0, 0.0)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for ii in range(1, 101):
            t = step(t, tend * ii / 100)
        torch.cuda.synchronize()
        el = time.perf_counter() - w0
        print(f"{name} {kind:14s}: {el*1e3:7.1f} ms for 100 calls / {nsteps} steps -> {el/100*1e6:6.0f} us per call, {el/nsteps*1e6:5.1f} us per step")
