#!/usr/bin/env python
"""A/B helper: the batched ensemble (cfg5: rows x 4096 cells, Burgers + Godunov, fast mode) with the state attached to
the integrator, for the (k, order) pairs given -- prints cell-stages/s and the fraction of the measured HBM peak.
    HRWENO_B200_LIB=path/to/lib.so python tools/ab_k1.py [--rows 65536] [--pairs 1:1,1:2,1:3,3:3] [--reps 2]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=65536)
ap.add_argument("--nc", type=int, default=4096)
ap.add_argument("--pairs", default="1:1,1:2,1:3,3:3")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--mode", default="fast")
args = ap.parse_args()
pkg = graft.load_package()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6552.3
MODE = pkg._abi.MODE_FAST if args.mode == "fast" else pkg._abi.MODE_STRICT
stream = torch.cuda.current_stream().cuda_stream
rows, nc = args.rows, args.nc
g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
rng = np.random.default_rng(2024)
va, vb = rng.uniform(0.5, 1.5, rows), rng.uniform(-1.0, 0.0, rows)
xa, xb = rng.uniform(-4.5, -3.0, rows), rng.uniform(1.0, 3.0, rows)
if rows == 1:  # cfg3's shape: one long row, built on the device
    xd = torch.linspace(-5.0, 5.0, nc, dtype=torch.float64, device="cuda")
    ud0 = torch.clamp(1.0 - 0.25 * (xd + 4.0), -0.5, 1.0)
    del xd
else:
    x = g.center[None, :]
    u0 = np.clip(va[:, None] + ((vb - va) / (xb - xa))[:, None] * (x - xa[:, None]), np.minimum(va, vb)[:, None], np.maximum(va, vb)[:, None])
    ud0 = torch.from_numpy(u0.reshape(-1)).cuda()
dt = 0.1 * 10.0 / nc
bytes_step = {1: 16.0, 2: 40.0, 3: 64.0}


def steps_to(t, dt, k):
    return t + (k - 0.5) * dt


for pair in args.pairs.split(","):
    k, order = (int(v) for v in pair.split(":"))
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=k, rows=rows, width=[g.width], mode=MODE))
    ode = pkg.hrweno_tvdode.rktvd(fv, rows * nc, order)
    ud = ud0.clone()
    ode.attach(ud.data_ptr(), stream)
    t = ode.integrate_attached(0.0, steps_to(0.0, dt, 3), dt, 1, stream)
    res = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        t = ode.integrate_attached(t, steps_to(t, dt, 10), dt, 1, stream)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        res.append(rows * nc * bytes_step[order] * 10 / sec / 1e9 / PEAK)
    print(f"{os.path.basename(os.environ.get('HRWENO_B200_LIB', 'default')):22s} {rows}x{nc} k={k} rktvd{order} ({args.mode}): " + " ".join(f"{r:.3f}" for r in res)
          + f"  ({rows * nc * order * 10 / sec:.3e} cell-stages/s)", flush=True)
    del ode, fv, ud
