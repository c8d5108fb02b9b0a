#!/usr/bin/env python
"""The REAL32 path on the shape bench.py's `real32` extra uses (rows x 4096 cells, WENO5 + Godunov + rktvd3), a few steps:
a target for `ncu -k regex:fvgen_stage_kernel` and a quick throughput print.
    python tools/f32_probe.py [--rows 16384] [--steps 4]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=16384)
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()
pkg = graft.load_package()
stream = torch.cuda.current_stream().cuda_stream
rows, nc = args.rows, 4096
F = np.float32
e = (F(-5) + (F(5) - F(-5)) / F(nc) * np.arange(nc + 1, dtype=F)).astype(F)
x = ((e[:-1] + e[1:]) / F(2)).astype(np.float64)
amp = np.random.default_rng(12345).uniform(0.5, 1.5, rows)
u0 = (np.clip(1.0 + (-1.5 / 6.0) * (x + 4.0), -0.5, 1.0)[None, :] * amp[:, None]).astype(F).reshape(-1)
ode = pkg.real32.rktvd(pkg.real32.FV(pkg.real32.make_desc(nc, k=3, eps=1e-6, rows=rows, linear=(-5.0, 5.0))), 3)
ud = torch.from_numpy(u0).cuda()
dt = 0.1 * 10.0 / nc
t = ode.integrate_dev(ud.data_ptr(), 0.0, float(F(1.5) * F(dt)), dt, 1, stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
l0 = ode.launches
e0.record()
ode.integrate_dev(ud.data_ptr(), t, float(F(t) + F(args.steps - 0.5) * F(dt)), dt, 1, stream)
e1.record()
torch.cuda.synchronize()
sec, st = e0.elapsed_time(e1) * 1e-3, (ode.launches - l0)
print(f"real32 {rows}x{nc} k=3 rktvd3: {st} stage launches in {sec * 1e3:.2f} ms -> {rows * nc * st / sec:.3e} cell-stages/s")
