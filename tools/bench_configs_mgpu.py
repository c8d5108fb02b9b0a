#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] on N GPUs of one box (one process per GPU, torchrun):

  cfg4: 2D linear advection WENO5+mstvd on a 16384 x 16384 grid, slabs along x2, k halo ROWS per stage over NVLink
        peer memory (csrc/halo.cu); strong scaling (the grid is fixed, each rank owns 16384/N rows)
  cfg5: 65536 independent rows x 4096 cells, WENO5 + rktvd3, rows split over the ranks, no communication

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs_mgpu.py [--mode fast]

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex2_ic  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fast")
ap.add_argument("--n2d", type=int, default=16384)
ap.add_argument("--rows", type=int, default=65536)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = graft.load_package()
MODE = pkg._abi.MODE_STRICT if args.mode == "strict" else pkg._abi.MODE_FAST
PEAK = 6538.9
stream = torch.cuda.current_stream().cuda_stream
gather = pkg.slab.torch_all_gather(world)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def steps_to(t, dt, k):
    tt = t
    for _ in range(k - 1):
        tt = tt + dt
    return tt


# ---- cfg4 -------------------------------------------------------------------------------------------------
n = args.n2d
off, n2 = pkg.slab.partition(n, world, rank)
g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
fv = pkg.fv.FV(pkg.fv.make_desc((n, n2), flux_model=1, bc=1, width=[g.width, g.width[off:off + n2]], mode=MODE, rank=rank,
                                nranks=world, global_n=n, global_offset=off))
pkg.slab.connect(fv, rank, world, gather)
ode = pkg.hrweno_tvdode.mstvd(fv, n * n2)
u = ex2_ic(g.center, g.center[off:off + n2]) + 1e-3 * np.random.default_rng(12345 + rank).standard_normal((n2, n))
ud = torch.from_numpy(u.reshape(-1)).cuda()
dt = 0.125 * 10.0 / n
t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt, 6), dt, 1, stream)
K = 20
el = timed(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt, K), dt, 1, stream))
if rank == 0:
    cells = n * n
    gbs = cells * 40.0 * K / el / 1e9
    print(f"cfg4 2D {n}x{n} WENO5+mstvd ({args.mode}) on {world} GPU(s): {K} steps in {el*1e3:.1f} ms -> {cells*K/el:.3e} cell-steps/s, "
          f"{gbs/world:.0f} GB/s algorithmic per GPU = {gbs/world/PEAK:.3f} of measured HBM peak", flush=True)
del ode, fv, ud

# ---- cfg5 -------------------------------------------------------------------------------------------------
rows_g, nc = args.rows, 4096
roff, rows = pkg.slab.partition(rows_g, world, rank)
g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
rng = np.random.default_rng(2024)
va, vb = rng.uniform(0.5, 1.5, rows_g), rng.uniform(-1.0, 0.0, rows_g)
xa, xb = rng.uniform(-4.5, -3.0, rows_g), rng.uniform(1.0, 3.0, rows_g)
sl = slice(roff, roff + rows)
x = g.center[None, :]
u0 = np.clip(va[sl, None] + ((vb - va) / (xb - xa))[sl, None] * (x - xa[sl, None]), np.minimum(va, vb)[sl, None], np.maximum(va, vb)[sl, None])
ud0 = torch.from_numpy(u0.reshape(-1)).cuda()
dt = 0.1 * 10.0 / nc
for k in (1, 2, 3):
    fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=k, rows=rows, width=[g.width], mode=MODE))
    ode = pkg.hrweno_tvdode.rktvd(fv, rows * nc, 3)
    ud = ud0.clone()
    t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt, 3), dt, 1, stream)
    K = 10
    el = timed(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt, K), dt, 1, stream))
    if rank == 0:
        cells = rows_g * nc
        gbs = cells * 64.0 * K / el / 1e9
        print(f"cfg5 ensemble {rows_g}x{nc} k={k} rktvd3 ({args.mode}) on {world} GPU(s): {cells*3*K/el:.3e} cell-stages/s, "
              f"{gbs/world:.0f} GB/s algorithmic per GPU = {gbs/world/PEAK:.3f} of peak", flush=True)
    del ode, fv, ud
if world > 1:
    dist.destroy_process_group()
