"""Multi-GPU parity worker: run under torchrun with N ranks (one per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py

Every rank integrates its slab on its GPU (halo cells over NVLink peer memory) and compares it bit
for bit with the same slab of the oracle's global single-domain result.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import ex1_ic, ex2_ic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg, ref = graft.load_package(), graft.load_oracle()
    gather = pkg.slab.torch_all_gather(world)
    fails = 0

    # ---- 1D: WENO k=1..3, rktvd 1..3 and mstvd, linear grid (analytic widths) and width arrays ----------
    for nglob, k, kind in [(1000, 3, "rk3"), (1003, 2, "rk2"), (997, 1, "rk1"), (4099, 3, "ms"), (50000, 3, "rk3w")]:
        off, n = pkg.slab.partition(nglob, world, rank)
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
        rng = np.random.default_rng(nglob)
        u0 = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nglob)
        common = dict(k=k, eps=1e-6)
        if kind.endswith("w"):
            desc = pkg.fv.make_desc(n, width=[g.width[off:off + n]], rank=rank, nranks=world, global_n=nglob, global_offset=off, **common)
        else:
            desc = pkg.fv.make_desc(n, linear=(-5.0, 5.0), rank=rank, nranks=world, global_n=nglob, global_offset=off, **common)
        fv = pkg.fv.FV(desc)
        pkg.slab.connect(fv, rank, world, gather)
        rfv = ref.FV(pkg.fv.make_desc(nglob, width=[g.width], **common))
        if kind == "ms":
            ode, rode = pkg.hrweno_tvdode.mstvd(fv, n), ref.mstvd(rfv)
        else:
            order = int(kind[2])
            ode, rode = pkg.hrweno_tvdode.rktvd(fv, n, order), ref.rktvd(rfv, order)
        u, ur, t, tr = np.ascontiguousarray(u0[off:off + n]), u0.copy(), 0.0, 0.0
        dt = 0.2 * 10.0 / nglob
        for tout in (0.0, 10 * dt, 25 * dt):
            t = ode.integrate(u, t, tout, dt)
            tr = rode.integrate(ur, tr, tout, dt)
            ok = t == tr and np.array_equal(u, ur[off:off + n])
            if not ok:
                fails += 1
                print(f"[rank {rank}] FAIL 1D n={nglob} k={k} {kind} tout={tout}: max|d|={np.max(np.abs(u - ur[off:off + n])):.3e}", flush=True)
        dist.barrier()

    # ---- 2D: slabs along x2, mstvd (example2 type) and rk3 -------------------------------------------------
    for (n1, n2g), kind in [((130, 96), "ms"), ((70, 61), "rk3")]:
        off, n2 = pkg.slab.partition(n2g, world, rank)
        g1 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1)
        g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2g)
        rng = np.random.default_rng(n1 + n2g)
        u0 = ex2_ic(g1.center, g2.center) + 1e-3 * rng.standard_normal((n2g, n1))
        common = dict(k=3, eps=1e-6, flux_model=1, bc=1)
        fv = pkg.fv.FV(pkg.fv.make_desc((n1, n2), width=[g1.width, g2.width[off:off + n2]], rank=rank, nranks=world,
                                        global_n=n2g, global_offset=off, **common))
        pkg.slab.connect(fv, rank, world, gather)
        rfv = ref.FV(pkg.fv.make_desc((n1, n2g), width=[g1.width, g2.width], **common))
        if kind == "ms":
            ode, rode = pkg.hrweno_tvdode.mstvd(fv, n1 * n2), ref.mstvd(rfv)
        else:
            ode, rode = pkg.hrweno_tvdode.rktvd(fv, n1 * n2, 3), ref.rktvd(rfv, 3)
        u = np.ascontiguousarray(u0[off:off + n2]).reshape(-1)
        ur = u0.reshape(-1).copy()
        t, tr = 0.0, 0.0
        for tout in (0.0, 0.1, 0.25):
            t = ode.integrate(u, t, tout, 1e-2)
            tr = rode.integrate(ur, tr, tout, 1e-2)
            want = ur.reshape(n2g, n1)[off:off + n2].reshape(-1)
            if not (t == tr and np.array_equal(u, want)):
                fails += 1
                print(f"[rank {rank}] FAIL 2D {n1}x{n2g} {kind} tout={tout}: max|d|={np.max(np.abs(u - want)):.3e}", flush=True)
        dist.barrier()

    # ---- FAST mode: the slab result must equal the single-domain FAST result bit for bit (per-cell arithmetic does
    # not depend on the decomposition) and stay within the north-star tolerance of the oracle ----------------------------
    FAST = pkg._abi.MODE_FAST
    for nglob, k, order in [(40000, 3, 3), (30011, 2, 2)]:
        off, n = pkg.slab.partition(nglob, world, rank)
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
        rng = np.random.default_rng(nglob)
        u0 = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nglob)
        common = dict(k=k, eps=1e-6, linear=(-5.0, 5.0), mode=FAST)
        fv = pkg.fv.FV(pkg.fv.make_desc(n, rank=rank, nranks=world, global_n=nglob, global_offset=off, **common))
        pkg.slab.connect(fv, rank, world, gather)
        ode = pkg.hrweno_tvdode.rktvd(fv, n, order)
        gode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(nglob, **common)), nglob, order)  # whole domain on this GPU
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nglob, width=[g.width], k=k, eps=1e-6)), order)
        u, ug, ur = np.ascontiguousarray(u0[off:off + n]), u0.copy(), u0.copy()
        dt = 0.2 * 10.0 / nglob
        t = ode.integrate(u, 0.0, 25 * dt, dt)
        tg = gode.integrate(ug, 0.0, 25 * dt, dt)
        rode.integrate(ur, 0.0, 25 * dt, dt)
        drift = np.max(np.abs(u - ur[off:off + n])) / np.max(np.abs(ur))
        if not (t == tg and np.array_equal(u, ug[off:off + n]) and drift < 1e-12):
            fails += 1
            print(f"[rank {rank}] FAIL 1D fast n={nglob} k={k}: slab vs single-domain max|d|={np.max(np.abs(u - ug[off:off + n])):.3e}, drift {drift:.3e}", flush=True)
        dist.barrier()

    # ---- adaptive Lax-Friedrichs alpha (extension): device reduction per rank + ONE NCCL max all-reduce ---------------
    nglob = 50001
    off, n = pkg.slab.partition(nglob, world, rank)
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(3).standard_normal(nglob)
    fv = pkg.fv.FV(pkg.fv.make_desc(n, k=3, flux_scheme=1, alpha=1.0, linear=(-5.0, 5.0), rank=rank, nranks=world, global_n=nglob,
                                    global_offset=off))
    pkg.slab.connect(fv, rank, world, gather)
    ud = torch.from_numpy(np.ascontiguousarray(u0[off:off + n])).cuda()
    alpha = pkg.slab.global_max_wavespeed(fv, ud, world)
    ode = pkg.hrweno_tvdode.rktvd(fv, n, 3)
    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nglob, k=3, flux_scheme=1, alpha=alpha, width=[g.width])), 3)
    u, ur = u0[off:off + n].copy(), u0.copy()  # copies: integrate works in place
    dt = 0.2 * 10.0 / nglob
    ode.integrate(u, 0.0, 10 * dt, dt)
    rode.integrate(ur, 0.0, 10 * dt, dt)
    if not (alpha == float(np.max(np.abs(u0))) and np.array_equal(u, ur[off:off + n])):
        fails += 1
        print(f"[rank {rank}] FAIL adaptive alpha: {alpha!r} vs {float(np.max(np.abs(u0)))!r}", flush=True)
    dist.barrier()

    # ---- host-pointer integrate on slabs: per-slab chunk pipeline on the slab extended by the wide halos (exchanged once
    # per call through the IPC-mapped mailboxes), no per-stage traffic; forced at a small size by a small chunk --------
    os.environ["HRWENO_PIPE_CHUNK_TILES"] = "29"
    for mode, order in [(pkg._abi.MODE_STRICT, 3), (pkg._abi.MODE_STRICT, 1), (FAST, 3)]:
        nglob = world * 140000 + 777
        off, n = pkg.slab.partition(nglob, world, rank)
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
        u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(world + order).standard_normal(nglob)
        fv = pkg.fv.FV(pkg.fv.make_desc(n, k=3, linear=(-5.0, 5.0), mode=mode, rank=rank, nranks=world, global_n=nglob, global_offset=off))
        pkg.slab.connect(fv, rank, world, gather)
        ode = pkg.hrweno_tvdode.rktvd(fv, n, order)
        ref.set_threads(max(1, ref.max_threads() // world))
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nglob, k=3, width=[g.width])), order)
        u, ur, t, tr = np.ascontiguousarray(u0[off:off + n]), u0.copy(), 0.0, 0.0
        dt = 0.2 * 10.0 / nglob
        l0 = ode.launches
        for nsteps in (6, 1, 11):
            tt, ttr = t, tr
            for _ in range(nsteps - 1):
                tt, ttr = tt + dt, ttr + dt
            t = ode.integrate(u, t, tt, dt)
            tr = rode.integrate(ur, tr, ttr, dt)
            if mode == FAST:
                ok = t == tr and np.max(np.abs(u - ur[off:off + n])) / np.max(np.abs(ur)) < 1e-12
            else:
                ok = t == tr and np.array_equal(u, ur[off:off + n])
            if not ok:
                fails += 1
                print(f"[rank {rank}] FAIL slab host pipeline mode={mode} order={order} nsteps={nsteps}: max|d|={np.max(np.abs(u - ur[off:off + n])):.3e}", flush=True)
        if ode.launches - l0 <= 3 * 18 * order:
            fails += 1
            print(f"[rank {rank}] FAIL slab host pipeline not taken: {ode.launches - l0} launches", flush=True)
        ref.set_threads(1)
        dist.barrier()
    del os.environ["HRWENO_PIPE_CHUNK_TILES"]

    tot = torch.tensor([fails], device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        print(f"mgpu parity: world={world} failures={int(tot[0])}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(tot[0]) else 0)


if __name__ == "__main__":
    main()
