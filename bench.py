#!/usr/bin/env python
"""bench.py -- WENO5 + Godunov + RK3-TVD cell-updates/s on B200 (BASELINE.json metric).

Workload (configs[2] of BASELINE.json, SURVEY 8d "cfg3"): 1D inviscid Burgers, k=3, eps=1e-6,
rktvd order 3, 2^28 cells per GPU on a linear grid over [-5,5], initial data = example1's clipped
ramp + 1e-3*N(0,1) (default_rng(12345)), dt = 0.1*dx.  One "step" = one RK3 time step = 3 fused
stage kernels over all cells; one cell-update = one cell-stage.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode strict|fast]

N > 1 is launched by torchrun (one rank per GPU); weak scaling: 2^28 cells per GPU, slab
decomposition along x with k halo cells per stage.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

RK3_BYTES_PER_CELL_STEP = 64.0  # 16 + 24 + 24
XMIN, XMAX = -5.0, 5.0


def make_ic(n, offset=0, n_global=None, seed=12345):
    """example1's ramp (example1:128-131) at the cell centres of grid1%linear plus a seeded perturbation"""
    n_global = n_global or n
    rx = (XMAX - XMIN) / n_global
    i = np.arange(offset, offset + n + 1, dtype=np.float64)
    edges = XMIN + rx * i
    x = (edges[:-1] + edges[1:]) / 2
    u = np.clip(1.0 + (-1.5 / 6.0) * (x + 4.0), -0.5, 1.0)
    rng = np.random.default_rng(seed + offset // max(n, 1))
    u += 1e-3 * rng.standard_normal(n)
    return u


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled DURING the timed region (NVML every 5 ms; the nvidia-smi query
    of B200_PROFILING.md once as a cross-check)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.power, self.bits, self._stop_evt = index, [], [], 0, threading.Event()
        self.sm_max, self.err = None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: fall back to nvidia-smi sampling
            self.nv, self.err = None, repr(e)

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            r = [c.strip() for c in out.split(",")]
            self.sm.append(float(r[0]))
            self.sm_max = self.sm_max or float(r[1])
            self.power.append(float(r[2]))
            for bit, val in zip((0x8, 0x40, 0x20, 0x4), r[3:7]):
                if val.lower().startswith("active"):
                    self.bits |= bit
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            if self.nv is not None:
                try:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    if len(self.sm) % 16 == 8:  # the power query is slow (tens of ms on some boxes): rarely, never first
                        self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception as e:
                    self.err = repr(e)
                self._stop_evt.wait(0.002)
            else:
                self._smi_once()
                self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_min_mhz": sm[0] if sm else None,
            "sm_max_mhz": self.sm_max,
            "power_w_max": max(self.power, default=None),
            "reasons": sorted(name for bit, name in self.REASONS.items() if self.bits & bit),
            "samples": len(sm),
            "source": "nvml 5 ms" if self.nv is not None else "nvidia-smi",
        }


def bind_to_gpu_numa_node(torch, index):
    """Host-side placement, as any MPI launcher would do it: run this rank on the cores of the NUMA node its GPU hangs off,
    so that the pinned host buffers (first touch) and the copy threads sit next to the PCIe root of that GPU.  Measured at
    8 ranks: without it the 8 x (2 GiB in + 2 GiB out) of an e2e call share ~130 GB/s.  Best effort; returns what it did."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"bdf": bdf, "numa_node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"bdf": bdf, "numa_node": node, "cpus": len(cpus), "bound": bool(cpus)}
    except Exception as e:  # no sysfs / no such attribute: leave the placement to the OS
        return {"bound": False, "why": repr(e)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference path (no Fortran compiler exists in this image,
    so oracle/_ref cannot be built; kind = "port") on all host threads, bounded sample per step."""
    if rank != 0:
        return
    pkg = graft.load_package()
    ref = graft.load_oracle()
    cores = ref.max_threads()
    ref.set_threads(cores)
    n = 1 << 24
    u = make_ic(n)
    dt = 0.1 * (XMAX - XMIN) / n
    fv = ref.FV(pkg.fv.make_desc(n, k=3, eps=1e-6, linear=(XMIN, XMAX)))
    ode = ref.rktvd(fv, 3)
    t = 0.0
    for _ in range(args.warmup):
        t = ode.integrate(u, t, 1e9, dt, itask=2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t = ode.integrate(u, t, 1e9, dt, itask=2)
    el = time.perf_counter() - t0
    value = n * 3 * args.steps / el
    sample = (f"2^24 cells x {args.steps} RK3 steps per run (1/16 of the 2^28-cell workload; a per-cell rate: u, ui, udot are 128 MiB "
              f"each, beyond the host's last-level cache, like the full size), OpenMP over cells on {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "WENO5+RK3 cell-updates/s", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3: 1D Burgers WENO5+Godunov+RK3, linear grid [-5,5], ramp+1e-3*N(0,1) IC, dt=0.1dx",
                   "cells_per_step": n, "note": "C restatement of the reference (oracle, bit-identical to the executed reference source), not a gfortran build"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


PARITY_CELLS = 1 << 20


def parity_head(n, world, total_steps, cells=None):
    """(cells compared, cells the oracle integrates): the first 2^20 cells of rank 0's slab plus the cells whose values can
    reach them in `total_steps` RK3 steps (k = 3 cells per stage) -- beyond that the cut is invisible"""
    cells = cells or PARITY_CELLS
    halo = 9 * total_steps + 32
    if world == 1 and n <= cells + halo:
        return n, n  # small run: the whole domain
    m = min(cells, n - halo)
    return (m, m + halo) if m > 0 else (0, 0)


def parity_check_leg(pkg, u0_head, u_gpu_head, n_global, dt, calls):
    """the oracle (reference operation order, oracle/hrweno_oracle.c) on the head of the SAME initial data for the SAME
    steps the timed state went through; checker only, outside every timed region"""
    ref = graft.load_oracle()
    ref.set_threads(ref.max_threads())
    mh, m = len(u0_head), len(u_gpu_head)
    rx = (XMAX - XMIN) / n_global
    edges = XMIN + rx * np.arange(mh + 1, dtype=np.float64)  # grids.f90:76-79
    ode = ref.rktvd(ref.FV(pkg.fv.make_desc(mh, k=3, eps=1e-6, width=[edges[1:] - edges[:-1]])), 3)
    u, t = u0_head.copy(), 0.0
    for k in calls:
        tt = t
        for _ in range(k - 1):
            tt = tt + dt
        t = ode.integrate(u, t, tt, dt)
    ref.set_threads(1)
    err = float(np.max(np.abs(u[:m] - u_gpu_head)) / np.max(np.abs(u[:m])))
    return {"cells": m, "oracle_cells": mh, "steps": int(sum(calls)), "max_normwise": err, "tolerance": 1e-12,
            "bit_identical": bool(np.array_equal(u[:m], u_gpu_head)), "ok": bool(err <= 1e-12),
            "what": "state after warm-up + timed steps on rank 0's first cells vs the CPU oracle on the same initial data"}


def cpu_baseline_leg(pkg):
    """the oracle on ONE host core (the shipped reference is serial), bounded sample"""
    ref = graft.load_oracle()
    ref.set_threads(1)
    n, steps = 1 << 22, 10
    u = make_ic(n)
    dt = 0.1 * (XMAX - XMIN) / n
    ode = ref.rktvd(ref.FV(pkg.fv.make_desc(n, k=3, eps=1e-6, linear=(XMIN, XMAX))), 3)
    t = ode.integrate(u, 0.0, 1e9, dt, itask=2)
    t0 = time.perf_counter()
    for _ in range(steps):
        t = ode.integrate(u, t, 1e9, dt, itask=2)
    el = time.perf_counter() - t0
    return {"value": n * 3 * steps / el, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "sample": f"2^22 cells x {steps} RK3 steps, 1 thread (C restatement of the reference, not gfortran); u, ui, udot are 32 MiB "
                      "each: a streaming working set at or beyond the host's last-level cache, as at the full 2^28 cells"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default=os.environ.get("HRWENO_BENCH_MODE", "fast"), choices=["strict", "fast"])
    ap.add_argument("--single-mode", action="store_true", help="skip the short measurement of the other arithmetic mode")
    ap.add_argument("--log2-cells", type=int, default=28)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip BASELINE.json configs[3] (cfg4) and configs[4] (cfg5)")
    ap.add_argument("--n2d", type=int, default=16384, help="cfg4 grid edge")
    ap.add_argument("--rows", type=int, default=65536, help="cfg5 rows")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE JSON line: everything libraries print while the job runs (NCCL's version banner, ...)
    # goes to stderr; the descriptor is restored just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    import torch
    import torch.distributed as dist

    pkg = graft.load_package()
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)  # before any pinned allocation: first touch decides where the pages live
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = pkg.lib()
    assert lib.hrweno_device_count() >= 1, "no CUDA device: there is no CPU fallback"

    n = 1 << args.log2_cells  # cells per GPU (weak scaling)
    n_global = n * world
    dt = 0.1 * (XMAX - XMIN) / n_global
    u_host = torch.from_numpy(make_ic(n, rank * n, n_global)).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream
    gather = pkg.slab.torch_all_gather(world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(*vals):
        if world == 1:
            return vals
        tmax = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in tmax)

    def measure(mode_name, steps, warmup, with_e2e):
        """one full measurement in one arithmetic mode: device-resident K steps, stage-kernel time, e2e"""
        mode = pkg._abi.MODE_STRICT if mode_name == "strict" else pkg._abi.MODE_FAST
        desc = pkg.fv.make_desc(n, k=3, eps=1e-6, linear=(XMIN, XMAX), mode=mode, rank=rank, nranks=world,
                                global_n=n_global, global_offset=rank * n)
        fv = pkg.fv.FV(desc)
        pkg.slab.connect(fv, rank, world, gather)
        ode = pkg.hrweno_tvdode.rktvd(fv, n, 3)
        u_dev = u_host.cuda(non_blocking=False)
        state = {"t": 0.0, "te": 0.0}
        # initial data of the cells the parity check re-integrates on the oracle (the e2e leg below advances u_host in place)
        ic_head = u_host.numpy()[:parity_head(n, world, warmup + steps)[1]].copy() if rank == 0 else None

        def run_steps(k):
            # k steps in ONE integrate call: tout = t after k-1 steps, the strict is_done test then takes exactly k
            tt = state["t"]
            for _ in range(k - 1):
                tt = tt + dt
            state["t"] = ode.integrate_dev(u_dev.data_ptr(), state["t"], tt, dt, 1, stream)

        run_steps(warmup)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ode.launches
        barrier()
        e0.record()
        run_steps(steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ode.launches - launches0
        # the state that was just timed (rank 0's first cells), for the parity check against the oracle after the run
        head = None
        if rank == 0:
            m, _ = parity_head(n, world, warmup + steps)
            head = u_dev[:m].cpu().numpy() if m > 0 else None
        # stage-kernel-only time: T(K steps) - T(1 step) removes the pack/unpack copies of the call
        ms1 = None
        for _ in range(3):  # the shortest of three single-step calls: a one-off hiccup here would skew every stage time
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            run_steps(1)
            e3.record()
            torch.cuda.synchronize()
            ms1 = e2.elapsed_time(e3) if ms1 is None else min(ms1, e2.elapsed_time(e3))
        clocks = sampler.stop()
        ms, ms1 = allmax(ms, ms1)
        out = {"ms": ms, "ms1": ms1, "launches": int(launches), "clocks": clocks, "head": head, "ic_head": ic_head, "calls": (warmup, steps)}
        if with_e2e:
            # host (pinned) u through the host-pointer C-ABI call, copies inside the timed region
            u_np = u_host.numpy()

            def e2e_call(k):
                tt = state["te"]
                for _ in range(k - 1):
                    tt = tt + dt
                state["te"] = ode.integrate(u_np, state["te"], tt, dt)

            e2e_call(1)
            barrier()
            w0 = time.perf_counter()
            e2e_call(steps)
            barrier()
            (out["e2e_s"],) = allmax(time.perf_counter() - w0)
            # the same API used one step per call (every step pays H2D + D2H of the whole state): the un-amortised figure
            ks = max(1, min(int(steps), 5))
            barrier()
            w1 = time.perf_counter()
            for _ in range(ks):
                e2e_call(1)
            barrier()
            (out["e2e_step_s"],) = allmax((time.perf_counter() - w1) / ks)
            # the ceiling the host side puts on e2e: pinned H2D and D2H of this rank's u, all ranks copying at once
            # (both directions at once as well, as the chunk pipeline does), CUDA events, slowest rank
            d_tmp = torch.empty(n, dtype=torch.float64, device="cuda")
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            barrier()
            with torch.cuda.stream(s_up):
                ev[0].record()
                d_tmp.copy_(u_host, non_blocking=True)
                ev[1].record()
            barrier()
            h2d_s = ev[0].elapsed_time(ev[1]) * 1e-3
            d_tmp2 = torch.empty(n, dtype=torch.float64, device="cuda")
            h_tmp = torch.empty(n, dtype=torch.float64).pin_memory()
            barrier()
            with torch.cuda.stream(s_up):
                ev[0].record()
                d_tmp.copy_(u_host, non_blocking=True)
                ev[1].record()
            with torch.cuda.stream(s_dn):
                ev[2].record()
                h_tmp.copy_(d_tmp2, non_blocking=True)
                ev[3].record()
            barrier()
            both_s = max(ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])) * 1e-3
            h2d_s, both_s = allmax(h2d_s, both_s)
            out["pcie"] = {"h2d_alone_gbs_per_gpu": n * 8 / h2d_s / 1e9, "h2d_and_d2h_at_once_gbs_per_gpu_each_way": n * 8 / both_s / 1e9,
                           "copy_floor_s": both_s, "note": "pinned host <-> device copies of one rank's u (8n B), every rank copying at the same time; "
                                                           "slowest rank; an e2e call cannot finish faster than copy_floor_s"}
            del d_tmp, d_tmp2, h_tmp
        del ode, fv, u_dev
        torch.cuda.empty_cache()
        return out

    def steps_to(t, dtx, k):
        tt = t
        for _ in range(k - 1):
            tt = tt + dtx
        return tt

    def timed_max(fn):
        """device time of fn() on the launching stream (CUDA events), barrier + synchronize on both sides, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0.record()
        fn()
        e1.record()
        barrier()
        clocks = sampler.stop()
        (sec,) = allmax(e0.elapsed_time(e1) * 1e-3)
        return sec, clocks

    def extra_cfg4(mode_name):
        """BASELINE.json configs[3]: 2D linear advection (f = v in both directions, Godunov, zero-flux walls) WENO5 + mstvd
        on an n x n grid, example2's box IC + 1e-3*N(0,1), dt = 0.125 dx, 20 multistep steps timed after the start-up.
        N GPUs: strong scaling, slabs along x2 with k halo rows per stage over NVLink peer memory."""
        mode = pkg._abi.MODE_STRICT if mode_name == "strict" else pkg._abi.MODE_FAST
        n2d = args.n2d
        off, n2 = pkg.slab.partition(n2d, world, rank)
        g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2d)
        fv = pkg.fv.FV(pkg.fv.make_desc((n2d, n2), flux_model=1, bc=1, width=[g.width, g.width[off:off + n2]], mode=mode, rank=rank,
                                        nranks=world, global_n=n2d, global_offset=off))
        pkg.slab.connect(fv, rank, world, gather)
        ode = pkg.hrweno_tvdode.mstvd(fv, n2d * n2)
        c1, c2 = g.center, g.center[off:off + n2]
        u = ((c2 >= 1.0) & (c2 <= 3.0))[:, None] & ((c1 >= 1.0) & (c1 <= 3.0))[None, :]
        u = u.astype(np.float64)
        u += 1e-3 * np.random.default_rng(12345 + rank).standard_normal((n2, n2d))
        # corner block the oracle re-integrates (rank 0): 256^2 compared, + the cells that can reach them (3 per stage)
        kw, ksteps = 6, 20
        nstages = 12 + (kw - 4) + ksteps
        mc, mh = 256, 256 + 3 * nstages + 16
        chk = rank == 0 and min(n2d, n2) > mh
        ic_blk = u[:mh, :mh].copy() if chk else None
        ud = torch.from_numpy(u.reshape(-1)).cuda()
        del u
        dt2 = 0.125 * 10.0 / n2d
        t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dt2, kw), dt2, 1, stream)  # start-up (4 RK3) + 2 multistep steps
        l0 = ode.launches
        sec, clocks = timed_max(lambda: ode.integrate_dev(ud.data_ptr(), t, steps_to(t, dt2, ksteps), dt2, 1, stream))
        launches = ode.launches - l0
        res = None
        if rank == 0:
            cells = n2d * n2d
            gbs = cells * 40.0 * ksteps / sec / 1e9 / world
            res = {"workload": f"cfg4: 2D linear advection WENO5+mstvd {n2d}x{n2d}, f=v, Godunov, zero-flux walls, box IC + 1e-3*N(0,1), dt=0.125dx; "
                               f"{ksteps} multistep steps after the 4-step RK3 start-up" + (f"; strong scaling, x2 slabs on {world} GPUs" if world > 1 else ""),
                   "mode": mode_name, "value": cells * ksteps / sec, "unit": "cell-steps/s", "ms_per_step": 1e3 * sec / ksteps, "gpu_launches": int(launches),
                   "scaling": "strong", "clocks": clocks,
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks()[0], "unit": "GB/s", "frac": gbs / peaks()[0], "traffic": None,
                                "kernel": "fv2d_stage_kernel<K=3, C_MS> (one launch per multistep step)", "per_gpu": True,
                                "algorithmic_bytes_per_launch": cells * 40.0 / world, "avg_launch_ms": 1e3 * sec / ksteps}}
            if chk:
                ref = graft.load_oracle()
                ref.set_threads(ref.max_threads())
                rode = ref.mstvd(ref.FV(pkg.fv.make_desc((mh, mh), flux_model=1, bc=1, width=[g.width[:mh], g.width[:mh]])))
                ur, tr = ic_blk.reshape(-1).copy(), 0.0
                tr = rode.integrate(ur, tr, steps_to(0.0, dt2, kw), dt2)
                tr = rode.integrate(ur, tr, steps_to(tr, dt2, ksteps), dt2)
                ref.set_threads(1)
                got = ud[: mc * n2d].cpu().numpy().reshape(mc, n2d)[:, :mc]
                want = ur.reshape(mh, mh)[:mc, :mc]
                err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
                res["parity_check"] = {"cells": mc * mc, "oracle_cells": mh * mh, "steps": kw + ksteps, "max_normwise": err, "tolerance": 1e-12,
                                       "bit_identical": bool(np.array_equal(got, want)), "ok": bool(err <= 1e-12),
                                       "what": "corner block of the timed state vs the CPU oracle on the same initial data"}
        del ode, fv, ud
        torch.cuda.empty_cache()
        return res

    def extra_cfg5(mode_name):
        """BASELINE.json configs[4]: 65536 independent rows x 4096 cells, Burgers + Godunov, k in {1,2,3} x rktvd order in
        {1,2,3}, 10 steps each; rows split over the GPUs, no halos."""
        mode = pkg._abi.MODE_STRICT if mode_name == "strict" else pkg._abi.MODE_FAST
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
        dt5 = 0.1 * 10.0 / nc
        bytes_step = {1: 16.0, 2: 40.0, 3: 64.0}
        kw, ksteps, prow = 3, 10, 8
        sweep, worst, launches = {}, 0.0, 0
        for k in (1, 2, 3):
            for order in (1, 2, 3):
                fv = pkg.fv.FV(pkg.fv.make_desc(nc, k=k, rows=rows, width=[g.width], mode=mode))
                ode = pkg.hrweno_tvdode.rktvd(fv, rows * nc, order)
                ud = ud0.clone()
                # the state is handed to the integrator once (hrweno_ode_attach) and fetched when output is due: no dense <->
                # padded copies inside the timed calls (they were 20 % of an RK1 call's traffic)
                ode.attach(ud.data_ptr(), stream)
                t = ode.integrate_attached(0.0, steps_to(0.0, dt5, kw), dt5, 1, stream)
                l0 = ode.launches
                sec, clocks = timed_max(lambda: ode.integrate_attached(t, steps_to(t, dt5, ksteps), dt5, 1, stream))
                launches += ode.launches - l0
                ode.fetch(ud.data_ptr(), stream)
                torch.cuda.synchronize()
                if rank == 0:
                    cells = rows_g * nc
                    gbs = cells * bytes_step[order] * ksteps / sec / 1e9 / world
                    ent = {"value": cells * order * ksteps / sec, "ms_per_step": 1e3 * sec / ksteps, "roofline_frac": gbs / peaks()[0], "achieved_gbs_per_gpu": gbs,
                           "sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"]}
                    # the first rows against the oracle (rows are independent)
                    ref = graft.load_oracle()
                    rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, k=k, rows=prow, width=[g.width])), order)
                    ur = u0[:prow].reshape(-1).copy()
                    tr = rode.integrate(ur, 0.0, steps_to(0.0, dt5, kw), dt5)
                    tr = rode.integrate(ur, tr, steps_to(tr, dt5, ksteps), dt5)
                    got = ud[: prow * nc].cpu().numpy()
                    ent["parity_normwise"] = float(np.max(np.abs(got - ur)) / np.max(np.abs(ur)))
                    worst = max(worst, ent["parity_normwise"])
                    sweep[f"k{k}_rktvd{order}"] = ent
                del ode, fv, ud
        del ud0
        torch.cuda.empty_cache()
        if rank != 0:
            return None
        head = sweep["k3_rktvd3"]
        return {"workload": f"cfg5: {rows_g} independent rows x {nc} cells, Burgers+Godunov, per-row ramp IC (rng 2024), dt=0.1dx, {ksteps} steps per (k, order); "
                            "rows split over the GPUs, no halos; state attached to the integrator (hrweno_ode_attach / _integrate_attached / _fetch)",
                "mode": mode_name, "unit": "cell-updates/s (cell-stages)", "value": head["value"], "value_is": "k=3, rktvd order 3", "scaling": "strong",
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": head["achieved_gbs_per_gpu"], "peak": peaks()[0], "unit": "GB/s", "frac": head["roofline_frac"],
                             "traffic": None, "kernel": "fv1d_stage_kernel<K=3> on 4096-cell rows (half-size tiles)", "per_gpu": True,
                             "algorithmic_bytes": "16 / 40 / 64 B per cell-step for rktvd 1 / 2 / 3"},
                "sweep": sweep,
                "parity_check": {"rows": prow, "steps": kw + ksteps, "max_normwise": worst, "tolerance": 1e-12, "ok": bool(worst <= 1e-12),
                                 "what": "first rows of every (k, order) run vs the CPU oracle"}}

    def extra_f32():
        """REAL32 build of the path (hrweno_kinds.F90:9-17): cfg3's problem at 2^26 cells in float32 on the general kernels
        (reference order, not the tuned TMA kernels) -- a coverage number, never the headline"""
        if rank != 0:
            return None
        # a float32 grid cannot resolve 2^26 cells on [-5, 5] (dx = 1.5e-7 is below the spacing of float32 near 5): the
        # REAL32 workload is the ensemble shape, 16384 rows x 4096 cells = 2^26 cells (256 MiB per array, beyond L2)
        rows, nc = 16384, 4096
        nf = rows * nc
        e = (np.float32(XMIN) + (np.float32(XMAX) - np.float32(XMIN)) / np.float32(nc) * np.arange(nc + 1, dtype=np.float32)).astype(np.float32)
        x = ((e[:-1] + e[1:]) / np.float32(2)).astype(np.float64)
        amp = np.random.default_rng(12345).uniform(0.5, 1.5, rows)
        u0 = (np.clip(1.0 + (-1.5 / 6.0) * (x + 4.0), -0.5, 1.0)[None, :] * amp[:, None]).astype(np.float32).reshape(-1)
        ode = pkg.real32.rktvd(pkg.real32.FV(pkg.real32.make_desc(nc, k=3, eps=1e-6, rows=rows, linear=(XMIN, XMAX))), 3)
        ud = torch.from_numpy(u0).cuda()
        dtf = 0.1 * (XMAX - XMIN) / nc
        kw, ks = 2, 10
        t = ode.integrate_dev(ud.data_ptr(), 0.0, steps_to(0.0, dtf, kw), dtf, 1, stream)
        tw = t
        l0 = ode.launches
        # rank 0 only: LOCAL timing (CUDA events, no barrier / all-reduce -- the other ranks are not here)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0.record()
        ode.integrate_dev(ud.data_ptr(), tw, float(np.float32(tw) + np.float32(ks - 0.5) * np.float32(dtf)), dtf, 1, stream)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        sec = e0.elapsed_time(e1) * 1e-3
        launches = ode.launches - l0
        ksteps = launches // 3
        gbs = nf * 32.0 * ksteps / sec / 1e9  # fp32 halves the bytes: 8 / 12 / 12 B per cell-stage
        res = {"workload": f"REAL32 (rk = real32): {rows} rows x {nc} cells 1D Burgers WENO5+Godunov+rktvd(3), general kernels (reference order, true "
                           "float divisions), not the tuned TMA kernels",
               "dtype": "f32", "value": nf * 3 * ksteps / sec, "unit": "cell-updates/s", "steps": int(ksteps), "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks()[0], "unit": "GB/s", "frac": gbs / peaks()[0], "traffic": None,
                            "kernel": "fvgen_stage_kernel<3, false, float> (8 / 12 / 12 algorithmic B per cell for the three RK3 stages + 4 B width)"},
               "clocks": clocks}
        # parity: the first rows of the state against the REAL32 oracle on the same initial data for the same steps
        from oracle import ref32
        prow = 8
        rode = ref32.rktvd(ref32.FV(pkg.real32.make_desc(nc, k=3, eps=1e-6, rows=prow, linear=(XMIN, XMAX))), 3)
        ur, tr = u0[: prow * nc].copy(), 0.0
        tr = rode.integrate(ur, tr, steps_to(0.0, dtf, kw), dtf)
        tr = rode.integrate(ur, tr, float(np.float32(tr) + np.float32(ks - 0.5) * np.float32(dtf)), dtf)
        got = ud[: prow * nc].cpu().numpy()
        res["parity_check"] = {"rows": prow, "steps": kw + int(ksteps), "bit_identical": bool(np.array_equal(got, ur)),
                               "max_normwise": float(np.max(np.abs(got - ur)) / np.max(np.abs(ur))),
                               "what": "first rows of the timed state vs the REAL32 build of the CPU oracle (bar: bit-identical)"}
        del ode, ud
        torch.cuda.empty_cache()
        return res

    K = args.steps
    main_m = measure(args.mode, K, args.warmup, True)
    other = "strict" if args.mode == "fast" else "fast"
    other_m = measure(other, max(2, min(K, 5)), 3, False) if not args.single_mode else None
    extra = None
    if not args.no_extra_configs:
        def guarded(fn, *a):
            """an extra configuration must never cost the headline line at N = 1: report its failure instead.  With several
            ranks an exception is re-raised (a rank that skipped ahead would leave the others inside a collective)"""
            if world > 1:
                return fn(*a)
            try:
                return fn(*a)
            except Exception as e:  # noqa: BLE001
                torch.cuda.empty_cache()
                return {"error": repr(e)}

        extra = {"cfg4": guarded(extra_cfg4, args.mode), "cfg5": guarded(extra_cfg5, args.mode), "real32": guarded(extra_f32)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()

    def derive(m, k):
        stage_ms = (m["ms"] - m["ms1"]) / (3 * (k - 1)) if k > 1 else m["ms"] / 3
        achieved = n * (RK3_BYTES_PER_CELL_STEP / 3) / (stage_ms * 1e-3) / 1e9
        return n_global * 3 * k / (m["ms"] * 1e-3), stage_ms, achieved

    value, stage_ms, achieved = derive(main_m, K)
    # DRAM bytes per launch from the committed ncu capture of the same kernels (profiles/traffic.json): the AVERAGE over
    # the three RK3 stage launches, the same averaging as algorithmic_bytes_per_launch next to it -- "from profile", not live
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(args.mode, {})
        if tj.get("log2_cells") == args.log2_cells:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = {"from": "profile (ncu --set full, not measured in this run)", "file": "profiles/traffic.json",
                           "per_stage_ratio_to_algorithmic": {k: v["ratio"] for k, v in tj.get("per_stage", {}).items()}}
    line = {
        "metric": "WENO5+RK3 cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": main_m["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "cfg3: 1D Burgers WENO5(k=3,eps=1e-6)+Godunov+rktvd(3), 2^%d cells per GPU, linear grid [-5,5], "
                        "ramp+1e-3*N(0,1) IC (rng 12345), dt=0.1dx" % args.log2_cells,
            "cells_per_gpu": n, "mode": args.mode,
            "parity": "strict = bit-identical to the oracle, which is bit-identical to the reference's source text executed by tools/f90exec (tests/test_reference_source_exec.py); fast = within 1e-12 normwise per output time (tests/test_gpu_parity.py)",
            "parallelism": "slab x%d (halo k=3 per stage over NVLink peer memory)" % world if world > 1 else "1 GPU",
            "l2": "state vectors are 2 GiB each, far larger than the 126 MB L2 (no flush needed)",
            "unit_note": "one cell-update = one cell-stage (rhs + stage combination); cell-steps/s = value/3",
        },
        "cell_steps_per_s": value / 3,
        "gpu_launches": main_m["launches"],
        "clocks": main_m["clocks"],
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "fv1d_stage_kernel<K=3> (average of the 3 RK3 stage instantiations; 16/24/24 algorithmic B per cell)",
            "algorithmic_bytes_per_launch": n * RK3_BYTES_PER_CELL_STEP / 3, "avg_launch_ms": stage_ms, "peak_source": peak_src,
            "note": "fp64 WENO5 is bound by fp64 instruction issue on B200, not by HBM (DESIGN.md section 5)",
        },
        "e2e": {
            "value": n_global * 3 * K / main_m["e2e_s"], "unit": "cell-updates/s",
            "h2d_bytes_per_step": n * 8 / K, "d2h_bytes_per_step": n * 8 / K,
            "note": "one hrweno_ode_integrate call with a pinned host u advancing K steps: H2D u (8n B), 3K fused stages, D2H u (8n B); "
                    "time-skewed chunk pipeline (slabs: on the slab extended by wide halos exchanged once per call)",
            "seconds_per_call": main_m["e2e_s"], "host_copy_ceiling": main_m.get("pcie"), "host_placement": numa,
            "one_call_per_step": {
                "value": n_global * 3 / main_m["e2e_step_s"], "unit": "cell-updates/s", "seconds_per_call": main_m["e2e_step_s"],
                "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
                "note": "hrweno_ode_integrate called once PER STEP with the pinned host u (itask-free: tout one step ahead): every step "
                        "pays the H2D and the D2H of the whole state, nothing is amortised; bound by host_copy_ceiling.copy_floor_s per call",
            },
        },
    }
    def parity_of(m):
        if m["head"] is None:
            return None
        return parity_check_leg(pkg, m["ic_head"], m["head"], n_global, dt, m["calls"])

    line["parity_check"] = parity_of(main_m)
    if other_m is not None:
        ko = max(2, min(K, 5))
        v2, sm2, ach2 = derive(other_m, ko)
        line["other_mode"] = {"mode": other, "value": v2, "steps": ko, "avg_launch_ms": sm2, "roofline_frac": ach2 / peak,
                              "parity_check": parity_of(other_m)}
    if extra is not None:
        line["configs"] = extra
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(pkg)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
