"""One process, N GPUs through the C ABI alone (hrweno_mgpu_*, csrc/mgpu.cu; SURVEY 8b "ngpus", 8e): the library cuts
the global problem into slabs, wires the halo mailboxes over peer memory and reduces the Lax-Friedrichs alpha itself.
Strict mode must equal the oracle's single-domain run bit for bit for every GPU count the box offers (1 included, so
the entry points are exercised on a single-GPU box too); examples/example4_multi_gpu.cpp drives cfg3 with no Python."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, ex1_ic, ex2_ic

pytestmark = pytest.mark.gpu


def _counts(gpu_lib):
    n = gpu_lib.hrweno_device_count()
    return [w for w in (1, 2, 4, 8) if w <= n]


def _steps_to(t, dt, k):
    for _ in range(k - 1):
        t = t + dt
    return t


@pytest.mark.parametrize("kind,k,nglob", [("rk3", 3, 40000), ("rk2", 2, 10007), ("rk1", 1, 9973), ("ms", 3, 25013)])
def test_mgpu_1d_slabs_bitwise(gpu_lib, pkg, ref, kind, k, nglob):
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(nglob).standard_normal(nglob)
    dt = 0.2 * 10.0 / nglob
    rfv = ref.FV(pkg.fv.make_desc(nglob, k=k, width=[g.width]))
    rode = ref.mstvd(rfv) if kind == "ms" else ref.rktvd(rfv, int(kind[2]))
    ur, tr = u0.copy(), 0.0
    outs = []
    for tout in (0.0, 10 * dt, 25 * dt):
        tr = rode.integrate(ur, tr, tout, dt)
        outs.append((tr, ur.copy()))
    for world in _counts(gpu_lib):
        for linear in (True, False):
            kw = dict(linear=(-5.0, 5.0)) if linear else dict(width=[g.width])
            m = pkg.mgpu.MultiGPU(pkg.fv.make_desc(nglob, k=k, **kw), world)
            assert m.ngpus == world
            assert sum(m.slab(r)[2] for r in range(world)) == nglob
            m.mstvd() if kind == "ms" else m.rktvd(int(kind[2]))
            u, t = u0.copy(), 0.0
            for tout, (tr_, ur_) in zip((0.0, 10 * dt, 25 * dt), outs):
                t = m.integrate(u, t, tout, dt)
                assert t == tr_ and np.array_equal(u, ur_), f"world={world} linear={linear} {kind} k={k}"
            assert m.fevals == rode.fevals
            del m


def test_mgpu_resident_state_and_alpha_reduction(gpu_lib, pkg, ref):
    """device-resident state between calls; alpha = max|u| over all slabs reduced inside the library, then
    Lax-Friedrichs with that alpha equals the oracle given the same alpha"""
    nglob = 30011
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(7).standard_normal(nglob)
    u0[nglob - 5] = -1.75  # the global maximum of |u| sits in the last slab
    dt = 0.1 * 10.0 / nglob
    for world in _counts(gpu_lib):
        m = pkg.mgpu.MultiGPU(pkg.fv.make_desc(nglob, k=3, flux_scheme=1, alpha=0.0, width=[g.width]), world).rktvd(3)
        m.upload(u0)
        alpha = m.max_wavespeed(install=True)
        assert alpha == np.max(np.abs(u0)) == 1.75
        t = m.integrate_resident(0.0, _steps_to(0.0, dt, 7), dt)
        t = m.integrate_resident(t, _steps_to(t, dt, 5), dt)
        u = m.download(np.empty(nglob))
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nglob, k=3, flux_scheme=1, alpha=alpha, width=[g.width])), 3)
        ur = u0.copy()
        tr = rode.integrate(ur, 0.0, _steps_to(0.0, dt, 7), dt)
        tr = rode.integrate(ur, tr, _steps_to(tr, dt, 5), dt)
        assert t == tr and np.array_equal(u, ur), f"world={world}"
        del m


def test_mgpu_2d_slabs_and_rows_bitwise(gpu_lib, pkg, ref):
    n1, n2 = 300, 260
    g1, g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    u0 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(3).standard_normal((n2, n1))).reshape(-1)
    dt = 5e-3
    kw = dict(flux_model=1, bc=1, width=[g1.width, g2.width])
    rode = ref.mstvd(ref.FV(pkg.fv.make_desc((n1, n2), **kw)))
    ur = u0.copy()
    tr = rode.integrate(ur, 0.0, _steps_to(0.0, dt, 9), dt)
    rows, nc = 37, 4096
    gr = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    r0 = (ex1_ic(gr.center)[None, :] * np.linspace(0.5, 1.5, rows)[:, None]).reshape(-1)
    rrode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, k=3, rows=rows, width=[gr.width])), 3)
    rr = r0.copy()
    rrode.integrate(rr, 0.0, _steps_to(0.0, 2e-4, 6), 2e-4)
    for world in _counts(gpu_lib):
        m = pkg.mgpu.MultiGPU(pkg.fv.make_desc((n1, n2), **kw), world).mstvd()
        u = u0.copy()
        t = m.integrate(u, 0.0, _steps_to(0.0, dt, 9), dt)
        assert t == tr and np.array_equal(u, ur), f"2D world={world}"
        del m
        # independent rows dealt out over the GPUs (cfg5 shape): no halos
        m = pkg.mgpu.MultiGPU(pkg.fv.make_desc(nc, k=3, rows=rows, width=[gr.width]), world).rktvd(3)
        u = r0.copy()
        m.integrate(u, 0.0, _steps_to(0.0, 2e-4, 6), 2e-4)
        assert np.array_equal(u, rr), f"rows world={world}"
        del m


def test_example4_cpp_multi_gpu(gpu_lib, pkg, ref, tmp_path):
    """the C++ host program (no Python on the path) over every GPU count, strict mode, against the oracle"""
    exe = os.path.join(ROOT, "examples", "example4_multi_gpu")
    assert os.path.exists(exe), "examples not built (python __graft_entry__.py)"
    log2n, steps = 18, 12
    nc = 1 << log2n
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    s, pert = 12345, np.empty(nc)
    for i in range(nc):  # the program's own deterministic perturbation
        s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        pert[i] = (s >> 11) / 9007199254740992.0 - 0.5
    u0 = ex1_ic(g.center) + 1e-3 * pert
    dt = 0.1 * 10.0 / nc
    ref.set_threads(min(16, ref.max_threads()))
    try:
        rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nc, k=3, width=[g.width])), 3)
        ur = u0.copy()
        tr = rode.integrate(ur, 0.0, _steps_to(0.0, dt, steps), dt)
    finally:
        ref.set_threads(1)
    for world in _counts(gpu_lib):
        out = subprocess.run([exe, str(tmp_path), str(log2n), str(world), str(steps), "0"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert f"on {world} GPU(s)" in out.stdout and f"fevals = {3 * steps}" in out.stdout
        u = np.fromfile(tmp_path / "u_final.bin")
        assert np.array_equal(u, ur), f"world={world}: max|d| = {np.max(np.abs(u - ur)):.3e}"


@pytest.mark.parametrize("order", [1, 3])
def test_mgpu_host_pipeline_on_slabs_bitwise(gpu_lib, pkg, ref, order, monkeypatch):
    """host-pointer integrate on slabs: every slab runs its own time-skewed chunk pipeline on the slab extended by the
    wide halos exchanged once per call (ode.cu: pipeline_operator; halo.cu: fv_exchange_wide) -- no per-stage traffic.
    Must equal the oracle's single-domain run bit for bit, call after call (the second call reuses the wide slots)."""
    monkeypatch.setenv("HRWENO_PIPE_CHUNK_TILES", "29")
    for world in _counts(gpu_lib):
        nglob = world * 140000 + 777
        g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
        u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(world).standard_normal(nglob)
        dt = 0.2 * 10.0 / nglob
        ref.set_threads(min(16, ref.max_threads()))
        try:
            rode = ref.rktvd(ref.FV(pkg.fv.make_desc(nglob, k=3, width=[g.width])), order)
            m = pkg.mgpu.MultiGPU(pkg.fv.make_desc(nglob, k=3, linear=(-5.0, 5.0)), world).rktvd(order)
            u, ur, t, tr = u0.copy(), u0.copy(), 0.0, 0.0
            l0 = m.launches
            for nsteps in (6, 1, 11):
                t = m.integrate(u, t, _steps_to(t, dt, nsteps), dt)
                tr = rode.integrate(ur, tr, _steps_to(tr, dt, nsteps), dt)
                assert t == tr and np.array_equal(u, ur), f"world={world} order={order} after {nsteps} steps: {np.max(np.abs(u - ur)):.3e}"
            # more launches than stages: the chunked pipeline ran (18 steps x order stages x world slabs otherwise)
            assert m.launches - l0 > 3 * 18 * order * world
        finally:
            ref.set_threads(1)
        del m


def _g_of_t(t):
    return 1.0 + 0.5 * np.sin(3.0 * t) + 0.25 * t


def test_mgpu_general_operators_on_slabs_bitwise(gpu_lib, pkg, ref):
    """non-uniform grids (per-cell tables from the GLOBAL edges), x- and t-dependent fluxes on slabs: the general kernels read
    the neighbour's cells from the exchanged ghost cells and reconstruct them with the neighbour's tables"""
    # 1D: geometric grid, face coefficient, time factor, rktvd3 and mstvd
    nglob = 20011
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.0002, nglob)
    u0 = ex1_ic(g.center) + 1e-3 * np.random.default_rng(4).standard_normal(nglob)
    fc = 1.0 + 0.1 * np.cos(g.edges)
    for kind in ("rk3", "ms"):
        rfv = ref.FV(pkg.fv.make_desc(nglob, k=3, width=[g.width]))
        rfv.set_xedges(0, g.edges)
        rfv.set_flux_coef(0, fc, None)
        rfv.set_flux_time_fn(_g_of_t)
        rode = ref.mstvd(rfv) if kind == "ms" else ref.rktvd(rfv, 3)
        ur, dt = u0.copy(), 2e-5
        tr = rode.integrate(ur, 0.0, _steps_to(0.0, dt, 9), dt)
        for world in _counts(gpu_lib):
            m = pkg.mgpu.MultiGPU(pkg.fv.make_desc(nglob, k=3, width=[g.width]), world)
            m.set_xedges(0, g.edges)
            m.set_flux_coef(0, fc, None)
            m.set_flux_time_fn(_g_of_t)
            m.mstvd() if kind == "ms" else m.rktvd(3)
            u = u0.copy()
            t = m.integrate(u, 0.0, _steps_to(0.0, dt, 9), dt)
            assert t == tr and np.array_equal(u, ur), f"1D general {kind} world={world}: {np.max(np.abs(u - ur)):.3e}"
            del m
    # 2D: example2 on geometric x geometric grids with the growth terms of example2:140,153 and a time factor
    n1, n2 = 150, 131
    g1, g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.01, n1), pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.007, n2)
    v0 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(6).standard_normal((n2, n1))).reshape(-1)
    kw = dict(flux_model=1, bc=1, width=[g1.width, g2.width])

    def setup(f):
        f.set_xedges(0, g1.edges)
        f.set_xedges(1, g2.edges)
        f.set_flux_coef(0, g1.edges**2, 1.0 + 0.01 * g2.center)
        f.set_flux_coef(1, g2.edges, g1.center)
        f.set_flux_time_fn(_g_of_t)

    for kind in ("ms", "rk3"):
        rfv = ref.FV(pkg.fv.make_desc((n1, n2), **kw))
        setup(rfv)
        rode = ref.mstvd(rfv) if kind == "ms" else ref.rktvd(rfv, 3)
        vr, dt = v0.copy(), 1e-4
        tr = rode.integrate(vr, 0.0, _steps_to(0.0, dt, 8), dt)
        for world in _counts(gpu_lib):
            m = pkg.mgpu.MultiGPU(pkg.fv.make_desc((n1, n2), **kw), world)
            setup(m)
            m.mstvd() if kind == "ms" else m.rktvd(3)
            v = v0.copy()
            t = m.integrate(v, 0.0, _steps_to(0.0, dt, 8), dt)
            assert t == tr and np.array_equal(v, vr), f"2D general {kind} world={world}: {np.max(np.abs(v - vr)):.3e}"
            del m
