"""Committed fixtures (tests/golden/*.npz, oracle-generated -- see make_golden.py): the oracle on CPU and the CUDA
strict path on GPU must both reproduce them bit for bit."""
import os

import numpy as np
import pytest

from conftest import ROOT, ex1_ic, ex2_ic, pulse

G = os.path.join(ROOT, "tests", "golden")


def _check_reconstruct(fn):
    z = np.load(os.path.join(G, "reconstruct.npz"))
    for k in (1, 2, 3):
        for name, v in (("pulse", pulse(30)), ("rand", z["rand_v"])):
            vl, vr = fn(v, k)
            assert np.array_equal(vl, z[f"{name}_k{k}_vl"]) and np.array_equal(vr, z[f"{name}_k{k}_vr"])


def _check_example1(make_ode, grid):
    z = np.load(os.path.join(G, "example1.npz"))
    ode = make_ode()
    u, t = ex1_ic(grid.center), 0.0
    for ii in range(101):
        t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
        assert t == z["times"][ii]
        if ii in (0, 50, 100):
            assert np.array_equal(u, z[f"u_{ii}"])


def _check_example2(make_ode, grid, n=250):
    z = np.load(os.path.join(G, "example2.npz"))
    ode = make_ode()
    u, t = ex2_ic(grid.center, grid.center).reshape(-1), 0.0
    for ii in range(101):
        t = ode.integrate(u, t, 5.0 * ii / 100, 5e-3)
        assert t == z["times"][ii]
        U = u.reshape(n, n)
        assert U.min() == z["umin"][ii] and U.max() == z["umax"][ii]
    assert np.array_equal(U[20:30], z["rows_20_29"]) and np.array_equal(U[60:70], z["rows_60_69"])
    assert np.array_equal(U.sum(axis=1), z["row_sums"])


def test_oracle_reconstruct_golden(ref):
    _check_reconstruct(lambda v, k: ref.reconstruct(v, k, 1e-6))


def test_oracle_example1_golden(ref, pkg):
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    _check_example1(lambda: ref.rktvd(ref.FV(pkg.fv.make_desc(100, width=[g.width])), 3), g)


@pytest.mark.slow
def test_oracle_example2_golden(ref, pkg):
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, 250)
    ref.set_threads(min(8, ref.max_threads()))
    try:
        _check_example2(lambda: ref.mstvd(ref.FV(pkg.fv.make_desc((250, 250), flux_model=1, bc=1, width=[g.width, g.width]))), g)
    finally:
        ref.set_threads(1)


@pytest.mark.gpu
def test_gpu_reconstruct_golden(gpu_lib, pkg):
    _check_reconstruct(lambda v, k: pkg.hrweno_weno.weno(len(v), k, 1e-6).reconstruct(v))


@pytest.mark.gpu
def test_gpu_example1_golden(gpu_lib, pkg):
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    _check_example1(lambda: pkg.hrweno_tvdode.rktvd(pkg.fv.FV(pkg.fv.make_desc(100, width=[g.width])), 100, 3), g)


@pytest.mark.gpu
def test_gpu_example2_golden(gpu_lib, pkg):
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, 250)
    _check_example2(lambda: pkg.hrweno_tvdode.mstvd(pkg.fv.FV(pkg.fv.make_desc((250, 250), flux_model=1, bc=1, width=[g.width, g.width])), 62500), g)
