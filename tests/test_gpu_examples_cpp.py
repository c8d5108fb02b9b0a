"""The C++ host programs (examples/*.cpp, mirrors of the reference's example programs written against the
reference-shaped C++ host API) run on the GPU and must reproduce the oracle's final state bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, ex1_ic, ex2_ic

pytestmark = pytest.mark.gpu


def _run(exe, tmp_path):
    path = os.path.join(ROOT, "examples", exe)
    assert os.path.exists(path), "examples not built (python __graft_entry__.py)"
    out = subprocess.run([path, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout, np.fromfile(tmp_path / "u_final.bin")


def test_example1_cpp(gpu_lib, pkg, ref, tmp_path):
    stdout, u = _run("example1_burgers_1d_fv", tmp_path)
    assert "fevals = 3603" in stdout and "12.009999999999788" in stdout
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, 100)
    ode = ref.rktvd(ref.FV(pkg.fv.make_desc(100, width=[g.width])), 3)
    ur, t = ex1_ic(g.center), 0.0
    for ii in range(101):
        t = ode.integrate(ur, t, 12.0 * ii / 100, 1e-2)
    assert np.array_equal(u, ur)
    rows = open(tmp_path / "u.txt").read().splitlines()
    assert len(rows) == 102 and len(rows[1].split()) == 101  # header + 101 output rows (example1:61-65,172-177)
    assert float(rows[1].split()[0]) == 0.01  # first output is at t = 0.01 (strict is_done)


def test_example2_cpp(gpu_lib, pkg, ref, tmp_path):
    stdout, u = _run("example2_pbe_2d_fv", tmp_path)
    assert "fevals = 1009" in stdout and "5.0049999999999155" in stdout
    n = 250
    g = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n)
    ref.set_threads(min(16, ref.max_threads()))
    try:
        ode = ref.mstvd(ref.FV(pkg.fv.make_desc((n, n), flux_model=1, bc=1, width=[g.width, g.width])))
        ur, t = ex2_ic(g.center, g.center).reshape(-1), 0.0
        for ii in range(101):
            t = ode.integrate(ur, t, 5.0 * ii / 100, 5e-3)
    finally:
        ref.set_threads(1)
    assert np.array_equal(u, ur)
