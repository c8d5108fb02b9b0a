"""CPU-only: the host-side mirrors of src/hrweno_grids.f90 (Python `hrweno_grids.grid1`, C++ `hrweno::hrweno_grids::grid1`).

Grids stay on the host (north star); the mirrors exist so that C++ / Python host programs feed the device path the
same widths, edges and centres a Fortran host would.  Pins: every assertion of test/test_grid.f90 (rtol = 1e-5 there),
and the two mirrors against each other -- bit for bit where the reference uses only + - * / and real**integer
(linear, bilinear, geometric), to 2 ulp where it calls libm's log/exp (log grid).
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

RTOL = 1e-5  # test_grid.f90:13


def _check_derived(g, nc):  # the common tail of every test in test_grid.f90
    assert g.ncells == nc
    assert np.array_equal(g.left, g.edges[:nc]) and np.array_equal(g.right, g.edges[1:])
    assert np.array_equal(g.width, g.right - g.left)
    assert np.array_equal(g.center, (g.left + g.right) / 2)


def test_log_reference_test(pkg):  # test_grid.f90:62-92
    g = pkg.hrweno_grids.grid1().log(1e-1, 1e3, 10**4, name="T [K]")
    assert abs(g.edges[0] - 1e-1) <= RTOL * 1e-1 and abs(g.edges[-1] - 1e3) <= RTOL * 1e3
    _check_derived(g, 10**4)
    r = g.edges[1:] / g.edges[:-1]
    assert np.max(np.abs(r / r[0] - 1.0)) < 1e-12  # linear in log(x): constant edge ratio


def test_geometric_reference_test(pkg):  # test_grid.f90:94-126
    g = pkg.hrweno_grids.grid1().geometric(1e1, 1e3, 1.1, 10**2, name="P [W]")
    assert abs(g.edges[0] - 1e1) <= RTOL * 1e1 and abs(g.edges[-1] - 1e3) <= RTOL * 1e3
    _check_derived(g, 100)
    assert np.max(np.abs(g.width[1:] / g.width[:-1] - 1.1)) < 1e-9  # width(i+1) = R*width(i)  (grids.f90:184)


def test_bilinear_reference_test(pkg):  # test_grid.f90:128-160
    nc = (124, 365)
    g = pkg.hrweno_grids.grid1().bilinear(0.0, 1e1, 1e3, nc, name="W [J]")
    assert g.edges[0] == 0.0 and abs(g.edges[nc[0]] - 1e1) <= RTOL * 1e1 and abs(g.edges[-1] - 1e3) <= RTOL * 1e3
    _check_derived(g, sum(nc))


def test_grid_input_validation(pkg):  # grids.f90:66-72, 110-118, 157-165, 205-213: the reference would error stop
    G = pkg.hrweno_grids.grid1
    for bad in (lambda: G().linear(1.0, 1.0, 10), lambda: G().log(0.0, 1.0, 10), lambda: G().log(2.0, 1.0, 10),
                lambda: G().geometric(0.0, 1.0, 0.0, 10), lambda: G().geometric(1.0, 0.0, 1.1, 10),
                lambda: G().bilinear(0.0, 0.0, 1.0, (2, 2)), lambda: G().bilinear(0.0, 2.0, 1.0, (2, 2)),
                lambda: G().bilinear(0.0, 1.0, 2.0, (0, 2))):
        with pytest.raises(ValueError):
            bad()


def test_integer_power_is_binary_exponentiation(pkg):
    """`ratio**i` with integer i (grids.f90:223-225) is repeated squaring in gfortran, not libm pow"""
    powi = pkg.hrweno_grids._powi
    x = 1.1
    assert powi(x, np.array([0]))[0] == 1.0 and powi(x, np.array([1]))[0] == x
    assert powi(x, np.array([5]))[0] == x * ((x * x) * (x * x))       # 101b: y = x; x4 = (x^2)^2; y = y*x4
    assert powi(x, np.array([6]))[0] == (x * x) * ((x * x) * (x * x))  # 110b: y = x^2; y = y*x^4
    i = np.arange(200)
    assert np.max(np.abs(powi(x, i) / x**i - 1.0)) < 1e-13


def test_cpp_mirror_equals_python_mirror(pkg):
    exe = os.path.join(ROOT, "examples", "grid_dump")
    assert os.path.exists(exe), "examples not built (python __graft_entry__.py)"
    out = subprocess.run([exe], capture_output=True, text=True, check=True, timeout=60).stdout.splitlines()
    got = {}
    for line in out:
        name, n, *vals = line.split()
        got[name] = np.array([float.fromhex(v) for v in vals])
        assert got[name].size == int(n) + 1
    G = pkg.hrweno_grids.grid1
    assert np.array_equal(got["linear"], G().linear(-5.0, 5.0, 100).edges)
    assert np.array_equal(got["geometric"], G().geometric(1e1, 1e3, 1.1, 100).edges)
    assert np.array_equal(got["bilinear"], G().bilinear(0.0, 1e1, 1e3, (124, 365)).edges)
    ref_log = G().log(1e-1, 1e3, 10**4).edges
    assert np.max(np.abs(got["log"] / ref_log - 1.0)) <= 2 * np.finfo(float).eps  # libm log/exp: numpy vs glibc scalar
