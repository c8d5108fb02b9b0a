"""GPU parity of the general fused stage (csrc/fvgen.cu, K7): non-uniform grids through `weno(ncells,k,eps,xedges)`
(weno.f90:100-112,177,221-297) and x-dependent fluxes (fluxes.f90:12-18; example2:100-101,109-110,140,153).

Bar: bit-identical to the oracle (the general kernel always runs the reference's operation order, in both modes).
This file sorts after the established suite on purpose: under `pytest -x` a failure here cannot hide those results.
"""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic, normwise

pytestmark = pytest.mark.gpu


def _pair(pkg, ref, kw, xedges=(None, None), face=(None, None), cross=(None, None)):
    """the product operator and the oracle operator with the same general-path settings"""
    out = []
    for mod in (pkg.fv, ref):
        f = mod.FV(pkg.fv.make_desc(**kw))
        for a in range(2):
            if xedges[a] is not None:
                f.set_xedges(a, xedges[a])
            if face[a] is not None or cross[a] is not None:
                f.set_flux_coef(a, face[a], cross[a])
        out.append(f)
    return out


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nc", [2, 3, 5, 100, 255, 256, 257, 1000, 5000])
def test_rhs1d_nonuniform_bitwise_sizes(gpu_lib, pkg, ref, k, nc):
    rng = np.random.default_rng(31 * k + nc)
    g = pkg.hrweno_grids.grid1().geometric(0.5, 10.0, 1.0 + 2.0 / nc, nc)
    v = ex1_ic(g.center - 5.0) + 1e-3 * rng.standard_normal(nc)
    fv, rfv = _pair(pkg, ref, dict(n=nc, k=k, width=[g.width]), xedges=(g.edges, None))
    assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))


@pytest.mark.parametrize("scheme,model,bc", [(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 1), (0, 0, 1), (1, 1, 0)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_rhs1d_general_variants_batched(gpu_lib, pkg, ref, k, scheme, model, bc):
    """rows of independent problems, both schemes, both models, both boundary rules, x-dependent face coefficient"""
    rng = np.random.default_rng(7 * k + scheme + 2 * model)
    nc, rows = 300, 5
    g = pkg.hrweno_grids.grid1().log(0.3, 12.0, nc)
    v = rng.standard_normal((rows, nc)).ravel()
    kw = dict(n=nc, rows=rows, k=k, width=[g.width], flux_model=model, flux_scheme=scheme, flux_coef=(0.7, 1.0),
              alpha=1.3, bc=bc)
    fv, rfv = _pair(pkg, ref, kw, xedges=(g.edges, None), face=(g.edges**2, None))
    assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))


def test_rhs1d_coefficients_only_on_a_linear_grid(gpu_lib, pkg, ref):
    """x-dependent flux with the uniform tables, widths from the analytic linear grid (dictionary expanded on device)"""
    nc = 777
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    v = ex1_ic(g.center) + 1e-3 * np.random.default_rng(5).standard_normal(nc)
    fc = 1.0 + 0.1 * g.edges**2
    fv = pkg.fv.FV(pkg.fv.make_desc(n=nc, k=3, linear=(-5.0, 5.0)))
    fv.set_flux_coef(0, fc)
    rfv = ref.FV(pkg.fv.make_desc(n=nc, k=3, linear=(-5.0, 5.0)))
    rfv.set_flux_coef(0, fc)
    assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))


def test_unit_coefficients_equal_the_tuned_path(gpu_lib, pkg):
    """coefficient arrays of ones (exact multiplications) through the general kernel == the tuned strict kernels"""
    rng = np.random.default_rng(11)
    nc = 1500
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    v = ex1_ic(g.center) + 1e-3 * rng.standard_normal(nc)
    kw = dict(n=nc, k=3, width=[g.width])
    base = pkg.fv.FV(pkg.fv.make_desc(**kw)).rhs(0.0, v)
    fv = pkg.fv.FV(pkg.fv.make_desc(**kw))
    fv.set_flux_coef(0, np.ones(nc + 1))
    assert np.array_equal(fv.rhs(0.0, v), base)
    n1, n2 = 70, 45
    g1 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1)
    g2 = pkg.hrweno_grids.grid1().linear(0.0, 7.0, n2)
    v = ex2_ic(g1.center, g2.center).ravel() + 1e-3 * rng.standard_normal(n1 * n2)
    kw = dict(n=(n1, n2), k=3, flux_model=1, bc=1, width=[g1.width, g2.width])
    base = pkg.fv.FV(pkg.fv.make_desc(**kw)).rhs(0.0, v)
    fv = pkg.fv.FV(pkg.fv.make_desc(**kw))
    fv.set_flux_coef(0, np.ones(n1 + 1), np.ones(n2))
    fv.set_flux_coef(1, np.ones(n2 + 1), np.ones(n1))
    assert np.array_equal(fv.rhs(0.0, v), base)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("n1,n2", [(2, 2), (5, 7), (31, 9), (32, 8), (33, 9), (65, 17), (150, 90)])
def test_rhs2d_general_bitwise_sizes(gpu_lib, pkg, ref, k, n1, n2):
    """geometric x log grid, growth terms flux1 = v*x(1)**2, flux2 = v*x(1)*x(2) (example2:140,153), zero-flux walls"""
    rng = np.random.default_rng(1000 * n1 + n2 + k)
    g1 = pkg.hrweno_grids.grid1().geometric(0.1, 10.0, 1.03, n1)
    g2 = pkg.hrweno_grids.grid1().log(0.2, 8.0, n2)
    v = rng.standard_normal(n1 * n2)
    kw = dict(n=(n1, n2), k=k, flux_model=1, bc=1, width=[g1.width, g2.width])
    fv, rfv = _pair(pkg, ref, kw, xedges=(g1.edges, g2.edges), face=(g1.edges**2, g2.edges), cross=(None, g1.center))
    assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))


@pytest.mark.parametrize("scheme,model,bc", [(0, 0, 0), (1, 0, 0), (1, 1, 1), (0, 0, 1), (1, 1, 0)])
def test_rhs2d_general_variants(gpu_lib, pkg, ref, scheme, model, bc):
    """one axis non-uniform only, the other on the uniform tables; coefficients on one axis only"""
    rng = np.random.default_rng(17 + scheme)
    n1, n2 = 97, 41
    g1 = pkg.hrweno_grids.grid1().geometric(0.0, 4.0, 1.02, n1)
    g2 = pkg.hrweno_grids.grid1().linear(0.5, 9.0, n2)
    v = rng.standard_normal(n1 * n2)
    kw = dict(n=(n1, n2), k=3, flux_model=model, flux_scheme=scheme, flux_coef=(1.7, -0.6), alpha=1.3, bc=bc,
              width=[g1.width, g2.width])
    fv, rfv = _pair(pkg, ref, kw, xedges=(g1.edges, None), face=(None, 1.0 + g2.edges), cross=(None, g1.center))
    assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [2, 3])
def test_rktvd_fused_nonuniform_1d_bitwise(gpu_lib, pkg, ref, k, order):
    """rktvd orders 1-3 on a geometric grid (the small-problem single-launch path must stand aside)"""
    nc = 400
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.004, nc)
    fv, rfv = _pair(pkg, ref, dict(n=nc, k=k, width=[g.width]), xedges=(g.edges, None))
    ode, rode = pkg.hrweno_tvdode.rktvd(fv, nc, order), ref.rktvd(rfv, order)
    u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
    for tout in (0.0, 0.05, 0.05, 0.2):
        t = ode.integrate(u, t, tout, 2e-3)
        tr = rode.integrate(ur, tr, tout, 2e-3)
        assert t == tr and np.array_equal(u, ur), f"normwise {normwise(u, ur):.3e}"
    assert ode.fevals == rode.fevals


def test_mstvd_fused_nonuniform_1d_bitwise(gpu_lib, pkg, ref):
    nc = 500
    g = pkg.hrweno_grids.grid1().geometric(-5.0, 5.0, 1.003, nc)
    fv, rfv = _pair(pkg, ref, dict(n=nc, k=3, width=[g.width]), xedges=(g.edges, None))
    ode, rode = pkg.hrweno_tvdode.mstvd(fv, nc), ref.mstvd(rfv)
    u, ur, t, tr = ex1_ic(g.center), ex1_ic(g.center), 0.0, 0.0
    for tout in (0.0, 0.1, 0.1, 0.25):
        t = ode.integrate(u, t, tout, 2e-3)
        tr = rode.integrate(ur, tr, tout, 2e-3)
        assert t == tr and np.array_equal(u, ur), f"normwise {normwise(u, ur):.3e}"
    assert ode.fevals == rode.fevals


def test_pbe_growth_on_geometric_grid_mstvd_every_output(gpu_lib, pkg, ref):
    """example2's program with what its comments point at: geometric grids, weno(nc,k,eps,xedges) per axis, growth
    fluxes v*x(1)**2 and v*x(1)*x(2), mstvd, zero-flux walls -- bitwise at every output time, mass conserved"""
    n = 60
    g1 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.02, n)
    g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.03, n)
    kw = dict(n=(n, n), k=3, eps=1e-6, flux_model=1, bc=1, width=[g1.width, g2.width])
    fv, rfv = _pair(pkg, ref, kw, xedges=(g1.edges, g2.edges), face=(g1.edges**2, g2.edges), cross=(None, g1.center))
    ode, rode = pkg.hrweno_tvdode.mstvd(fv, n * n), ref.mstvd(rfv)
    u = ex2_ic(g1.center, g2.center).reshape(-1)
    ur = u.copy()
    w = (g1.width[None, :] * g2.width[:, None]).ravel()
    mass0 = float(np.sum(u * w))
    t, tr = 0.0, 0.0
    for ii in range(21):
        tout = 0.1 * ii / 20
        t = ode.integrate(u, t, tout, 2.5e-4)
        tr = rode.integrate(ur, tr, tout, 2.5e-4)
        assert t == tr and np.array_equal(u, ur), f"output {ii}: normwise {normwise(u, ur):.3e}"
    assert ode.fevals == rode.fevals
    assert abs(float(np.sum(u * w)) - mass0) <= 1e-12 * mass0
    assert u.min() > -1e-3 and u.max() < 1.0  # stable and essentially non-oscillatory at CFL 0.1


def test_rktvd_fused_general_2d_bitwise(gpu_lib, pkg, ref):
    n1, n2 = 70, 45
    g1 = pkg.hrweno_grids.grid1().log(0.5, 10.0, n1)
    g2 = pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.05, n2)
    kw = dict(n=(n1, n2), k=3, flux_model=1, bc=1, width=[g1.width, g2.width])
    for order in (1, 2, 3):
        fv, rfv = _pair(pkg, ref, kw, xedges=(g1.edges, g2.edges), face=(g1.edges, None))
        ode, rode = pkg.hrweno_tvdode.rktvd(fv, n1 * n2, order), ref.rktvd(rfv, order)
        u = ex2_ic(g1.center, g2.center).reshape(-1)
        ur = u.copy()
        t, tr = 0.0, 0.0
        for tout in (0.0, 0.02, 0.05):
            t = ode.integrate(u, t, tout, 1e-3)
            tr = rode.integrate(ur, tr, tout, 1e-3)
            assert t == tr and np.array_equal(u, ur), f"order {order}: normwise {normwise(u, ur):.3e}"


def test_general_path_ignores_fast_mode(gpu_lib, pkg, ref):
    """mode = FAST selects the tuned kernels' arithmetic only; the general kernel stays in the reference order"""
    nc = 300
    g = pkg.hrweno_grids.grid1().geometric(0.5, 10.0, 1.01, nc)
    v = np.random.default_rng(2).standard_normal(nc)
    fv = pkg.fv.FV(pkg.fv.make_desc(n=nc, k=3, width=[g.width], mode=pkg._abi.MODE_FAST))
    fv.set_xedges(0, g.edges)
    rfv = ref.FV(pkg.fv.make_desc(n=nc, k=3, width=[g.width]))
    rfv.set_xedges(0, g.edges)
    assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))


def test_general_setters_validate(gpu_lib, pkg):
    g = pkg.hrweno_grids.grid1().linear(0.0, 1.0, 10)
    fv = pkg.fv.FV(pkg.fv.make_desc(n=10, k=3, width=[g.width]))
    with pytest.raises(pkg.HrwenoError):
        fv.set_xedges(0, g.edges[:-1])  # size(xedges) /= ncells + 1 (weno.f90:101-108)
    with pytest.raises(pkg.HrwenoError):
        fv.set_flux_coef(0, g.edges, g.center)  # a cross coefficient needs ndim == 2
    lib = pkg.lib()
    assert lib.hrweno_fv_set_xedges(fv._h, 1, g.edges.ctypes.data) == pkg._abi.EINVAL
    assert lib.hrweno_fv_set_xedges(fv._h, 0, None) == pkg._abi.EINVAL
    fv.set_flux_coef(0, 1.0 + g.edges)
    import torch

    v = torch.zeros(10, dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    with pytest.raises(pkg.HrwenoError):
        fv.max_wavespeed_dev(v.data_ptr(), out.data_ptr())


def test_example3_cpp_growth_on_geometric_grids(gpu_lib, pkg, ref, tmp_path):
    """examples/example3_pbe_2d_growth.cpp: example2's driver with geometric grids, per-axis xedges and the growth fluxes
    of example2:140,153, written against the C++ host mirror; final state bit-identical to the oracle on the same grids"""
    import os
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "examples", "example3_pbe_2d_growth")
    assert os.path.exists(exe), "examples not built (python __graft_entry__.py)"
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    e1, e2 = np.fromfile(tmp_path / "edges1.bin"), np.fromfile(tmp_path / "edges2.bin")
    u = np.fromfile(tmp_path / "u_final.bin")
    n1, n2 = e1.size - 1, e2.size - 1
    assert (n1, n2) == (80, 60)
    # the C++ and the Python mirror of grid1%geometric agree bit for bit (tests/test_grids.py); use the program's own edges
    assert np.array_equal(e1, pkg.hrweno_grids.grid1().geometric(0.0, 10.0, 1.02, n1).edges)
    w1, w2, c1, c2 = e1[1:] - e1[:-1], e2[1:] - e2[:-1], (e1[:-1] + e1[1:]) / 2, (e2[:-1] + e2[1:]) / 2
    rfv = ref.FV(pkg.fv.make_desc((n1, n2), k=3, eps=1e-6, flux_model=1, bc=1, width=[w1, w2]))
    rfv.set_xedges(0, e1)
    rfv.set_xedges(1, e2)
    rfv.set_flux_coef(0, e1 * e1, None)
    rfv.set_flux_coef(1, e2, c1)
    rode = ref.mstvd(rfv)
    ur, t = ex2_ic(c1, c2).reshape(-1), 0.0
    for ii in range(21):
        t = rode.integrate(ur, t, 0.1 * ii / 20, 2.5e-4)
    assert ur.min() > -1e-3 and ur.max() < 1.0  # a stable, essentially non-oscillatory run (CFL 0.1)
    assert f"fevals = {rode.fevals}" in out.stdout
    assert np.array_equal(u, ur), f"normwise {normwise(u, ur):.3e}"


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_single_cell_rows_zero_flux(gpu_lib, pkg, ref, k, general):
    """the smallest legal problem, ncells = 1 (weno.f90:72-76 only asks for ncells > 0): every stencil cell is an edge
    replica and both faces are walls, so vdot = -(0 - 0)/w; 1D rows and a 1 x n / n x 1 2D grid, tuned and general path"""
    rng = np.random.default_rng(k)
    w = np.array([0.37])
    for rows in (1, 4):
        v = rng.standard_normal(rows)
        kw = dict(n=1, rows=rows, k=k, width=[w], bc=1)
        fv, rfv = pkg.fv.FV(pkg.fv.make_desc(**kw)), ref.FV(pkg.fv.make_desc(**kw))
        if general:
            fv.set_flux_coef(0, np.array([2.0, 3.0]))
            rfv.set_flux_coef(0, np.array([2.0, 3.0]))
        assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))
    g = pkg.hrweno_grids.grid1().geometric(0.0, 1.0, 1.1, 9)
    for shape, widths, ax in (((9, 1), [g.width, w], 0), ((1, 9), [w, g.width], 1)):
        v = rng.standard_normal(9)
        kw = dict(n=shape, k=k, width=widths, flux_model=1, bc=1)
        fv, rfv = pkg.fv.FV(pkg.fv.make_desc(**kw)), ref.FV(pkg.fv.make_desc(**kw))
        if general:
            fv.set_xedges(ax, g.edges)
            rfv.set_xedges(ax, g.edges)
        assert np.array_equal(fv.rhs(0.0, v), rfv.rhs(0.0, v))
