"""CPU-only: the oracle's restatement of the general finite-volume path (non-uniform grids through
`weno(ncells,k,eps,xedges)`, weno.f90:100-112,177,221-297, and x-dependent fluxes through the `x` argument of the
flux callback, fluxes.f90:12-18, example2:100-101,109-110,140,153).

The reference has no test of an rhs on a non-uniform grid; its own pins for this path are test_hrweno.f90:69-161
(cnu == c on a uniform grid to rtol 1e-5; pulse reconstruction on a cubic grid to 1e-6).  Those are re-run on the
reconstruction in test_oracle_reference_pins.py; here the composed rhs is held to the structural facts that follow from
them and to an independent NumPy restatement bit for bit.
"""
import numpy as np
import pytest

from conftest import ex2_ic


def _grids(pkg, n1, n2):
    g1 = pkg.hrweno_grids.grid1().geometric(0.1, 10.0, 1.05, n1)
    g2 = pkg.hrweno_grids.grid1().log(0.2, 8.0, n2)
    return g1, g2


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("scheme,model,bc", [(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 1), (0, 0, 1)])
def test_rhs1d_general_c_oracle_equals_numpy_oracle(pkg, ref, npo, k, scheme, model, bc):
    rng = np.random.default_rng(10 * k + scheme)
    nc = 61
    g = pkg.hrweno_grids.grid1().geometric(0.5, 10.0, 1.07, nc)
    v = rng.standard_normal((3, nc))
    fc = g.edges**2
    d = pkg.fv.make_desc(n=nc, rows=3, k=k, width=[g.width], flux_model=model, flux_scheme=scheme, alpha=1.3,
                         flux_coef=(0.7, 1.0), bc=bc)
    f = ref.FV(d)
    f.set_xedges(0, g.edges)
    f.set_flux_coef(0, fc)
    a = f.rhs(0.0, v.ravel()).reshape(3, nc)
    b = npo.rhs1d(v, g.width, k=k, scheme=["godunov", "lax_friedrichs"][scheme], model=["burgers", "linear"][model],
                  coef=0.7, alpha=1.3, bc=["copy", "zero"][bc], cnu=npo.calc_cnu(g.edges, k), fcoef=fc)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_rhs2d_general_c_oracle_equals_numpy_oracle(pkg, ref, npo, k):
    rng = np.random.default_rng(k)
    n1, n2 = 23, 17
    g1, g2 = _grids(pkg, n1, n2)
    v = rng.standard_normal((n2, n1))
    d = pkg.fv.make_desc(n=(n1, n2), k=k, width=[g1.width, g2.width], flux_model=1, bc=1)
    f = ref.FV(d)
    f.set_xedges(0, g1.edges)
    f.set_xedges(1, g2.edges)
    f.set_flux_coef(0, g1.edges**2, None)      # flux1 = v*x(1)**2      (example2:140)
    f.set_flux_coef(1, g2.edges, g1.center)    # flux2 = v*x(1)*x(2)    (example2:153)
    a = f.rhs(0.0, v.ravel()).reshape(n2, n1)
    b = npo.rhs2d(v, g1.width, g2.width, k=k, cnu=(npo.calc_cnu(g1.edges, k), npo.calc_cnu(g2.edges, k)),
                  fcoef=(g1.edges**2, g2.edges), ccoef=(None, g1.center))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_cnu_on_a_uniform_grid_gives_the_uniform_rhs(pkg, ref, k):
    """test_hrweno.f90:104-108 (cnu == c on a uniform grid, rtol 1e-5) carried through the rhs"""
    nc = 100
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    x = g.center
    v = np.clip(1.0 - 0.25 * (x + 4.0), -0.5, 1.0)
    d = pkg.fv.make_desc(n=nc, k=k, width=[g.width])
    base = ref.FV(d).rhs(0.0, v)
    f = ref.FV(d)
    f.set_xedges(0, g.edges)
    got = f.rhs(0.0, v)
    assert np.max(np.abs(got - base)) <= 1e-9 * max(1.0, np.max(np.abs(base)))


def test_unit_coefficients_change_nothing(pkg, ref):
    """multiplying by 1.0 is exact: coefficient arrays of ones reproduce the x-independent path bit for bit"""
    rng = np.random.default_rng(3)
    n1, n2 = 19, 13
    g1, g2 = _grids(pkg, n1, n2)
    v = rng.standard_normal(n1 * n2)
    for model, scheme in ((0, 0), (1, 1)):
        d = pkg.fv.make_desc(n=(n1, n2), k=3, width=[g1.width, g2.width], flux_model=model, flux_scheme=scheme, bc=1)
        base = ref.FV(d).rhs(0.0, v)
        f = ref.FV(d)
        f.set_flux_coef(0, np.ones(n1 + 1), np.ones(n2))
        f.set_flux_coef(1, np.ones(n2 + 1), np.ones(n1))
        assert np.array_equal(f.rhs(0.0, v), base)


def test_zero_flux_walls_conserve_mass_with_growth_terms(pkg, ref):
    """finite-volume telescoping: with fedges = 0 on the walls (example2:117-120), sum(vdot*w1*w2) vanishes whatever the
    interior fluxes are -- non-uniform grid, growth-type coefficients included"""
    n1, n2 = 40, 30
    g1, g2 = _grids(pkg, n1, n2)
    v = ex2_ic(g1.center, g2.center).ravel()
    d = pkg.fv.make_desc(n=(n1, n2), k=3, width=[g1.width, g2.width], flux_model=1, bc=1)
    f = ref.FV(d)
    f.set_xedges(0, g1.edges)
    f.set_xedges(1, g2.edges)
    f.set_flux_coef(0, g1.edges**2, None)
    f.set_flux_coef(1, g2.edges, g1.center)
    vdot = f.rhs(0.0, v).reshape(n2, n1)
    scale = float(np.sum(np.abs(vdot) * g1.width[None, :] * g2.width[:, None]))
    assert scale > 0.0
    assert abs(float(np.sum(vdot * g1.width[None, :] * g2.width[:, None]))) <= 1e-13 * scale


def test_general_setters_validate(pkg, ref):
    g = pkg.hrweno_grids.grid1().linear(0.0, 1.0, 10)
    f = ref.FV(pkg.fv.make_desc(n=10, k=3, width=[g.width]))
    L = ref.lib()
    assert L.hrweno_ref_fv_set_xedges(f._h, 1, g.edges.ctypes.data) == pkg._abi.EINVAL  # a 1D operator has no axis 1
    assert L.hrweno_ref_fv_set_flux_coef(f._h, 0, g.edges.ctypes.data, g.center.ctypes.data) == pkg._abi.EINVAL  # cross needs 2D
    assert L.hrweno_ref_fv_set_flux_coef(f._h, 0, g.edges.ctypes.data, None) == pkg._abi.OK


@pytest.mark.parametrize("k", [1, 2, 3])
def test_slab_tables_from_global_edges(ref, k):
    """design fact for a slab decomposition of the non-uniform path: cnu(:,:,i) reads the edges
    i-k .. i+k (weno.f90:263-292: xl(i-r+m) with r = -1..k-1, m = 0..k), so the tables of a slab [off, off+n) computed from
    its own edges plus k-1 neighbour edges on the left and k on the right are the global tables bit for bit; one edge
    fewer on the right and they are not"""
    rng = np.random.default_rng(k)
    nc = 60
    xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.1, 2.0, nc))])
    glob = ref.calc_cnu(xe, k)
    for off, n in ((0, 20), (20, 20), (40, 20), (7, 3), (55, 5)):
        lo, hi = max(0, off - (k - 1)), min(nc, off + n + k)
        loc = ref.calc_cnu(xe[lo : hi + 1], k)
        assert np.array_equal(loc[off - lo : off - lo + n], glob[off : off + n])
    lo, hi = 20 - (k - 1), 40 + k - 1  # one neighbour edge too few on the right
    loc = ref.calc_cnu(xe[lo : hi + 1], k)
    if k >= 2:  # k = 1 reconstructs cell averages: its tables are 1 on any grid
        assert not np.array_equal(loc[20 - lo : 20 - lo + 20], glob[20:40])
