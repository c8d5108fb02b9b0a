"""Fast mode at large magnitudes (weno_core.cuh: magnitude guard).  The division-light weights form products
~v^9; stencil windows holding |v| >= 2^100 are evaluated on a power-of-two scaled copy, so the fast kernels stay finite
and accurate wherever the reference is (it overflows itself from |v| ~ 1e77 on)."""
import numpy as np
import pytest

from conftest import ex1_ic, ex2_ic

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("scale", [1.0, 1e29, 1e31, 1e34, 1e40, 1e70])
def test_reconstruct_fast_large_magnitudes(gpu_lib, pkg, ref, k, scale):
    rng = np.random.default_rng(5)
    n = 20011
    v = np.concatenate([rng.standard_normal(n // 2), np.where(np.arange(n - n // 2) % 97 < 40, 1.0, -0.5)]) * scale
    v[1000:1100] *= 1e-25  # small values next to large ones inside the same runs
    w = pkg.hrweno_weno.weno(n, k, 1e-6, mode=pkg._abi.MODE_FAST)
    vl, vr = w.reconstruct(v)
    rl, rr = ref.reconstruct(v, k, 1e-6)
    assert np.all(np.isfinite(vl)) and np.all(np.isfinite(vr))
    tol = 6 * EPS * np.max(np.abs(v))
    assert np.max(np.abs(vl - rl)) <= tol and np.max(np.abs(vr - rr)) <= tol


@pytest.mark.parametrize("scale", [1e31, 1e40, 1e70])
def test_fused_stage_fast_large_magnitudes(gpu_lib, pkg, ref, scale):
    """tiled 1D stage (linear flux, Lax-Friedrichs: the flux itself stays finite) and the 2D stage"""
    nc = 30011
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nc)
    v = (ex1_ic(g.center) + 1e-3 * np.random.default_rng(1).standard_normal(nc)) * scale
    kw = dict(n=nc, k=3, flux_model=1, flux_scheme=1, alpha=1.0, width=[g.width])
    got = pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)).rhs(0.0, v)
    want = ref.FV(pkg.fv.make_desc(**kw)).rhs(0.0, v)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want)) <= 64 * EPS * np.max(np.abs(v)) / g.width.min()
    n1, n2 = 300, 200
    g1, g2 = pkg.hrweno_grids.grid1().linear(0.0, 10.0, n1), pkg.hrweno_grids.grid1().linear(0.0, 10.0, n2)
    v2 = (ex2_ic(g1.center, g2.center) + 1e-3 * np.random.default_rng(2).standard_normal((n2, n1))) * scale
    kw = dict(n=(n1, n2), k=3, flux_model=1, bc=1, width=[g1.width, g2.width])
    got = pkg.fv.FV(pkg.fv.make_desc(mode=pkg._abi.MODE_FAST, **kw)).rhs(0.0, v2)
    want = ref.FV(pkg.fv.make_desc(**kw)).rhs(0.0, v2)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want)) <= 64 * EPS * np.max(np.abs(v2)) / min(g1.width.min(), g2.width.min())
