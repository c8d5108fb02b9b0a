"""CPU check of the FAST-mode algebra (hr-weno_b200/csrc/weno_core.cuh: weno_run_k3_fast / weno_run_k2_fast).

The device code evaluates WENO3/WENO5 in differences of the cell averages with division-light weights; its NumPy
restatement (oracle/np_oracle.py: reconstruct_fast) must agree with the reference-order reconstruction
(weno.f90:174-216) to a few ULP of max|v| on smooth, discontinuous, noisy and constant data, including the cells
whose stencil reaches the replicated ghosts (weno.f90:171-173).  The GPU tests hold the kernel itself to the same bar.
"""
import numpy as np
import pytest

from oracle import np_oracle as o

EPS = np.finfo(float).eps


def _cases():
    rng = np.random.default_rng(7)
    n = 400
    x = np.linspace(-5, 5, n)
    yield "noise", rng.standard_normal(n)
    yield "smooth", np.sin(x) + 0.3 * np.cos(3 * x)
    yield "ramp+noise", np.clip(1.0 - 0.25 * (x + 4.0), -0.5, 1.0) + 1e-3 * rng.standard_normal(n)
    yield "step", np.where(x < 0.3, 1.0, 0.0)
    yield "constant", np.full(50, 0.7)
    yield "rows", rng.standard_normal((3, 64))
    yield "short", rng.standard_normal(5)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_fast_formulas_match_reference_order_to_a_few_ulp(k):
    for name, v in _cases():
        vl, vr = o.reconstruct(v, k, 1e-6)
        fl, fr = o.reconstruct_fast(v, k, 1e-6)
        tol = 4 * EPS * max(np.max(np.abs(v)), 1e-300)
        assert np.max(np.abs(vl - fl)) <= tol, (name, k)
        assert np.max(np.abs(vr - fr)) <= tol, (name, k)


def test_fast_formulas_weights_are_scale_consistent():
    """scaling v by 2^m scales vl, vr exactly when eps scales by 4^m (beta' = 4 beta and eps' = 4 eps are exact)"""
    v = np.random.default_rng(3).standard_normal(200)
    fl, fr = o.reconstruct_fast(v, 3, 1e-6)
    gl, gr = o.reconstruct_fast(8.0 * v, 3, 64.0 * 1e-6)
    assert np.array_equal(gl, 8.0 * fl) and np.array_equal(gr, 8.0 * fr)


@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("scale", [1e30, 1e33, 1e34, 1e40, 1e60, 1e70])
def test_fast_formulas_stay_finite_and_accurate_at_large_magnitudes(k, scale):
    """the division-light weights form products ~v^9; beyond |v| ~ 1e33 they overflowed while the reference
    (which only forms (eps+beta)^2 ~ v^4) stays finite up to ~1e76.  With the magnitude guard the fast formulas are
    evaluated on a stencil scaled by a power of two and stay within a few ULP of the reference order."""
    rng = np.random.default_rng(11)
    for v in (rng.standard_normal(300), np.where(np.arange(300) < 150, 1.0, -0.5), np.full(64, 0.7),
              np.concatenate([rng.standard_normal(100), 1e-30 * rng.standard_normal(100)])):
        v = v * scale
        vl, vr = o.reconstruct(v, k, 1e-6)
        fl, fr = o.reconstruct_fast(v, k, 1e-6)
        assert np.all(np.isfinite(fl)) and np.all(np.isfinite(fr))
        tol = 4 * EPS * np.max(np.abs(v))
        assert np.max(np.abs(vl - fl)) <= tol and np.max(np.abs(vr - fr)) <= tol
