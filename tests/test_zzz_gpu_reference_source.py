"""GPU: the CUDA path through the C ABI against the fixtures produced by executing the reference's own source text
(tests/golden/ref_exec_*.npz, see tests/test_reference_source_exec.py and tests/golden/make_ref_exec_golden.py).
Bar: bit-identical (strict mode / the general kernel).  Sorted last on purpose (`pytest -x`)."""
import numpy as np
import pytest

from test_reference_source_exec import _example1, _example1_variant, _example2, gold

@pytest.mark.gpu
def test_gpu_example1_equals_reference_source(gpu_lib, pkg):
    g = gold("example1")
    ode = _example1(pkg, lambda d: pkg.hrweno_tvdode.rktvd(pkg.fv.FV(d), 100, 3), g)
    assert ode.fevals == 3603


@pytest.mark.gpu
def test_gpu_example2_equals_reference_source(gpu_lib, pkg):
    _example2(pkg, lambda fv: pkg.hrweno_tvdode.mstvd(fv, 1600), gold("example2_40"), 40, 40, 5e-3, 5.0, mod=pkg.fv)
    _example2(pkg, lambda fv: pkg.hrweno_tvdode.mstvd(fv, 62500), gold("example2_250_first2"), 250, 250, 5e-3, 5.0, mod=pkg.fv)


@pytest.mark.gpu
def test_gpu_general_path_equals_reference_source(gpu_lib, pkg):
    _example2(pkg, lambda fv: pkg.hrweno_tvdode.mstvd(fv, 24 * 18), gold("example2_growth"), 24, 18, 2.5e-4, 0.5, growth=True,
              mod=pkg.fv)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3])
def test_gpu_reconstruct_equals_reference_source(gpu_lib, pkg, k):
    g = gold("reconstruct")
    for gname in ("none", "uniform", "cubic"):
        w = pkg.hrweno_weno.weno(30, k, 1e-6) if gname == "none" else pkg.hrweno_weno.weno(30, k, 1e-6, xedges=g["xe_" + gname])
        if gname != "none":
            assert np.array_equal(w.cnu, g[f"cnu_{gname}_k{k}"])
        for vname in ("pulse", "rand"):
            vl, vr = w.reconstruct(g["v_" + vname])
            assert np.array_equal(vl, g[f"vl_{gname}_{vname}_k{k}"]) and np.array_equal(vr, g[f"vr_{gname}_{vname}_k{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_gpu_example1_k_order_sweep_equals_reference_source(gpu_lib, pkg, k, order):
    g = gold("example1_sweep")
    make = lambda d, o: pkg.hrweno_tvdode.rktvd(pkg.fv.FV(d), 100, o)  # noqa: E731
    ode = _example1_variant(pkg, make, g[f"t_k{k}_o{order}"], g[f"u_k{k}_o{order}"], k, order)
    assert ode.fevals == int(g[f"fevals_k{k}_o{order}"])


@pytest.mark.gpu
def test_gpu_example1_lax_friedrichs_equals_reference_source(gpu_lib, pkg):
    g = gold("example1_lf")
    snaps = {ii: g[f"u_{ii}"] for ii in (0, 50, 100)}
    make = lambda d, o: pkg.hrweno_tvdode.rktvd(pkg.fv.FV(d), 100, o)  # noqa: E731
    _example1_variant(pkg, make, g["times"], None, 3, 3, scheme=1, upto=100, snaps=snaps)


@pytest.mark.gpu
def test_gpu_fast_mode_example1_within_north_star_tolerance_of_reference_source(gpu_lib, pkg):
    """north star: within 1e-12 relative (normwise) of the reference at every output time -- fast arithmetic mode against
    example1 as executed from the reference's source (stored outputs 0, 1, 50, 100)"""
    g = gold("example1")
    d = pkg.fv.make_desc(100, k=3, eps=1e-6, width=[g["width"]], mode=pkg._abi.MODE_FAST)
    ode = pkg.hrweno_tvdode.rktvd(pkg.fv.FV(d), 100, 3)
    u, t = np.clip(1.0 + (-1.5 / 6.0) * (g["center"] + 4.0), -0.5, 1.0), 0.0
    worst = 0.0
    for ii in range(101):
        t = ode.integrate(u, t, 12.0 * ii / 100, 1e-2)
        assert t == g["times"][ii]
        if f"u_{ii}" in g:
            r = g[f"u_{ii}"]
            worst = max(worst, float(np.max(np.abs(u - r)) / np.max(np.abs(r))))
    assert worst <= 1e-12, f"fast-mode drift {worst:.3e}"
