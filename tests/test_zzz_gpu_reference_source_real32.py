"""GPU: the REAL32 entry points (hrweno_*_f32) through the C ABI against the fixtures produced by executing the
reference's own source text as its REAL32 build (tests/golden/ref_exec_f32_*.npz, see
tests/test_reference_source_exec_real32.py).  Bar: bit-identical."""
import numpy as np
import pytest

from test_reference_source_exec_real32 import _example1, _example2, _grid32, gold


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3])
def test_gpu_real32_reconstruct_equals_reference_source(gpu_lib, pkg, k):
    g = gold("reconstruct")
    for gname in ("none", "uniform", "cubic"):
        w = pkg.real32.weno(30, k, 1e-6) if gname == "none" else pkg.real32.weno(30, k, 1e-6, xedges=g["xe_" + gname])
        for vname in ("pulse", "rand"):
            vl, vr = w.reconstruct(g["v_" + vname])
            assert np.array_equal(vl, g[f"vl_{gname}_{vname}_k{k}"]) and np.array_equal(vr, g[f"vr_{gname}_{vname}_k{k}"]), (gname, vname)


@pytest.mark.gpu
def test_gpu_real32_example1_equals_reference_source(gpu_lib, pkg):
    g = gold("example1")
    snaps = {ii: g[f"u_{ii}"] for ii in (0, 1, 10, 50, 100)}
    ode = _example1(pkg, pkg.real32, g["times"], g["ic"], g["width"], 3, 3, 0, snaps)
    assert ode.fevals == int(g["fevals"])


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_gpu_real32_example1_k_order_sweep_equals_reference_source(gpu_lib, pkg, k, order):
    g = gold("example1_sweep")
    w = _grid32(-5.0, 5.0, 100)[2]
    ode = _example1(pkg, pkg.real32, g[f"t_k{k}_o{order}"], g["ic"], w, k, order, 0, {}, final=g[f"u_k{k}_o{order}"])
    assert ode.fevals == int(g[f"fevals_k{k}_o{order}"])


@pytest.mark.gpu
def test_gpu_real32_example1_lax_friedrichs_equals_reference_source(gpu_lib, pkg):
    g = gold("example1_lf")
    snaps = {ii: g[f"u_{ii}"] for ii in (0, 10, 20)}
    _example1(pkg, pkg.real32, g["times"], g["ic"], g["width"], 3, 3, 1, snaps)


@pytest.mark.gpu
def test_gpu_real32_example2_equals_reference_source(gpu_lib, pkg):
    _example2(pkg, pkg.real32, gold("example2_40"), 40, 40, 5e-3, 5.0)


@pytest.mark.gpu
def test_gpu_real32_general_path_equals_reference_source(gpu_lib, pkg):
    _example2(pkg, pkg.real32, gold("example2_growth"), 24, 18, 2.5e-4, 0.5, growth=True)
