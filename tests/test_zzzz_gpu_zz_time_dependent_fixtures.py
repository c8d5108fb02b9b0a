"""GPU: t-dependent fluxes against the reference's source executed with g(t) = 1 + t/4 on its flux lines
(tests/golden/ref_exec_example1_tfactor.npz, ref_exec_example2_growth_tfactor.npz; CPU side:
tests/test_reference_source_exec.py::test_oracle_time_dependent_*).  Bar: bit-identical.  Written after the round's GPU budget
was spent, hence sorted after every test that has already run on a B200 (`pytest -x`)."""
import pytest

from test_reference_source_exec import _example1_tfactor, _example2, g_of_t, gold


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2, 3])
def test_gpu_time_dependent_flux_example1_equals_reference_source(gpu_lib, pkg, order):
    """example1 executed from source with flux = (v**2)/2*(1 + t/4): the fused stage with the separable time factor, g called on
    the host at the reference's stage times"""
    _example1_tfactor(pkg, pkg.fv.FV, lambda fv, o: pkg.hrweno_tvdode.rktvd(fv, 100, o), order)


@pytest.mark.gpu
def test_gpu_time_dependent_growth_example2_equals_reference_source(gpu_lib, pkg):
    _example2(pkg, lambda fv: pkg.hrweno_tvdode.mstvd(fv, 24 * 18), gold("example2_growth_tfactor"), 24, 18, 2.5e-4, 0.5, growth=True,
              mod=pkg.fv, time_fn=g_of_t)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2, 3])
def test_gpu_real32_time_dependent_flux_example1_equals_reference_source(gpu_lib, pkg, order):
    """the REAL32 entry points against example1 executed in real32 with the same patched flux line"""
    import test_reference_source_exec_real32 as r32

    r32._example1_tfactor32(pkg, pkg.real32, order)


@pytest.mark.gpu
def test_gpu_real32_time_dependent_growth_example2_equals_reference_source(gpu_lib, pkg):
    import test_reference_source_exec_real32 as r32

    r32._example2(pkg, pkg.real32, r32.gold("example2_growth_tfactor"), 24, 18, 2.5e-4, 0.5, growth=True, time_fn=r32.g_of_t32)
