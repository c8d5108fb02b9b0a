"""The Fortran shim EXECUTED against libhrweno_b200.so on the GPU (see tests/test_fortran_shim_exec.py for the method and for
the same runs on the CPU stand-in of the C ABI).  The reference tree does not exist on the GPU box, so these tests run the
self-contained programs of fortran/examples/ -- the fused 1D and 2D operators and a caller's own `pure` Fortran rhs -- through
fortran/hrweno_b200_shim.f90 into the CUDA library, and hold their outputs, bit for bit at every output the fixtures hold, to
the reference's source executed for the same problems (tests/golden/ref_exec_*.npz).  Strict mode: bit-identical is the bar."""
import ctypes
import os

import pytest

import shim_exec
from shim_exec import OWN_PROGRAMS, OWN_PROGRAMS_REAL32, f90py

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda_abi(gpu_lib, pkg):
    """a ctypes handle of its own on the product library: the interop layer sets argtypes from the FORTRAN declarations, which
    must not overwrite the prototypes the package put on its own handle"""
    path = os.path.join(os.path.dirname(pkg.__file__), "lib", "libhrweno_b200.so")
    lib = ctypes.CDLL(path)
    assert lib is not gpu_lib
    return lib


@pytest.mark.parametrize("name", sorted(OWN_PROGRAMS))
def test_own_fortran_program_on_the_shim_and_the_cuda_library(cuda_abi, name):
    fixture, snaps, must_call = OWN_PROGRAMS[name]
    ns, P = shim_exec.run_own_program(cuda_abi, name)
    shim_exec.assert_history_equals_fixture(ns, fixture, snaps)
    assert must_call <= set(P.interop.calls)


def test_error_stops_through_the_shim_carry_the_reference_messages(cuda_abi):
    P = shim_exec.program_on_shim(cuda_abi, [])
    ns = P.build()
    with pytest.raises(f90py.FortranStop, match="Invalid input 'ncells'. Valid range: ncells > 0."):  # weno.f90:75
        ns["weno"](0)
    with pytest.raises(f90py.FortranStop, match="Invalid input 'k'. Valid range: 1 <= k <= 3."):  # weno.f90:84
        ns["weno"](10, 4)
    with pytest.raises(f90py.FortranStop, match="Invalid input 'order' in 'rktvd'"):  # tvdode.f90:89
        ns["rktvd"](lambda t, u, udot: None, 10, 4)


def test_weno_type_of_the_shim_on_the_gpu(cuda_abi, ref):
    """type(weno) of the shim: the constructor with xedges fills the public `cnu` component (weno.f90:41,100-112); reconstruct
    of contiguous and strided actuals (example2:98,107) equals the oracle bit for bit; destroy() releases the handle"""
    shim_exec.check_weno_type(cuda_abi, ref)


def test_multi_gpu_fortran_program_on_every_visible_device(cuda_abi, gpu_lib, ref, pkg):
    """fortran/examples/burgers_multi_gpu.f90: ONE Fortran process, hrweno_mgpu_create(..., 0, c_null_ptr) = all visible GPUs,
    state resident between outputs; slabs + halos inside the library; bit-identical to the single-domain oracle"""
    assert shim_exec.check_multi_gpu_program(cuda_abi, ref, pkg) == gpu_lib.hrweno_device_count()


def test_time_dependent_growth_fortran_program_on_the_gpu(cuda_abi, ref, pkg):
    """fortran/examples/pbe2d_growth_time_factor.f90: the library calls the program's bind(c) g(t) back at every stage time"""
    shim_exec.check_time_factor_program(cuda_abi, ref, pkg)


@pytest.mark.parametrize("name", sorted(OWN_PROGRAMS_REAL32))
def test_own_fortran_program_on_the_real32_build_of_the_shim_and_the_cuda_library(cuda_abi, name):
    """the shim compiled with -DREAL32 binds hrweno_*_f32: the same programs in binary32 against the REAL32 executed-source fixtures"""
    fixture, snaps, must_call = OWN_PROGRAMS_REAL32[name]
    ns, P = shim_exec.run_own_program(cuda_abi, name, real32=True)
    shim_exec.assert_history_equals_fixture(ns, fixture, snaps)
    assert must_call <= set(P.interop.calls)


def test_weno_type_of_the_real32_build_of_the_shim_on_the_gpu(cuda_abi):
    from oracle import ref32

    shim_exec.check_weno_type_real32(cuda_abi, ref32)


def test_device_integrand_fortran_program_on_the_gpu(cuda_abi, ref, pkg):
    """fortran/examples/burgers_device_rhs.f90: rktvd_dev / mstvd_dev, the bind(c) integrand enqueues hrweno_fv_rhs_dev on the
    integrator's stream; equals the fused integrators (and the oracle) bit for bit"""
    shim_exec.check_device_integrand_program(cuda_abi, ref, pkg)
