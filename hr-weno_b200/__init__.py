"""hrweno_b200 -- host-side mirror of HR-WENO's module interface over libhrweno_b200.so.

The directory is called ``hr-weno_b200`` (not importable by name); load it through
``__graft_entry__.load_package()`` which registers it as ``hrweno_b200``.

Module names follow the reference's Fortran modules:
    hrweno_weno    (src/hrweno_weno.f90)    -> weno, c1, c2, c3
    hrweno_fluxes  (src/hrweno_fluxes.f90)  -> godunov, lax_friedrichs
    hrweno_tvdode  (src/hrweno_tvdode.f90)  -> rktvd, mstvd
    hrweno_grids   (src/hrweno_grids.f90)   -> grid1 (host side, like the reference)
    fv                                      -> the example `rhs` as a fused device operator
    real32                                  -> the REAL32 build of the path (hrweno_kinds.F90:9-17): weno, FV, rktvd, mstvd on float32
    mgpu                                    -> the same over the GPUs of one box from ONE process (hrweno_mgpu_*)
All compute goes through the C ABI into CUDA kernels; nothing here computes on the CPU
except grid set-up, which the north star keeps on the host.
"""
from . import _abi
from ._abi import HrwenoError, lib
from . import hrweno_grids, hrweno_weno, hrweno_fluxes, hrweno_tvdode, fv, slab, mgpu, real32

__all__ = ["_abi", "HrwenoError", "lib", "hrweno_grids", "hrweno_weno", "hrweno_fluxes", "hrweno_tvdode", "fv", "slab", "mgpu", "real32"]
