"""The Fortran programs of fortran/examples/ through the shim into libhrweno_b200.so, as fast as it can be done (no pytest,
no torch): one line per program.  Used to get the GPU leg of tests/test_zzzz_gpu_fortran_shim_exec.py onto a B200 inside a
few seconds of remaining GPU budget; the pytest file is the real test."""
import ctypes
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import shim_exec  # noqa: E402

lib = ctypes.CDLL(os.path.join(ROOT, "hr-weno_b200", "lib", "libhrweno_b200.so"))
names = sys.argv[1:] or ["burgers_fused", "pbe2d_fused", "pbe2d_growth_fused", "burgers_host_rhs"]
for name in names:
    fixture, snaps, must = shim_exec.OWN_PROGRAMS[name]
    ns, P = shim_exec.run_own_program(lib, name)
    shim_exec.assert_history_equals_fixture(ns, fixture, snaps)
    assert must <= set(P.interop.calls)
    print(f"{name}: bit-identical to {fixture} at outputs {snaps}, fevals {int(ns['nfev'])}  (t = {time.time() - t0:.1f} s)", flush=True)
