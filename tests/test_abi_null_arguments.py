"""The C ABI is the drop-in boundary: a caller's mistake must come back as a status, never as a crash.  Every entry point of
include/hrweno_b200.h is called with null handles / null pointers / zero sizes (in one child process, so that a segmentation
fault shows up as that process's exit status with the name of the call that was being made).  No GPU needed: argument
validation precedes any device work."""
import subprocess
import sys

from conftest import ROOT

CHILD = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
import __graft_entry__ as g
pkg = g.load_package()
lib = pkg.lib()
for name in sorted(pkg._abi.PROTOTYPES):
    res, args = pkg._abi.PROTOTYPES[name]
    vals = []
    for a in args:
        if a in (C.c_int, C.c_int32, C.c_int64):
            vals.append(0)
        elif a in (C.c_double, C.c_float):
            vals.append(0.0)
        elif isinstance(a, type) and issubclass(a, C._CFuncPtr):
            vals.append(a())
        else:
            vals.append(None)
    print("CALL", name, flush=True)
    r = getattr(lib, name)(*vals)
    print("RET", name, repr(r), flush=True)
print("DONE", flush=True)
"""


def test_every_entry_point_survives_null_arguments(pkg):
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, timeout=300)
    calls = [ln.split()[1] for ln in p.stdout.splitlines() if ln.startswith("CALL")]
    assert p.returncode == 0 and p.stdout.rstrip().endswith("DONE"), f"crashed inside {calls[-1] if calls else '?'}: rc={p.returncode}\n{p.stderr[-2000:]}"
    rets = {ln.split()[1]: ln.split(None, 2)[2] for ln in p.stdout.splitlines() if ln.startswith("RET")}
    assert set(rets) == set(pkg._abi.PROTOTYPES) and len(rets) >= 85
    import ctypes as C

    for name, (res, args) in pkg._abi.PROTOTYPES.items():
        if res is C.c_int and args and name not in ("hrweno_abi_version", "hrweno_device_count", "hrweno_mgpu_ngpus", "hrweno_ode_order"):
            if name.endswith("_istate"):
                assert rets[name] == "-1", name  # tvdode.f90:294: istate = -1 is the error state
            else:
                assert rets[name] == str(pkg._abi.EINVAL), f"{name} returned {rets[name]} for null arguments"
    assert rets["hrweno_godunov"] == "nan" and rets["hrweno_lax_friedrichs"] == "nan"
