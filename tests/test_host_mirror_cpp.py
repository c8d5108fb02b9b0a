"""CPU: the C++ host mirror (hr-weno_b200/host/hrweno.hpp -- the reference's module / type / procedure names over the C
ABI, incl. the one-process multi-GPU operator and the REAL32 types) compiles against include/hrweno_b200.h without
warnings and links against libhrweno_b200.so with every wrapper instantiated; started without a device it reports the ABI
version and exits (no compute call without a GPU)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_cpp_host_mirror_compiles_links_and_loads(pkg, tmp_path):
    exe = str(tmp_path / "host_mirror_link")
    lib_dir = os.path.join(ROOT, "hr-weno_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "hr-weno_b200", "host"),
           os.path.join(ROOT, "tests", "cpp", "host_mirror_link.cpp"), "-L", lib_dir, "-lhrweno_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith(f"abi {pkg._abi.ABI_VERSION}, devices"), r.stdout + r.stderr
