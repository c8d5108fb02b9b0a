"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the keys the driver
reads, and the B200 arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "WENO5+RK3 cell-updates/s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_parity_check_head_cut_is_invisible():
    """bench.py's parity_check integrates only the head of the domain on the oracle: the cut must not reach the compared
    cells (k = 3 cells per stage, 9 per RK3 step), i.e. the head of a whole-domain run equals the cut run bit for bit"""
    import numpy as np

    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    import bench

    pkg, ref = graft.load_package(), graft.load_oracle()
    n, calls = 20000, (2, 5)
    m, mh = bench.parity_head(n, 1, sum(calls), cells=4096)
    assert (m, mh) == (4096, 4096 + 9 * 7 + 32)
    assert bench.parity_head(3000, 1, 7, cells=4096) == (3000, 3000)  # small run: whole domain
    assert bench.parity_head(100, 2, 7, cells=4096) == (5, 100)      # slab barely larger than the halo
    assert bench.parity_head(90, 2, 7, cells=4096) == (0, 0)
    u0 = bench.make_ic(n)
    dt = 0.1 * 10.0 / n
    g = pkg.hrweno_grids.grid1().linear(bench.XMIN, bench.XMAX, n)
    full = ref.rktvd(ref.FV(pkg.fv.make_desc(n, k=3, eps=1e-6, width=[g.width])), 3)
    u, t = u0.copy(), 0.0
    for k in calls:
        tt = t
        for _ in range(k - 1):
            tt = tt + dt
        t = full.integrate(u, t, tt, dt)
    res = bench.parity_check_leg(pkg, u0[:mh].copy(), u[:m].copy(), n, dt, calls)
    assert res["bit_identical"] and res["ok"] and res["cells"] == m and res["steps"] == 7
    bad = u[:m].copy()
    bad[17] += 1e-9
    assert not bench.parity_check_leg(pkg, u0[:mh].copy(), bad, n, dt, calls)["ok"]
