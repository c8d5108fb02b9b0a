"""CPU, world_size 2 (gloo): the host-side logic of the N > 1 path -- slab partition, per-rank synthetic
initial data, neighbour mapping and the one-time handle exchange.  The per-stage halo traffic is device code
(csrc/halo.cu) and is covered on real GPUs by tests/mgpu_worker.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeFV:
    """stands in for fv.FV: records what slab.connect hands to import_halo"""

    def __init__(self, rank):
        self.rank, self.imported = rank, None

    def export_halo(self):
        return bytes([self.rank]) * 64

    def import_halo(self, left, right):
        self.imported = (left, right)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = graft.load_package()
    out = {}
    # neighbour wiring through the real connect() with the real all_gather_object
    fv = _FakeFV(rank)
    pkg.slab.connect(fv, rank, world, pkg.slab.torch_all_gather(world))
    out["imported"] = fv.imported
    # weak-scaling initial data: each rank's slab of the global ramp equals the slice of the global array
    n = 1000
    slab = bench.make_ic(n, rank * n, n * world)
    parts = pkg.slab.torch_all_gather(world)(slab)
    out["ic_edges_match"] = bool(np.array_equal(np.concatenate(parts)[rank * n:(rank + 1) * n], slab))
    # oracle per slab with k halo cells from the neighbour (sent with gloo) equals the global oracle rhs
    ref = graft.load_oracle()
    k, nglob = 3, 400
    off, nl = pkg.slab.partition(nglob, world, rank)
    g = pkg.hrweno_grids.grid1().linear(-5.0, 5.0, nglob)
    u = np.clip(1.0 - 0.25 * (g.center + 4.0), -0.5, 1.0) + 1e-3 * np.random.default_rng(0).standard_normal(nglob)
    full = ref.FV(pkg.fv.make_desc(nglob, k=k, width=[g.width])).rhs(0.0, u)
    mine = u[off:off + nl].copy()
    everyone = pkg.slab.torch_all_gather(world)(mine)
    left, right = pkg.slab.neighbours(world, rank)
    lo = everyone[left][-k:] if left is not None else np.empty(0)
    hi = everyone[right][:k] if right is not None else np.empty(0)
    ext = np.concatenate([lo, mine, hi])
    wext = g.width[off - len(lo):off + nl + len(hi)]
    part = ref.FV(pkg.fv.make_desc(len(ext), k=k, width=[wext])).rhs(0.0, ext)[len(lo):len(lo) + nl]
    # interior faces only: the artificial ends of the extended slab use the copy rule, k cells away from our cells
    out["halo_k_bitwise"] = bool(np.array_equal(part, full[off:off + nl]))
    # the alpha all-reduce of the adaptive Lax-Friedrichs extension: max over ranks of the local max |f'(u)|
    import torch

    out["alpha"] = pkg.slab.reduce_alpha(torch.tensor([np.max(np.abs(mine))], dtype=torch.float64), world)
    out["alpha_want"] = float(np.max(np.abs(u)))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_domain(pkg):
    for nglob, world in [(10, 3), (1 << 28, 8), (7, 7), (1003, 4)]:
        spans = [pkg.slab.partition(nglob, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == nglob
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1
    assert pkg.slab.neighbours(4, 0) == (None, 1) and pkg.slab.neighbours(4, 3) == (2, None)


@pytest.mark.timeout(300)
def test_world_size_2_gloo(pkg):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["imported"] == (None, bytes([1]) * 64)
    assert res[1]["imported"] == (bytes([0]) * 64, None)
    for r in range(world):
        assert res[r]["ic_edges_match"] and res[r]["halo_k_bitwise"]
        assert res[r]["alpha"] == res[r]["alpha_want"]
