"""Slab domain decomposition across GPUs, one process per GPU.

Host-side plumbing only: which cells a rank owns, and the one-time exchange of the CUDA IPC
handles of the halo mailboxes through torch.distributed (any backend; gloo on CPU in the tests).
The per-stage halo traffic itself never goes through here: the CUDA library stores the k
boundary cells straight into the neighbour's mailbox over NVLink (csrc/halo.cu).
"""
from __future__ import annotations


def partition(n_global, nranks, rank):
    """contiguous slabs along the decomposed axis -> (global_offset, n_local); remainder to the first ranks"""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_global), int(nranks))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def neighbours(nranks, rank):
    """(left, right) rank numbers or None at a physical boundary"""
    return (rank - 1 if rank > 0 else None, rank + 1 if rank < nranks - 1 else None)


def connect(fv, rank, nranks, all_gather_object):
    """Export this rank's halo mailbox, gather every rank's handle and import the two neighbours'.

    `all_gather_object(obj) -> list` is the only communication primitive needed, e.g.
    ``lambda o: _gather(o)`` around torch.distributed.all_gather_object.
    """
    if nranks <= 1:
        return
    handles = all_gather_object(fv.export_halo())
    left, right = neighbours(nranks, rank)
    fv.import_halo(handles[left] if left is not None else None, handles[right] if right is not None else None)


def torch_all_gather(world):
    import torch.distributed as dist

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    return gather


def reduce_alpha(local_max, nranks):
    """max over ranks of a one-element float64 tensor (NCCL for CUDA tensors, gloo for CPU ones) -> float"""
    import torch.distributed as dist

    if nranks > 1:
        dist.all_reduce(local_max, op=dist.ReduceOp.MAX)
    return float(local_max.item())


def global_max_wavespeed(fv, u_dev, nranks, device_scalar=None):
    """Extension (BASELINE north star; the reference has a caller-supplied alpha and no reduction): the Lax-Friedrichs
    alpha = max |f'(u)| over ALL slabs.  The local maximum is reduced on the GPU (csrc/fv.cu: max_abs_kernel), the ranks
    combine their scalars with ONE max all-reduce on the device (torch.distributed, NCCL on GPUs) and the result is
    installed with fv.set_alpha().  `u_dev` is a torch CUDA tensor holding this rank's dense state."""
    import torch

    out = device_scalar if device_scalar is not None else torch.zeros(1, dtype=torch.float64, device=u_dev.device)
    fv.max_wavespeed_dev(u_dev.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    alpha = reduce_alpha(out, nranks)
    fv.set_alpha(alpha)
    return alpha
