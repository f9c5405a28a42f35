"""Column sharding of the VB / Gibbs hot path across the GPUs of one box
(SURVEY.md section 8e): the N columns of V are split into contiguous blocks, one
per rank; the Beta-side quantities (f_ones, f_zeros, B1, B0) of a block live only
on its GPU; the Dirichlet-side statistic n_total (F x K) is replicated and is the
one thing exchanged each iteration (all-reduce, sum).  One process per GPU,
`torch.distributed` (NCCL over NVLink) is only the plumbing: the engine calls the
hook below from inside its iteration loop with a device pointer.
"""
from __future__ import annotations

import numpy as np


def shard_columns(N: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous column block [lo, hi) of rank `rank` out of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (N * rank) // world, (N * (rank + 1)) // world


def tensor_from_device_ptr(ptr: int, count: int, device):
    """fp64 torch view of `count` doubles at device address `ptr` (no copy)."""
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(a, device=device)


def init_nccl(device):
    """`init_process_group("nccl")` with the collective stream at high priority: the engine overlaps the
    exchange with its persistent cols-pass kernel, so the all-reduce's CTAs have to be placed first."""
    import torch.distributed as dist
    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:  # older torch: default priority still works, with less overlap
        opts = None
    if opts is not None:
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=device)


def make_allreduce_hook(device, group=None):
    """Hook for Engine.set_allreduce_hook: sums the buffer across ranks in place, ordered on the
    stream the engine passes (its main stream, or the side stream of an overlapped exchange)."""
    import torch
    import torch.distributed as dist

    def hook(ptr: int, count: int, stream: int):
        t = tensor_from_device_ptr(ptr, count, device)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
            dist.all_reduce(t, group=group)
    return hook


def init_library_comm(engine, rank: int, world: int, group=None):
    """Give the engine its own NCCL communicator (nbmf_comm_init_rank): rank 0 asks the library for a
    unique id and `torch.distributed` -- any backend -- only ships its 128 bytes.  From then on the
    library exchanges the F x K statistic itself; no hook is installed."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.tensor(list(engine.comm_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    engine.comm_init(bytes(t.cpu().tolist()), world, rank)


def combine_lowerbounds(terms3: np.ndarray, allreduce_sum) -> np.ndarray:
    """Global ELBO trace from per-rank terms (nbmf_vb_get_lb_terms):
    lb_i = sum_ranks(sum log D + (pH - qH)) + (pW - qW); the row term is replicated.
    `allreduce_sum(x)` must return the sum over ranks of the fp64 array x."""
    t = np.asarray(terms3, dtype=np.float64)
    local = t[:, 0] + t[:, 2]
    return allreduce_sum(local) + t[:, 1]
