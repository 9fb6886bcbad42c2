"""Chains shard over the GPUs of one box (one process per GPU); the only exchange is the per-bin reduction of the
observable accumulators and of the control counters, which replaces ALF's MPI_REDUCE calls
(Prog/observables_mod.F90:425-438,648-653,828-834; Prog/control_mod.F90:397-452) by NCCL over NVLink.

The device buffers of the C-ABI handle are wrapped zero-copy (``__cuda_array_interface__``) so NCCL reads them in place.
On CPU (tests, world_size-2 gloo) the same code path reduces host copies.
"""
from __future__ import annotations

import numpy as np


class _DevArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def shard_chains(n_chains_total: int, world: int, rank: int):
    """Chain c -> rank c mod world; returns the global chain ids owned by `rank`."""
    return list(range(rank, n_chains_total, world))


def seed_for_rank_file_order(seeds, n_ranks: int, r: int):
    """Set_random_mod.F90:79-84: with C ranks, rank r (0-based) takes line C-r (1-based) of the seeds file."""
    return seeds[n_ranks - r - 1]


def reduce_sum(t, dst: int = 0):
    """SUM-reduce a tensor (device: NCCL, host: gloo) to rank `dst`; no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t


def reduce_control(ctl_vec, dst: int = 0):
    """Control_Print's 18 scalar MPI_REDUCEs packed into one SUM and one MAX reduction.
    ctl_vec: the 16 doubles of alf_b200_get_control; entries 1, 3, 5 (XMAXG, XMAXP, XMAX_tau), 11, 12 (flags) are maxima."""
    import torch
    import torch.distributed as dist
    v = torch.as_tensor(np.asarray(ctl_vec, dtype=np.float64)).clone()
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return v.numpy()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    s = v.to(dev).clone(); m = v.to(dev).clone()        # .to() aliases on CPU: the two reductions need distinct buffers
    dist.reduce(s, dst=dst, op=dist.ReduceOp.SUM); dist.reduce(m, dst=dst, op=dist.ReduceOp.MAX)
    out = s.cpu().numpy()
    for i in (1, 3, 5, 11, 12):
        out[i] = float(m[i])
    return out


def reduce_bins(g, obs_host, world: int, dst: int = 0):
    """Per-bin reduction of the observable accumulators of handle `g` to rank `dst` (C1 of SURVEY 2.5).
    With NCCL the handle's device buffer is reduced in place from HBM (no host staging); returns the host copy on `dst`."""
    import torch
    import torch.distributed as dist
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        return obs_host
    if dist.get_backend() == "nccl":
        ptr, n = g.obs_device_ptr()
        t = torch.as_tensor(_DevArray(ptr, n), device="cuda").clone()
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
        return t.cpu().numpy() if dist.get_rank() == dst else None
    t = torch.as_tensor(np.asarray(obs_host, dtype=np.float64)).clone()
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t.numpy() if dist.get_rank() == dst else None
