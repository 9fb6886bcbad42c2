"""Chains shard over the GPUs of one box (one process per GPU); the only exchange is the per-bin reduction of the
observable accumulators and of the control counters, which replaces ALF's MPI_REDUCE calls
(Prog/observables_mod.F90:425-438,648-653,828-834; Prog/control_mod.F90:397-452) by NCCL over NVLink.

On the GPU the reduction itself is done by the C-ABI (alf_b200_comm_init / alf_b200_reduce_bins / alf_b200_reduce_control: NCCL, in place on the
handle's device buffers), so a Fortran or C++ host needs no Python; torch.distributed only carries the 128-byte communicator id.  On CPU
(world_size-2 gloo tests) reduce_host_bins / reduce_control reduce host copies of the same accumulators.
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


def init_comm(handles, world: int, rank: int):
    """One NCCL communicator per handle behind the C-ABI (alf_b200_comm_init).  The 128-byte unique ids are created on rank 0 and
    handed to the other ranks through the process group that launched the job (ALF: MPI_BCAST) -- plumbing only."""
    import torch.distributed as dist
    if world <= 1:
        return
    ids = [type(handles[0]).comm_unique_id() for _ in handles] if rank == 0 else [None] * len(handles)
    dist.broadcast_object_list(ids, src=0)
    for g, uid in zip(handles, ids):
        g.comm_init(world, rank, uid)


def reduce_bins(handles, world: int, dst: int = 0, rank: int = 0):
    """Per-bin reduction (C1 of SURVEY 2.5) of EVERY observable accumulator -- scalars, equal-time and time-displaced lattice observables with
    backgrounds and counters -- of the handles of this rank: the device buffers are SUM-reduced in place over NCCL by the C-ABI
    (alf_b200_reduce_bins), then the handles of the destination rank are added up on the host.  Returns on `dst` a dict
    {"obs", "eq": (acc, bg, n, sign) or None, "tau": ... or None}; None elsewhere."""
    for g in handles:
        if world > 1:
            g.reduce_bins(dst)
    if world > 1 and rank != dst:
        return None
    out = {"obs": sum(np.asarray(g.obs()) for g in handles), "eq": None, "tau": None}
    for key, getter, flag in (("eq", "obs_eq", "_obs_eq_on"), ("tau", "obs_tau", "_obs_tau_on")):
        parts = [getattr(g, getter)() for g in handles if getattr(g, flag, False)]
        if parts:
            out[key] = tuple(sum(p[i] for p in parts) for i in range(4))
    return out


def reduce_host_bins(arrays, dst: int = 0):
    """Host-side twin for process groups without GPUs (world-size-2 gloo tests on CPU): SUM-reduces a dict of numpy arrays (the same
    accumulators, already on the host) to rank `dst`."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return arrays
    out = {}
    for k in sorted(arrays):
        a = np.asarray(arrays[k]); cplx = np.iscomplexobj(a)
        t = torch.as_tensor(np.ascontiguousarray(a.view(np.float64) if cplx else a.astype(np.float64))).clone()
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
        r = t.numpy()
        out[k] = r.view(np.complex128).reshape(a.shape) if cplx else r.reshape(a.shape)
    return out if dist.get_rank() == dst else None
