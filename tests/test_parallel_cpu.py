"""World-size-2 gloo tests (CPU) of the only cross-rank step of the path: the per-bin reduction of observable accumulators
and control counters that replaces ALF's MPI_REDUCE calls (Prog/observables_mod.F90:425-438, Prog/control_mod.F90:397-452),
plus the chain -> rank sharding and the seeds-file scatter of Prog/Set_random_mod.F90:79-84."""
import os
import socket

import numpy as np
import pytest

from alf_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # bins: rank r contributes (r + 1) * arange
        obs = (rank + 1) * np.arange(24, dtype=np.float64)
        lat = (rank + 1) * (np.arange(12).reshape(3, 4) + 1j * np.ones((3, 4)))      # a lattice accumulator (complex, like Obs_Latt)
        both = parallel.reduce_host_bins({"obs": obs, "Green_eq": lat})
        red = None if both is None else both["obs"]
        if both is not None:
            assert np.allclose(both["Green_eq"], 3.0 * (np.arange(12).reshape(3, 4) + 1j * np.ones((3, 4))))
        # control: sums except the maxima entries 1, 3, 5, 11, 12
        ctl = np.arange(16, dtype=np.float64) + 100.0 * rank
        rc = parallel.reduce_control(ctl)
        chains = parallel.shard_chains(7, world, rank)
        q.put((rank, None if red is None else red.tolist(), rc.tolist() if rank == 0 else None, chains))
    finally:
        dist.destroy_process_group()


def test_shard_and_seed_order():
    assert parallel.shard_chains(7, 2, 0) == [0, 2, 4, 6]
    assert parallel.shard_chains(7, 2, 1) == [1, 3, 5]
    allc = sorted(sum((parallel.shard_chains(10, 4, r) for r in range(4)), []))
    assert allc == list(range(10))
    seeds = [11, 22, 33, 44]
    # rank C-1 takes line 1, rank 0 takes line C
    assert [parallel.seed_for_rank_file_order(seeds, 4, r) for r in range(4)] == [44, 33, 22, 11]


def test_reduce_is_noop_without_process_group():
    obs = {"obs": np.arange(5.0)}
    assert parallel.reduce_host_bins(obs) is obs
    out = parallel.reduce_control(np.arange(16.0))
    assert np.array_equal(out, np.arange(16.0))


@pytest.mark.timeout(120)
def test_bin_and_control_reduction_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port(); world = 2
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = {}
    for _ in range(world):
        r, red, rc, chains = q.get(timeout=100)
        res[r] = (red, rc, chains)
    [p.join(30) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    red0, rc0, ch0 = res[0]
    assert res[1][0] is None                                   # only the destination rank holds the reduced bin
    assert np.allclose(red0, 3.0 * np.arange(24))              # (1 + 2) * arange
    exp = 2.0 * np.arange(16) + 100.0
    for i in (1, 3, 5, 11, 12):
        exp[i] = i + 100.0                                     # maxima over ranks
    assert np.allclose(rc0, exp)
    assert ch0 == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]
