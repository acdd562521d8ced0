"""CPU tier: the N>1 host logic on world_size 2 over gloo -- ownership of recordings is a partition, and the
job throughput is all units over the slowest rank's time (the only cross-rank exchange bench.py does)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import btk_b200  # noqa: F401
from btk_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rec, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard(n_rec, rank, world)
    own = torch.zeros(n_rec, dtype=torch.int64)
    own[mine] = 1
    dist.all_reduce(own)                                  # every recording owned exactly once
    units = torch.tensor([float(len(mine)) * 480.0])      # channel-seconds this rank processed
    t = torch.tensor([0.010 * (rank + 1)])                # pretend rank 1 is slower
    tot = units.clone()
    dist.all_reduce(tot)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, own.tolist(), float(tot / tmax)))
    dist.destroy_process_group()


def test_partition_and_max_over_ranks_gloo():
    world, n_rec = 2, 7
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rec, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, own, rate in res:
        assert own == [1] * n_rec
        assert abs(rate - sharding.job_throughput([4 * 480.0, 3 * 480.0], [0.010, 0.020])) < 1e-6


def test_greedy_balances_lengths():
    owner = sharding.greedy_by_length([10, 1, 1, 1, 7, 3], 2)
    loads = [sum(l for l, o in zip([10, 1, 1, 1, 7, 3], owner) if o == r) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 1
    assert sorted(set(sharding.shard(5, 0, 2) + sharding.shard(5, 1, 2))) == list(range(5))
