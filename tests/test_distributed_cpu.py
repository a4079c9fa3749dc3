"""CPU, world_size 2 over gloo: the multi-process host logic -- replica
sharding and the one collective (episode-metric all-reduce)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytsc_b200.env import reduce_episode_metrics, shard_replicas


def test_shard_replicas_partitions_exactly():
    for total in (1, 7, 4096, 16384):
        for world in (1, 2, 3, 8):
            got = [shard_replicas(total, world, r) for r in range(world)]
            assert sum(c for _, c in got) == total
            assert got[0][0] == 0 and all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_replicas(10, world, rank)
    # every replica b contributes ATT = 100 + b, 3 finished, 2 running, reward -b, queue b, over 4 steps
    ids = torch.arange(first, first + count, dtype=torch.float64)
    vec = torch.stack([(100 + ids).sum(), 3.0 * count * torch.ones(()), 2.0 * count * torch.ones(()),
                       4 * (-ids).sum(), 4 * ids.sum(), torch.tensor(4.0 * count), torch.tensor(float(count))]).double()
    m = reduce_episode_metrics(vec)
    # the timing convention of bench.py: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((m, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_episode_metric_all_reduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    m, tmax = q.get()
    assert m["replicas"] == 10
    assert m["average_travel_time"] == pytest.approx(104.5)
    assert m["finished_vehicles"] == 30 and m["running_vehicles"] == 20
    assert m["mean_global_reward"] == pytest.approx(-4.5) and m["mean_n_queued"] == pytest.approx(4.5)
    assert tmax == 2.0


def test_reduce_without_process_group_is_local():
    vec = torch.tensor([200.0, 6, 4, -8, 8, 8, 2], dtype=torch.float64)
    m = reduce_episode_metrics(vec)
    assert m == {"average_travel_time": 100.0, "finished_vehicles": 6.0, "running_vehicles": 4.0,
                 "mean_global_reward": -1.0, "mean_n_queued": 1.0, "replicas": 2}
