"""CPU tests of the multi-rank plumbing with the gloo backend, world_size 2 (the step itself has no collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wfcrl_b200.dist import gather_episode_stats, gather_episode_sums, shard_range


def test_shard_range_partitions_exactly():
    for n, w in ((65536, 8), (10, 3), (7, 8), (8192, 1)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    returns = torch.arange(lo, hi, dtype=torch.float64)  # the "episode return" of global env i is i
    lengths = torch.full((hi - lo,), 99)
    stats = gather_episode_stats(returns, lengths)
    if rank == 0:
        out.put(stats)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_episode_stats_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert stats["episodes"] == 10 and stats["world_size"] == 2
    assert abs(stats["return_mean"] - 4.5) < 1e-12 and abs(stats["length_mean"] - 99) < 1e-12
    assert abs(stats["return_std"] - (sum((i - 4.5) ** 2 for i in range(10)) / 10) ** 0.5) < 1e-12


def _worker_sums(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    r = torch.arange(lo, hi, dtype=torch.float64)  # finished-episode returns of this rank's envs, one episode each
    stats = gather_episode_sums(float(r.sum()), float((r * r).sum()), float(hi - lo), 99.0 * (hi - lo), "cpu")
    if rank == 0:
        out.put(stats)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_episode_sums_gloo_world2_equals_stats_of_the_whole_job():
    """What VecWindFarmEnv.episode_statistics does with the per-env sums the step kernels keep: four numbers per rank."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_sums, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert stats["episodes"] == 10 and stats["world_size"] == 2
    assert abs(stats["return_mean"] - 4.5) < 1e-12 and abs(stats["length_mean"] - 99) < 1e-12
    assert abs(stats["return_std"] - (sum((i - 4.5) ** 2 for i in range(10)) / 10) ** 0.5) < 1e-12
    empty = gather_episode_sums(0.0, 0.0, 0.0, 0.0, "cpu")  # no finished episode yet: NaNs, not a division error
    assert empty["episodes"] == 0 and empty["return_mean"] != empty["return_mean"]


def test_single_process_stats():
    stats = gather_episode_stats(torch.tensor([1.0, 3.0]), torch.tensor([5, 7]))
    assert stats == {"episodes": 2.0, "return_mean": 2.0, "return_std": 1.0, "length_mean": 6.0, "world_size": 1}


def test_numa_binding_helper_is_best_effort():
    from wfcrl_b200.dist import _parse_cpulist, bind_to_gpu_numa_node

    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and _parse_cpulist("") == []
    import os

    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa_node(0, sysfs="/nonexistent")   # no GPU / no sysfs entry: reports, never raises
    assert info["bound"] is False and "error" in info and os.sched_getaffinity(0) == before
