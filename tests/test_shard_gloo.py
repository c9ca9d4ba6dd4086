"""N>1 host logic on CPU: world_size-2 gloo processes exercise trajectory sharding, the
max-over-ranks timing reduction and the statistics gather used by bench.py / the evaluator."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tante_b200.shard import gather_counts, max_over_ranks, micro_batches, trajectory_shard


def test_shards_are_a_disjoint_contiguous_cover():
    for n in (0, 1, 7, 8, 64, 513):
        for w in (1, 2, 3, 4, 8):
            got = []
            for r in range(w):
                a, b = trajectory_shard(n, w, r)
                assert 0 <= a <= b <= n
                got += list(range(a, b))
            assert got == list(range(n))
            sizes = [trajectory_shard(n, w, r)[1] - trajectory_shard(n, w, r)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert micro_batches(3, 20, 8) == [(3, 11), (11, 19), (19, 20)]
    with pytest.raises(ValueError):
        trajectory_shard(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = trajectory_shard(n_total, world, rank)
    # stand-in for the per-trajectory result of a rollout: the step count each trajectory needed
    local_steps = torch.arange(a, b, dtype=torch.int64) % 3 + 2
    ms = max_over_ranks(10.0 + 5.0 * rank)                 # rank 1 is the slow one
    parts = gather_counts(local_steps)
    total = torch.tensor([float(b - a)])
    dist.all_reduce(total)                                 # whole-job unit count (outside timed regions)
    q.put((rank, a, b, ms, [p.tolist() for p in parts], float(total.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_reduction():
    world, n_total = 2, 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, ms0, parts0, tot0), (r1, a1, b1, ms1, parts1, tot1) = res
    assert (a0, b0, a1, b1) == (0, 7, 7, 13)
    assert ms0 == ms1 == 15.0                              # max over ranks, identical everywhere
    assert tot0 == tot1 == float(n_total)
    flat = sum(parts0, [])
    assert flat == [(i % 3) + 2 for i in range(n_total)] and parts0 == parts1
