"""Multi-GPU host logic on CPU: the deterministic drop partition and the
results-only gather, exercised with a world_size-2 gloo group."""
import os
import sys

import numpy as np
import pytest

from axctdprocessor_b200 import batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_a_balanced_cover():
    rng = np.random.default_rng(0)
    sizes = [int(x) for x in rng.choice([31752000, 34560000], 1024)]
    for w in (1, 2, 4, 8):
        parts = batch.partition_drops(sizes, w)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(1024))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)
    assert batch.partition_drops([], 4) == [[], [], [], []]
    assert batch.partition_drops([5, 1], 4) == [[0], [1], [], []]


def test_wave_plan_respects_budget():
    sizes = [10, 10, 10, 50, 10]
    waves = batch.plan_waves(sizes, budget_bytes=240, bytes_per_sample=8.0)
    assert [i for w in waves for i in w] == list(range(5))
    assert all(sum(sizes[i] for i in w) * 8 <= 240 or len(w) == 1 for w in waves)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from axctdprocessor_b200 import batch as B
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sizes = [1000 + 37 * i for i in range(21)]
    mine = B.partition_drops(sizes, world)[rank]
    local = {i: ("drop", i, sizes[i] * 2) for i in mine}        # stands in for the per-drop summaries
    merged = B.gather_results(local, world, rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, sorted(merged), mine))


def test_two_rank_gloo_gather_covers_every_drop():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = sorted(got)
    assert a[1] == b[1] == list(range(21))
    assert sorted(a[2] + b[2]) == list(range(21)) and not set(a[2]) & set(b[2])
