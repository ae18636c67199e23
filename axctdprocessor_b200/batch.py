"""Batches of independent drops and their sharding across the GPUs of one box.

The path shards by drop: every drop is decoded end to end on
one GPU, there is no data-path collective, and only the small per-drop results
are gathered on the host.  Within a drop the chunk chain is sequential, so a
single drop does not scale across GPUs ("replicas only" for one recording).
"""
from __future__ import annotations

from typing import Callable, Sequence


def partition_drops(sizes: Sequence[int], world_size: int) -> list:
    """Longest-processing-time greedy on sample counts: returns, per rank, the
    (sorted) indices of the drops it decodes.  Deterministic, identical on every
    rank, so no exchange is needed to agree on the partition."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(x) for x in out]


def plan_waves(sizes: Sequence[int], budget_bytes: int, bytes_per_sample: float = 8.0) -> list:
    """Split a rank's drops into waves whose device footprint (PCM + crossing /
    bit / frame scratch, ~8 B per sample) stays under budget_bytes."""
    waves, cur, used = [], [], 0.0
    for i, n in enumerate(sizes):
        need = n * bytes_per_sample
        if cur and used + need > budget_bytes:
            waves.append(cur)
            cur, used = [], 0.0
        cur.append(i)
        used += need
    if cur:
        waves.append(cur)
    return waves


def process_drops(eng, pcm_list, fs_list, settings=None, triggerrange=None, budget_bytes=64 << 30):
    """Decode a list of mono int16 recordings on one GPU (waves as needed)."""
    cfgs = [eng.config(fs, settings=settings, triggerrange=triggerrange) for fs in fs_list]
    results = [None] * len(pcm_list)
    for wave in plan_waves([len(p) for p in pcm_list], budget_bytes):
        out = eng.process([pcm_list[i] for i in wave], [cfgs[i] for i in wave])
        for i, r in zip(wave, out):
            results[i] = r
    return results


def gather_results(local: dict, world_size: int, rank: int, all_gather_object: Callable | None = None) -> dict:
    """Host-side gather of {drop index: small result} from every rank (results only;
    audio never leaves the GPU that decoded it)."""
    if world_size == 1:
        return dict(local)
    if all_gather_object is None:
        import torch.distributed as dist
        bucket = [None] * world_size
        dist.all_gather_object(bucket, local)
    else:
        bucket = all_gather_object(local)
    merged = {}
    for part in bucket:
        merged.update(part)
    return merged
