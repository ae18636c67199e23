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


def bind_host_thread_to_gpu(device: int) -> bool:
    """Restrict the calling thread (and the threads it starts later) to the CPUs NVML reports as local to the GPU,
    so that pinned ingest buffers allocated afterwards sit on the NUMA node the GPU's PCIe link hangs off.  Returns
    False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:                                    # CUDA device order need not be NVML's (CUDA_VISIBLE_DEVICES): go by PCI address
            import torch
            pr = torch.cuda.get_device_properties(int(device))
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


class PipelinedDecoder:
    """Ingest pipeline for a stream of batches on one GPU.

    Two engines (two CUDA streams) take alternate batches: ``submit`` enqueues the host->device copy of a
    batch on its engine's stream and returns at once, ``collect`` runs the decode of the oldest submitted
    batch and returns its results, so the copy of batch i+1 overlaps the decode of batch i.  Host buffers
    should be pinned (e.g. ``torch.empty(..).pin_memory()``) for the copies to be asynchronous."""

    def __init__(self, device: int = 0, slots: int = 2, engine_options=None, engine_factory=None):
        from . import engine as _engine
        make = engine_factory or (lambda: _engine.Engine(device))      # (tests inject the host emulation here)
        self.engines = [make() for _ in range(slots)]
        for e in self.engines:
            for k, v in (engine_options or {}).items():
                e.set_option(k, v)
        self._batches = [None] * slots          # (key, Batch) cached per slot: same shapes reuse the allocation
        self._pending = []                      # slots in submission order
        self._next = 0

    def close(self):
        for kb in self._batches:
            if kb:
                kb[1].close()
        for e in self.engines:
            e.close()
        self._batches, self.engines = [], []

    def submit(self, host_ptrs, n_samples, fs_list, settings=None, triggerrange=None):
        """Enqueue one batch: host_ptrs[i] = address of n_samples[i] int16 samples at rate fs_list[i]."""
        slot = self._next
        self._next = (self._next + 1) % len(self.engines)
        if slot in self._pending:
            raise RuntimeError("collect() the oldest batch before submitting more than `slots` batches")
        eng = self.engines[slot]
        key = (tuple(int(x) for x in n_samples), tuple(float(f) for f in fs_list), repr(settings), repr(triggerrange))
        cached = self._batches[slot]
        if cached is None or cached[0] != key:
            if cached:
                cached[1].close()
            cfgs = [eng.config(fs, settings=settings, triggerrange=triggerrange) for fs in fs_list]
            cached = (key, eng.batch(list(n_samples), cfgs))
            self._batches[slot] = cached
        b = cached[1]
        for i, (p, n) in enumerate(zip(host_ptrs, n_samples)):
            b.upload_ptr(i, int(p), int(n))
        self._pending.append(slot)
        return slot

    def collect(self, full: bool = False):
        """Decode the oldest submitted batch and return its DropResults."""
        slot = self._pending.pop(0)
        b = self._batches[slot][1]
        b.run()
        return [b.result(i, full=full) for i in range(b.n)]


class ConcurrentDecoder:
    """A resident batch decoded as ``shards`` sub-batches on as many engines (CUDA streams), each driven by
    its own host thread (drops are independent, so a GPU's share can be cut further).

    One engine's decode is a chain of about 45 launches with a few host round trips; most of its kernels are
    small and latency bound (chunk chain, frame sync, calibration ...) and its results travel to the host at
    the end.  With two sub-batches in flight those parts run underneath the other sub-batch's big kernels.
    The engines take turns with the demodulation pass itself (``heavy_chain`` engine option), which fills the
    GPU alone.  Drops are dealt round-robin (sub-batch k takes drops k, k + shards, ...)."""

    def __init__(self, device: int, n_samples, fs_list, shards: int = 2, engine_options=None, settings=None,
                 triggerrange=None, engine_factory=None, streams=None):
        from concurrent.futures import ThreadPoolExecutor
        from . import engine as _engine
        n = len(n_samples)
        shards = max(1, min(int(shards), n))
        make = engine_factory or (lambda: _engine.Engine(device))      # (tests inject the host emulation here)
        self.parts = [list(range(k, n, shards)) for k in range(shards)]
        self.where = {}
        for k, part in enumerate(self.parts):
            for j, i in enumerate(part):
                self.where[i] = (k, j)
        self.engines = [make() for _ in range(shards)]
        self.batches = []
        for k, eng in enumerate(self.engines):
            if streams is not None:
                eng.set_stream(streams[k])
            for name, v in (engine_options or {}).items():
                eng.set_option(name, v)
            cfgs = [eng.config(fs_list[i], settings=settings, triggerrange=triggerrange) for i in self.parts[k]]
            self.batches.append(eng.batch([int(n_samples[i]) for i in self.parts[k]], cfgs))
        self.n = n
        self._pool = ThreadPoolExecutor(max_workers=shards)

    def close(self):
        if getattr(self, "_pool", None):
            self._pool.shutdown(wait=True)
            self._pool = None
        for b in getattr(self, "batches", []):
            b.close()
        for e in getattr(self, "engines", []):
            e.close()
        self.batches, self.engines = [], []

    def batch_of(self, i: int):
        """(sub-batch, index inside it) holding drop i."""
        k, j = self.where[i]
        return self.batches[k], j

    def upload(self, i: int, pcm):
        b, j = self.batch_of(i)
        b.upload(j, pcm)

    def upload_ptr(self, i: int, host_ptr: int, n: int):
        b, j = self.batch_of(i)
        b.upload_ptr(j, host_ptr, n)

    def synth_fill(self, i: int, spec):
        b, j = self.batch_of(i)
        b.synth_fill(j, spec)

    def download(self, i: int):
        b, j = self.batch_of(i)
        return b.download(j)

    def run(self, steps: int = 1, before=None, after=None):
        """Decode every sub-batch ``steps`` times, all sub-batches concurrently.  ``before(k)`` / ``after(k)``
        run in sub-batch k's thread around its steps (bench.py records its CUDA events there).  Returns the
        per-step timings of every sub-batch."""
        def work(k):
            b = self.batches[k]
            out = []
            if before:
                before(k)
            for _ in range(steps):
                b.run()
                out.append(b.timing())
            if after:
                after(k)
            return out
        futs = [self._pool.submit(work, k) for k in range(len(self.batches))]
        return [f.result() for f in futs]

    def summary(self, i: int):
        b, j = self.batch_of(i)
        return b.summary(j)

    def result(self, i: int, full: bool = True):
        b, j = self.batch_of(i)
        return b.result(j, full=full)

    def results(self, full: bool = True) -> list:
        return [self.result(i, full=full) for i in range(self.n)]

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.engines)
