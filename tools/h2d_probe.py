"""Aggregate host->device bandwidth of one box, one rank per GPU.

    python tools/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Every rank copies a pool of pinned host memory to its GPU at the same time (barrier-bracketed, CUDA events, max over
ranks) and rank 0 prints one JSON line per variant: plain pinned vs write-combined pinned memory, pool size
(cache-resident vs DRAM-resident), one or two copy streams in flight, host thread bound to the GPU's NUMA node or not.
The aggregate is the ceiling bench.py's end-to-end figure can reach on that box: the e2e leg moves 2 bytes per sample
over these links and nothing else is close to binding (decode 17 ms vs 156 ms of copy per 8.5 GB step).
"""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cudart():
    for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    import glob
    for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        return ctypes.CDLL(p)
    raise OSError("libcudart not found")


class HostPool:
    """cudaHostAlloc'ed buffer (flags: 0 default, 4 write-combined) exposed as a torch uint8 tensor."""

    def __init__(self, nbytes, flags):
        self.rt = cudart()
        self.ptr = ctypes.c_void_p()
        rc = self.rt.cudaHostAlloc(ctypes.byref(self.ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
        if rc != 0:
            raise RuntimeError(f"cudaHostAlloc failed ({rc})")
        self.nbytes = nbytes
        buf = (ctypes.c_uint8 * nbytes).from_address(self.ptr.value)
        self.t = torch.frombuffer(buf, dtype=torch.uint8)

    def free(self):
        self.t = None
        self.rt.cudaFreeHost(self.ptr)


def main():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from axctdprocessor_b200 import batch as axbatch
    dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    piece = 64 << 20                                   # one 12-minute drop is 63-69 MB
    variants = []
    for bound in (False, True):
        if bound and not axbatch.bind_host_thread_to_gpu(local):
            continue
        for flags, fname in ((0, "pinned"), (4, "write-combined")):
            for pool_mb in (512, 4096):
                try:
                    pool = HostPool(pool_mb << 20, flags)
                except RuntimeError:
                    continue
                pool.t[:: 4096] = 1                     # touch every page from this thread
                for nstreams in (1, 2):
                    streams = [torch.cuda.Stream() for _ in range(nstreams)]
                    total = 8 << 30
                    npieces = total // piece

                    def run():
                        for q in range(npieces):
                            off = (q * piece) % (pool.nbytes - piece + 1)
                            with torch.cuda.stream(streams[q % nstreams]):
                                dev[(q % 16) * piece:(q % 16 + 1) * piece].copy_(pool.t[off:off + piece], non_blocking=True)
                    run()
                    barrier()
                    t0 = time.perf_counter()
                    run()
                    torch.cuda.synchronize()
                    dt = reduce_max(time.perf_counter() - t0)
                    barrier()
                    variants.append({"memory": fname, "pool_mb": pool_mb, "streams": nstreams, "numa_bound": bound,
                                     "per_gpu_gbs": total / dt / 1e9, "aggregate_gbs": world * total / dt / 1e9})
                pool.free()
    if rank == 0:
        best = max(variants, key=lambda v: v["aggregate_gbs"])
        print(json.dumps({"n_gpus": world, "cpus_allowed": len(os.sched_getaffinity(0)), "best": best, "variants": variants}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
