"""Host->device copy bandwidth from pinned memory, per NUMA node of the allocating thread (e2e leg tuning)."""
import os, subprocess, sys, time
import torch

def nodes():
    out = {}
    base = "/sys/devices/system/node"
    for d in sorted(os.listdir(base)):
        if d.startswith("node") and d[4:].isdigit():
            cpus = open(f"{base}/{d}/cpulist").read().strip()
            s = set()
            for part in cpus.split(","):
                if not part:
                    continue
                a, _, b = part.partition("-")
                s.update(range(int(a), int(b or a) + 1))
            out[int(d[4:])] = s
    return out

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
allowed = os.sched_getaffinity(0)
print("allowed cpus:", len(allowed))
torch.cuda.init()
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for node, cpus in nodes().items():
    use = cpus & allowed
    if not use:
        print("node", node, "no allowed cpus"); continue
    os.sched_setaffinity(0, use)
    host = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    for chunks in (1, 16):
        n = (1 << 30) // chunks
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for rep in range(4):
            for c in range(chunks):
                dev[c * n:(c + 1) * n].copy_(host[c * n:(c + 1) * n], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"node {node} ({len(use)} cpus) chunks={chunks}: H2D {4 * (1 << 30) / dt / 1e9:.1f} GB/s")
    t0 = time.perf_counter()
    for rep in range(4):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    print(f"node {node}: D2H {4 * (1 << 30) / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
    del host
os.sched_setaffinity(0, allowed)
