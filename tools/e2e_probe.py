"""Where the end-to-end leg's time goes: upload loop, decode, host-side result extraction (one half-batch)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth
from axctdprocessor_b200 import engine

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = engine.Engine(0)
specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=720.0, seed=100 + i, snr_db=(40.0, 25.0, 10.0)[i % 3]) for i in range(nd)]
n = [int(round(s.duration_s * s.fs)) for s in specs]
cfg = {fs: eng.config(fs) for fs in (44100, 48000)}
b = eng.batch(n, [cfg[s.fs] for s in specs])
for i in range(2):
    b.synth_fill(i, specs[i])
pinned = []
for i in range(2):
    t = torch.empty(n[i], dtype=torch.int16).pin_memory(); t.numpy()[:] = b.download(i); pinned.append(t)
src = [pinned[i % 2] for i in range(nd)]
tot = 2 * sum(n)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(nd):
        b.upload_ptr(i, src[i].data_ptr(), n[i])
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    b.run(); t3 = time.perf_counter()
    res = [b.result(i, full=False) for i in range(nd)]; t4 = time.perf_counter()
    print(f"rep {rep}: upload calls {1e3*(t1-t0):.1f} ms, H2D done {1e3*(t2-t0):.1f} ms ({tot/(t2-t0)/1e9:.1f} GB/s), decode {1e3*(t3-t2):.1f} ms, results {1e3*(t4-t3):.1f} ms")
# decode while a long host->device copy runs on another stream
big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dbig = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for k in range(6):
            dbig.copy_(big, non_blocking=True)
    t0 = time.perf_counter()
    b.run()
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"concurrent rep {rep}: decode {1e3*(t1-t0):.1f} ms while a {6*(1<<30)/1e9:.1f} GB H2D takes {1e3*(t2-t0):.1f} ms")
