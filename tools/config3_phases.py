"""Where the wall clock of BASELINE config 3 goes (96 kHz, one hour, five drops; segment.process_recording):
prints the host-side phases of one warm call and the device phases of its two batches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth
from axctdprocessor_b200 import engine, segment

fs = 96000
eng = engine.Engine(0)
specs = [synth.DropSpec(fs=fs, duration_s=720.0, seed=3300 + i, snr_db=(40.0, 25.0, 10.0)[i % 3]) for i in range(5)]
n = [int(round(s.duration_s * s.fs)) for s in specs]
gen = eng.batch(n, [eng.config(fs / 2, decimate=2)] * len(n))
rec = torch.empty(sum(n), dtype=torch.int16).pin_memory()
o = 0
for i, s in enumerate(specs):
    gen.synth_fill(i, s)
    rec.numpy()[o:o + n[i]] = gen.download(i); o += n[i]
gen.close()
pcm = rec.numpy()
out = {}
for rep in range(3):
    torch.cuda.synchronize()
    T = {}
    t0 = time.perf_counter()
    eng.set_option("scan_only", 1)
    st = {"minr400": 1e300}
    cfg = eng.config(float(fs), settings=st)
    t1 = time.perf_counter(); T["scan_config"] = t1 - t0
    b = eng.batch([len(pcm)], [cfg])
    t2 = time.perf_counter(); T["scan_batch_create"] = t2 - t1
    b.upload(0, pcm)
    torch.cuda.synchronize()
    t3 = time.perf_counter(); T["scan_upload_sync"] = t3 - t2
    b.run()
    t4 = time.perf_counter(); T["scan_run"] = t4 - t3
    T["scan_device"] = b.timing()
    p, r400, _ = b.power(0)
    eng.set_option("scan_only", 0)
    t5 = time.perf_counter(); T["scan_power_download"] = t5 - t4
    raw = segment.find_drops(p, r400, float(fs), len(pcm))
    t6 = time.perf_counter(); T["find_drops"] = t6 - t5
    cfg2 = eng.config(fs / 2, decimate=2)
    b2 = eng.batch([hi - lo for lo, hi in raw], [cfg2] * len(raw))
    t7 = time.perf_counter(); T["seg_batch_create"] = t7 - t6
    for i, (lo, hi) in enumerate(raw):
        b2.copy_from(i, b, 0, lo, hi - lo)
    b2.run()
    t8 = time.perf_counter(); T["seg_copy_run"] = t8 - t7
    T["seg_device"] = b2.timing()
    res = [b2.result(i) for i in range(len(raw))]
    t9 = time.perf_counter(); T["seg_results"] = t9 - t8
    b2.close(); b.close()
    t10 = time.perf_counter(); T["close"] = t10 - t9
    T["total"] = t10 - t0
    out[f"rep{rep}"] = T
print(json.dumps(out, indent=1, default=str))
eng.close()
