"""A/B of the statistics + tone-block-sum pass: pure DMMA (tone_mma=1) against the tensor-core / vector-pipe split
(tone_mma = k-steps of every 16 that stay on the tensor cores).  64 drops x 720 s, device-resident."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
from axctdprocessor_b200 import engine

specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=720.0, seed=900 + i, snr_db=(40.0, 25.0, 10.0)[i % 3]) for i in range(64)]
n = [int(round(s.duration_s * s.fs)) for s in specs]
out = {}
for opt in (1, 8, 10, 12, 14, 0):
    eng = engine.Engine(0)
    eng.set_option("tone_mma", opt)
    b = eng.batch(n, [eng.config(s.fs) for s in specs])
    for i, s in enumerate(specs):
        b.synth_fill(i, s)
    ing, tot = [], []
    for rep in range(6):
        b.run()
        t = b.timing()
        if rep:
            ing.append(t["ingest_ms"]); tot.append(t["total_ms"])
    frames = sum(int(b.summary(i).n_frames) for i in range(len(n)))
    sig = [(int(b.summary(i).firstpulse400), int(b.summary(i).profstartind)) for i in range(len(n))]
    out[opt] = dict(ingest_ms=float(np.median(ing)), total_ms=float(np.median(tot)), frames=frames, sig=hash(tuple(sig)))
    b.close(); eng.close()
print(json.dumps(out, indent=1))
