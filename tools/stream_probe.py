"""Per-poll cost of the incremental decoder against re-decoding the whole prefix: a 12-minute drop arrives one second
at a time.  Prints one JSON line: device milliseconds per poll early / late in the recording for the streaming
decoder (axctd_batch_stream_*), and the same for a batch decode of the prefix."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
from axctdprocessor_b200 import engine
from axctdprocessor_b200.stream import StreamingDecoder, prefix_normalisation

eng = engine.Engine(0)
spec = synth.DropSpec(fs=44100, duration_s=720.0, seed=4711, snr_db=25.0)
n = int(round(spec.duration_s * spec.fs))
g = eng.batch([n], [eng.config(spec.fs)])
g.synth_fill(0, spec)
pcm = g.download(0)
g.close()
sd = StreamingDecoder(spec.fs, engine=eng, max_seconds=740.0, norm_seconds=2.0)
wall = []
for a in range(0, n, spec.fs):
    t0 = time.perf_counter()
    sd.push(pcm[a:a + spec.fs])
    sd.poll()
    wall.append(1e3 * (time.perf_counter() - t0))
r = sd.finish()
runs = sd.runs
def med(xs): return float(np.median(xs))
prefix = {}
cfg = eng.config(spec.fs)
for sec in (60, 360, 720):
    b = eng.batch([sec * spec.fs], [cfg])
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        b.upload(0, pcm[:sec * spec.fs]); b.run(); b.result(0, full=False)
        ts.append(1e3 * (time.perf_counter() - t0))
    prefix[f"{sec}s"] = {"wall_ms": med(ts[1:]), "device_ms": b.timing()["total_ms"], "h2d_bytes": 2 * sec * spec.fs}
    b.close()
line = {"workload": "720 s drop at 44.1 kHz arriving 1 s per poll", "polls": len(runs), "frames": int(r.summary.n_frames), "status": int(r.summary.status),
        "streaming": {"device_ms_per_poll_60_120s": med([x["device_ms"] for x in runs[60:120]]),
                      "device_ms_per_poll_last_60s": med([x["device_ms"] for x in runs[-61:-1]]),
                      "filter_ms_per_poll_60_120s": med([x["filter_ms"] for x in runs[60:120]]),
                      "filter_ms_per_poll_last_60s": med([x["filter_ms"] for x in runs[-61:-1]]),
                      "wall_ms_per_push_poll_60_120s": med(wall[60:120]), "wall_ms_per_push_poll_last_60s": med(wall[-61:-1]),
                      "h2d_bytes_per_poll": 2 * spec.fs},
        "whole_prefix_redecode": prefix}
print(json.dumps(line))
sd._own = False
sd.close()
eng.close()
