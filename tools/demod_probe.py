"""Where k_demod_fused's time goes (timing probe, results of probes 1 and 2 are wrong by construction):
engine option demod_probe = 0 (the real pass), 1 (window sums skipped: loop over the crossings and record stores
only), 2 (crossings not processed: staging + cascade + sign bits only).  Needs a library built with -DAX_DEMOD_PROBE
(the branch itself costs 5 %, so it is not in the default build).  32 drops x 720 s, one sub-batch; prints the
pass duration from the engine's own CUDA events (Batch.timing()['filter_ms'])."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from axctdprocessor_b200 import batch as axbatch


def main():
    drops = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    extra = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in sys.argv[2:]}
    specs = bench.drop_specs(drops, 720.0, 0)
    n = [int(round(s.duration_s * s.fs)) for s in specs]
    out = {}
    for probe in (0, 1, 2, 0):
        opts = dict(extra)
        opts["demod_probe"] = float(probe)
        b = axbatch.ConcurrentDecoder(0, n, [s.fs for s in specs], shards=1, engine_options=opts)
        for i, s in enumerate(specs):
            b.synth_fill(i, s)
        ms = []
        try:
            for _ in range(5):
                t = b.run(steps=1)
                ms.append(t[0][0]["filter_ms"])
        except Exception as ex:            # the probes hand garbage to the rest of the decode
            print("probe", probe, "decode raised:", repr(ex)[:200], file=sys.stderr)
        b.close()
        out.setdefault("probe%d" % probe, []).append(float(np.median(ms[1:])) if len(ms) > 1 else None)
    out["samples"] = int(sum(n))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
