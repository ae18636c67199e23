"""Small decodes that touch every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck) runs:

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py

Each case is checked against its reference fixture, so a sanitizer-clean run is also a parity run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_util import Golden  # noqa: E402
from parity_util import check_against_golden, run_engine  # noqa: E402

from axctdprocessor_b200 import engine  # noqa: E402

CASES = [("g44_40db", {}), ("g44_40db", {"bulk": 1}), ("g48_25db", {"ws": 1}), ("g44_bandpass", {}), ("g44_bandpass", {"bulk": 1}),
         ("g96_decim", {}), ("g44_stereo", {}), ("g44_timeout", {}), ("g44_chunk05", {"tone_mma": 0}), ("g44_10db", {"guard": 3e-6})]
only = sys.argv[1:]
for name, opts in CASES:
    if only and name not in only:
        continue
    g = Golden(name)
    e = engine.Engine(0)
    for k, v in opts.items():
        e.set_option(k, v)
    out = run_engine(e, g.pcm(), g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange)
    check_against_golden(out, g)
    print("ok", name, opts, "launches", e.launch_count, flush=True)
    e.close()
