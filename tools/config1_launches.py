"""One 12-minute 44.1 kHz drop decoded three times (BASELINE config 1): run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel list of a single-drop decode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from axctdprocessor_b200 import engine
eng = engine.Engine(0)
spec = synth.config_spec("config1")
n = int(round(spec.duration_s * spec.fs))
b = eng.batch([n], [eng.config(spec.fs)])
b.synth_fill(0, spec)
for rep in range(3):
    b.run()
print(b.timing(), int(b.summary(0).n_frames), eng.launch_count)
b.close(); eng.close()
