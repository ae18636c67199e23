set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decim or g96 or config3 or multi_drop or small" > gpurun_out/r9_pytest_decim.log 2>&1; echo pytest=$?
tail -30 gpurun_out/r9_pytest_decim.log
timeout 300 python tools/config3_phases.py > gpurun_out/r9_config3_phases.json 2> gpurun_out/r9_config3_phases.err; echo c3=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r9_config3_phases.json'))
for k,v in d['rep2'].items(): print(k, v)
PY


