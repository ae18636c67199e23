set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decim or g96 or config3 or multi_drop or small" > gpurun_out/r8_pytest_decim.log 2>&1; echo pytest=$?
tail -30 gpurun_out/r8_pytest_decim.log
timeout 300 python tools/config3_phases.py > gpurun_out/r8_config3_phases.json 2> gpurun_out/r8_config3_phases.err; echo c3=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r8_config3_phases.json'))
for k,v in d['rep2'].items(): print(k, v)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r8_c3_launches.csv python tools/config3.py > gpurun_out/r8_c3_ncu.log 2>&1; echo c3list=$?
python tools/launch_summary.py gpurun_out/r8_c3_launches.csv | head -12
