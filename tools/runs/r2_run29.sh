set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r29_pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/r29_pytest_gpu.log
timeout 600 python bench.py --no-configs --no-cpu --steps 10 > gpurun_out/r29_bench.json 2> gpurun_out/r29_bench.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/r29_bench.json')); print(d['clocks'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
