set -x
mkdir -p gpurun_out
timeout 105 python -m pytest tests -m gpu -q -n 3 > gpurun_out/r51_pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/r51_pytest_gpu.log
timeout 40 python bench.py --no-configs --no-cpu --steps 10 --e2e-steps 1 > gpurun_out/r51_bench.json 2> gpurun_out/r51_bench.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/r51_bench.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
