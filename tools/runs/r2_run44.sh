set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r44_pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/r44_pytest_gpu.log
timeout 600 python bench.py --no-configs --no-cpu --steps 10 > gpurun_out/r44_bench.json 2> gpurun_out/r44_bench.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/r44_bench.json')); print(d['clocks'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 128 -c 160 --csv --log-file gpurun_out/r44_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/r44_ncu_list.log 2>&1; echo list=$?
python tools/launch_summary.py gpurun_out/r44_launches.csv > gpurun_out/r44_launch_summary.txt; head -14 gpurun_out/r44_launch_summary.txt
