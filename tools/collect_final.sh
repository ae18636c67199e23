#!/bin/bash
# copy the outputs of tools/final_run.sh (gpurun_out/) into profiles/ under the round's names
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/final_pytest_gpu.log profiles/r1_final_pytest_gpu.log
cp gpurun_out/final_bench_default.json profiles/r1_final_bench_default.json
cp gpurun_out/final_bench_ref.json profiles/r1_final_bench_ref.json
cp gpurun_out/final_bench_shards1.json profiles/r1_final_bench_one_subbatch.json
cp gpurun_out/final_launches.csv profiles/r1_final_launches_128drops.csv
python tools/launch_summary.py gpurun_out/final_launches.csv > profiles/r1_final_launch_summary.txt
ncu -i gpurun_out/final_prof.ncu-rep --page raw --csv 2>/dev/null > /tmp/raw_final.csv
python tools/ncu_keys.py /tmp/raw_final.csv > profiles/r1_final_ncu_full_128drops.txt
