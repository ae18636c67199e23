#!/bin/bash
# copy the outputs of tools/final_run.sh (gpurun_out/) into profiles/ under the round's names
set -e
cd "$(dirname "$0")/.."
R=${1:-r2_final}
cp gpurun_out/final_pytest_gpu.log profiles/${R}_pytest_gpu.log
cp gpurun_out/final_bench_default.json profiles/${R}_bench_default.json
cp gpurun_out/final_bench_ref.json profiles/${R}_bench_ref.json
cp gpurun_out/final_launches.csv profiles/${R}_launches_128drops.csv
python tools/launch_summary.py gpurun_out/final_launches.csv > profiles/${R}_launch_summary.txt
ncu -i gpurun_out/final_prof.ncu-rep --page raw --csv 2>/dev/null > /tmp/raw_final.csv
python tools/ncu_keys.py /tmp/raw_final.csv > profiles/${R}_ncu_full_128drops.txt
if [ -f gpurun_out/final_prof_c3.ncu-rep ]; then
  ncu -i gpurun_out/final_prof_c3.ncu-rep --page raw --csv 2>/dev/null > /tmp/raw_final_c3.csv
  python tools/ncu_keys.py /tmp/raw_final_c3.csv > profiles/${R}_ncu_full_config3.txt
fi
