set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo bench=$?
timeout 300 python tools/config3_phases.py > gpurun_out/r5_config3_phases.json 2> gpurun_out/r5_config3_phases.err; echo c3=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 128 -c 150 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/r5_ncu_list.log 2>&1; echo list=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5_c3_launches.csv python tools/config3.py > gpurun_out/r5_c3_ncu.log 2>&1; echo c3list=$?
