set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r33_pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/r33_pytest_gpu.log
OUT=gpurun_out/r33_ab_tone_int8.txt bash tools/ab_bench.sh "tone_int8=1" "tone_int8=0" "tone_int8=1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 128 -c 160 --csv --log-file gpurun_out/r33_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/r33_ncu_list.log 2>&1; echo list=$?
python tools/launch_summary.py gpurun_out/r33_launches.csv > gpurun_out/r33_launch_summary.txt; head -14 gpurun_out/r33_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stats_tones_imma" --launch-skip 0 -c 2 -o gpurun_out/r33_prof_tones -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/r33_ncu_tones.log 2>&1; echo full=$?
