set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo pytest=$?
timeout 600 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo bench=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo ref=$?
timeout 300 python bench.py --shards 1 --no-cpu --e2e-steps 1 > gpurun_out/final_bench_shards1.json 2> /dev/null; echo sh1=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 128 -c 150 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --shards 1 > gpurun_out/final_ncu_list.log 2>&1; echo list=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_demod_fused|k_stats_tones|k_tone_windows" --launch-skip 0 -c 10 -o gpurun_out/final_prof -f python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --shards 1 > gpurun_out/final_ncu_full.log 2>&1; echo full=$?
