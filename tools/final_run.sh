# Round-end measurement set on one B200 (outputs under gpurun_out/final_*; tools/collect_final.sh copies them to profiles/)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo pytest=$?
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1; echo smoke=$?; tail -4 gpurun_out/final_smoke.log
tail -3 gpurun_out/final_pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err ) 2> gpurun_out/final_bench_default.time; echo bench=$?; cat gpurun_out/final_bench_default.time
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo ref=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 128 -c 160 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/final_ncu_list.log 2>&1; echo list=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_demod_fused|k_stats_tones|k_tone_windows" --launch-skip 0 -c 10 -o gpurun_out/final_prof -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/final_ncu_full.log 2>&1; echo full=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decim_fused|k_demod_fused" --launch-skip 0 -c 6 -o gpurun_out/final_prof_c3 -f python tools/config3.py > gpurun_out/final_ncu_full_c3.log 2>&1; echo fullc3=$?
