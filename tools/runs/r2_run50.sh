set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_emit_chunk|k_valid|k_calib" --launch-skip 0 -c 8 -o gpurun_out/r50_prof_emit -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 4 > gpurun_out/r50_ncu_emit.log 2>&1; echo emit=$?
