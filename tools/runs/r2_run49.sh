set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_compact_warp|k_emit_chunk|k_tiles_reg|k_tone_windows_mma|k_valid|k_calib" --launch-skip 0 -c 12 -o gpurun_out/r49_prof_book -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 4 > gpurun_out/r49_ncu_book.log 2>&1; echo book=$?
