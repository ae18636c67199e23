set -x
mkdir -p gpurun_out
timeout 600 python bench.py --no-configs --no-cpu --steps 10 > gpurun_out/r25_bench.json 2> gpurun_out/r25_bench.err; echo bench=$?
python -c "
import json; d=json.load(open('gpurun_out/r25_bench.json')); print(d['clocks'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_compact_warp|k_emit_chunk|k_tiles_reg|k_chain_warp|k_levels|k_headers_warp|k_valid" --launch-skip 0 -c 14 -o gpurun_out/r25_prof_book -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-steps 1 --shards 1 > gpurun_out/r25_ncu_book.log 2>&1; echo book=$?
