set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r48_bench_n8.json 2> gpurun_out/r48_bench_n8.err; echo bench8=$?
wc -l gpurun_out/r48_bench_n8.json
python -c "
import json; d=json.load(open('gpurun_out/r48_bench_n8.json')); print(d['n_gpus'], d['clocks'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['fraction_of_h2d_ceiling'])"
