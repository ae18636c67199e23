set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r46_bench_n2.json 2> gpurun_out/r46_bench_n2.err; echo bench2=$?
wc -l gpurun_out/r46_bench_n2.json; head -c 100 gpurun_out/r46_bench_n2.json; echo
python -c "
import json; d=json.load(open('gpurun_out/r46_bench_n2.json')); print(d['n_gpus'], d['clocks'], d['ms_per_step'], d['value'], d['e2e']['value'])"
grep -c "NCCL version" gpurun_out/r46_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r46_bench_ref_n2.json 2> gpurun_out/r46_bench_ref_n2.err; echo ref2=$?
wc -l gpurun_out/r46_bench_ref_n2.json; head -c 100 gpurun_out/r46_bench_ref_n2.json; echo
