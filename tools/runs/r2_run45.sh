set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r45_bench_n2.json 2> gpurun_out/r45_bench_n2.err; echo bench2=$?
python -c "
import json; d=json.load(open('gpurun_out/r45_bench_n2.json')); print(d['n_gpus'], d['clocks'], d['ms_per_step'], d['value'], d['e2e']['value'])"
tail -3 gpurun_out/r45_bench_n2.err
timeout 600 python -m pytest tests -m gpu -q -k "two or rank or devices or sharded" > gpurun_out/r45_pytest_2gpu.log 2>&1; echo pytest2=$?
tail -3 gpurun_out/r45_pytest_2gpu.log
