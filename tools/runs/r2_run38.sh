set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r38_pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/r38_pytest_gpu.log
out=gpurun_out/r38_ab_pair_launch.txt
: > $out
run() {
  echo "== $*" >> $out
  python bench.py --steps 10 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --host-pool 2 "$@" 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('ms_per_step %.3f  demod_ms %.3f  tone_ms %.3f  frac %.4f  value %.0f' % (d['ms_per_step'], r['kernel_ms'], r['tone_kernels_ms'], r['frac'], d['value']))
" >> $out
}
run --shards 1 --opt pair_launch=1
run --shards 1 --opt pair_launch=0
run --shards 2 --opt pair_launch=1
run --shards 2 --opt pair_launch=0
run --shards 4
cat $out
