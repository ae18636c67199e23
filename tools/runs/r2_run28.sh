set -x
mkdir -p gpurun_out
out=gpurun_out/r28_ab_shards.txt
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
run --shards 4
run --shards 2
run --shards 3
run --shards 6
run --shards 8
run --shards 4 --opt heavy_chain=0
run --shards 4 --opt heavy_prio=0
run --shards 8 --opt heavy_chain=0
run --shards 4
cat $out
