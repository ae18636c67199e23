set -x
mkdir -p gpurun_out
out=gpurun_out/r47_ab_shards.txt
: > $out
for o in 4 3 5 6 8 4; do
  echo "== shards=$o" >> $out
  python bench.py --steps 15 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --host-pool 2 --shards $o 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('ms_per_step %.3f  demod_ms %.3f  tone_ms %.3f  frac %.4f  value %.0f' % (d['ms_per_step'], r['kernel_ms'], r['tone_kernels_ms'], r['frac'], d['value']))
" >> $out
done
cat $out
