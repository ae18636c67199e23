set -x
mkdir -p gpurun_out
out=gpurun_out/r43_ab_seg_target.txt
: > $out
for o in 16384 24576 32768 49152 16384 32768; do
  echo "== seg_target=$o" >> $out
  python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --host-pool 2 --opt seg_target=$o 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('ms_per_step %.3f  demod_ms %.3f  tone_ms %.3f  frac %.4f  value %.0f' % (d['ms_per_step'], r['kernel_ms'], r['tone_kernels_ms'], r['frac'], d['value']))
" >> $out
done
cat $out
