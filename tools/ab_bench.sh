#!/bin/bash
# A/B of engine options on the default bench workload (device-resident figure only): tools/ab_bench.sh "bulk=0" "bulk=1" ...
out=${OUT:-gpurun_out/ab.txt}
: > $out
for o in "$@"; do
  args=""
  for kv in $o; do args="$args --opt $kv"; done
  echo "== $o" >> $out
  python bench.py --steps 10 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --host-pool 2 $args 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('ms_per_step %.3f  demod_ms %.3f  tone_ms %.3f  frac %.4f  value %.0f' % (d['ms_per_step'], r['kernel_ms'], r['tone_kernels_ms'], r['frac'], d['value']))
    else: print(line)
" >> $out
done
cat $out
