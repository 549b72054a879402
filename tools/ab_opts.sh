#!/bin/bash
# A/B runtime options of ONE build on the same box:  WL="c2 c3" tools/ab_opts.sh "fused_frame=0" "fused_frame=1" ...
WL=${WL:-c2 c3 c4}
for w in $WL; do
  for rep in 1 2; do
    for o in "$@"; do
      args=""; for kv in $o; do args="$args --opt $kv"; done
      python bench.py --workload $w --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline $args 2>/dev/null | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w [$o]', round(d['ms_per_step'],4), 'ms', round(d['value']), 'Mrays/s  e2e', round(d['e2e']['ms_per_step'],4), {k: round(v,4) for k,v in d['kernel_ms'].items() if not isinstance(v, dict)})"
    done
  done
done
