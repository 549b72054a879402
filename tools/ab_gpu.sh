#!/bin/bash
# A/B several builds of librt_b200.so on the same box:  WL="c2 c3" tools/ab_gpu.sh <lib> <lib> ...
WL=${WL:-c2 c3 c4}
for w in $WL; do
  for rep in 1 2; do
    for lib in "$@"; do
      RT_B200_LIB=$PWD/$lib python bench.py --workload $w --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['config']['workload'][:24], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v,4) for k,v in d['kernel_ms'].items() if not isinstance(v, dict)})"
    done
  done
done
