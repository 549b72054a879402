#!/bin/bash
# Round evidence on one B200: GPU tests, the bench lines of every workload, the reference arm, the
# launch list and the ncu captures the profiles/ summaries are made from.  Outputs -> gpurun_out/ev/
set -u
O=gpurun_out/ev; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_c2.log 2>&1; tail -1 $O/bench_c2.log > $O/bench_c2.json
for w in c1 c3 c4 c5; do python bench.py --workload $w --steps 100 --warmup 5 > $O/bench_$w.log 2>&1; tail -1 $O/bench_$w.log > $O/bench_$w.json; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log > $O/bench_ref.json
for w in c1 c2 c3 c4 c5 ref; do python -c "
import json,sys
d=json.load(open('$O/bench_$w.json')); print('$w', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d.get('kernel_ms'))"; done
# launch list (cold-cache, serialised): shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
