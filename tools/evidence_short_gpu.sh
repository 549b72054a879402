set -u
O=gpurun_out/ev; mkdir -p $O
python bench.py > $O/bench_c2.log 2>&1; tail -1 $O/bench_c2.log > $O/bench_c2.json
for w in c1 c3 c4 c5; do python bench.py --workload $w --steps 100 --warmup 5 > $O/bench_$w.log 2>&1; tail -1 $O/bench_$w.log > $O/bench_$w.json; done
for w in c1 c2 c3 c4 c5; do python -c "
import json,sys
d=json.load(open('$O/bench_$w.json')); print('$w', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['pipelined']['ms_per_step'],4), d.get('kernel_ms'))"; done
