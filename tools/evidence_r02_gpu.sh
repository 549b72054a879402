#!/bin/bash
# Round-2 evidence on one B200: GPU tests, the bench lines of every workload, the reference arm, the launch list
# and the ncu captures the profiles/r02_* summaries are made from.  Outputs -> gpurun_out/ev2/
set -u
O=gpurun_out/ev2; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_c2.log 2>$O/bench_c2.err; tail -1 $O/bench_c2.log > $O/bench_c2.json
for w in c1 c3 c4 c5; do python bench.py --workload $w --steps 100 --warmup 5 > $O/bench_$w.log 2>&1; tail -1 $O/bench_$w.log > $O/bench_$w.json; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log > $O/bench_ref.json
for w in c1 c2 c3 c4 c5 ref; do python -c "
import json,sys
d=json.load(open('$O/bench_$w.json')); print('$w', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('ms_per_step'), d.get('kernel_ms'), (d.get('roofline') or {}).get('frac'))"; done
python tools/frame_overhead.py > $O/frame_overhead.log 2>&1; tail -5 $O/frame_overhead.log
python tools/build_times.py c3 > $O/build_times.json 2>$O/build_times.err; cat $O/build_times.json
# launch list (cold-cache, serialised): shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launch_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launch_c3.log 2>&1
# full captures of the frame kernels (third frame of tools/one_frame.py)
ncu --set full --clock-control none --import-source on -k regex:'k_frame' --launch-skip 2 -c 1 -f -o $O/prof_c2 python tools/one_frame.py c2 4 > $O/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace_nearest|k_shadow|k_shade' --launch-skip 6 -c 3 -f -o $O/prof_c3 python tools/one_frame.py c3 4 > $O/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace_nearest|k_shadow|k_shade' --launch-skip 6 -c 3 -f -o $O/prof_c5 python tools/one_frame.py c5 4 > $O/ncu_c5.log 2>&1
ls -la $O
