#!/bin/bash
# tools/multigpu_gpu.sh N : C2 and C5 bench lines on N GPUs of one box -> gpurun_out/ev/
N=$1; O=gpurun_out/ev; mkdir -p $O
for w in c2 c5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $w --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline > $O/bench_${w}_${N}gpu.log 2>&1
  grep '^{' $O/bench_${w}_${N}gpu.log | tail -1 > $O/bench_${w}_${N}gpu.json
  python -c "
import json
d=json.load(open('$O/bench_${w}_${N}gpu.json')); print('$w', $N, 'gpus', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d.get('nccl_gather_baseline'))"
done
