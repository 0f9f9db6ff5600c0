#!/bin/bash
# Multi-GPU bench lines of one box: usage scale_run.sh N tag [bench args...]   (writes gpurun_out/r02_bench_n${N}_${tag}.json)
N=$1; tag=$2; shift 2
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --no-cpu-baseline "$@" > gpurun_out/r02_bench_n${N}_${tag}.json 2> gpurun_out/r02_bench_n${N}_${tag}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline "$@" \
    > gpurun_out/r02_bench_n${N}_${tag}.json 2> gpurun_out/r02_bench_n${N}_${tag}.err
fi
cat gpurun_out/r02_bench_n${N}_${tag}.json | python -c "
import sys, json
try:
    d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
    print('$N $tag', d['value'], d['ms_per_step'], d.get('e2e', {}).get('value'), d.get('dp_check'), d['config'].get('workload', '')[:60])
except Exception as e:
    print('parse failed', e)
"
