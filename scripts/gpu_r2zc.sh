#!/bin/bash
# round 2, session 5: bench.py --gpus N as the driver launches it (weak + the strong record + sharded parity before timing), no CPU baseline
N=${1:-8}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench exit $?" >> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/bench_${N}gpu.err
python - $N <<'P'
import json, sys
n = sys.argv[1]
d=json.loads(open(f'gpurun_out/bench_{n}gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'scaling', d['scaling'], 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2)))
print({k: d[k] for k in d if 'parity' in k or 'strong' in k})
P
