#!/bin/bash
# round 2, session 5: 2 GPUs - the hardware sharded-vs-single test, then bench.py --gpus 2 as the driver launches it (weak), no CPU baseline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "sharded or two_gpu or nccl" 2>&1 | tail -6 > gpurun_out/pytest_2gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_2gpu.log
tail -3 gpurun_out/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench exit $?" >> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'scaling', d['scaling'], 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2)))
print({k: d[k] for k in d if 'parity' in k or 'sharded' in k or 'strong' in k})
P
