#!/bin/bash
# bench on N GPUs exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit $?" >> gpurun_out/bench_${N}gpu.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N=$N", d["scaling"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2)), "strong", d.get("strong_same_job") and (round(d["strong_same_job"]["value"]), round(d["strong_same_job"]["ms_per_step"],3)), d.get("sharded_parity","")[:40])
P
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_${N}gpu.err | tail -4
