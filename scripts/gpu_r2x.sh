#!/bin/bash
# round 2, session 5, call D: rank kernel (ground-truth key after the staging loads): tests + micro bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "rank or topk or recall or sharded or index" 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_new.json 2> gpurun_out/mb_new.err
python - <<'P'
import json
d = json.load(open("gpurun_out/micro_bench_new.json"))
for r in d["kernels"]:
    if "rank" in r["kernel"]:
        print(f"{r['kernel']:40s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
P
