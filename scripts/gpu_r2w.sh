#!/bin/bash
# round 2, session 5, call C: rank kernel with the inline ground-truth key; gallery-first copy order of the e2e step (A/B)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "rank or topk or recall or sharded or gallery or full_size or index or pinned or host or ingest or h2d" 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
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
MADE_PRIME_GALLERY=1 timeout 200 python scripts/diag_e2e.py dma 2>&1 | tail -2 | sed 's/^/prime=1 /'
MADE_PRIME_GALLERY=0 timeout 200 python scripts/diag_e2e.py dma 2>&1 | tail -2 | sed 's/^/prime=0 /'
MADE_PRIME_GALLERY=1 timeout 200 python scripts/diag_e2e.py dma 2>&1 | tail -2 | sed 's/^/prime=1 /'
