#!/bin/bash
# round 2, session 5, call B: rank / top-k tests on the counting order + wide staging, micro bench, ncu --set full of both shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "rank or topk or recall or sharded or gallery or full_size or index" 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_new.json 2> gpurun_out/mb_new.err
tail -3 gpurun_out/mb_new.err
python - <<'P'
import json
d = json.load(open("gpurun_out/micro_bench_new.json"))
for r in d["kernels"]:
    if "rank" in r["kernel"] or "postproc" in r["kernel"]:
        print(f"{r['kernel']:40s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rank_topk_staged -s 1 -c 2 -f -o gpurun_out/prof_rank python scripts/diag_rank_ncu.py > gpurun_out/prof_rank.log 2>&1
ncu -i gpurun_out/prof_rank.ncu-rep --page raw --csv > gpurun_out/prof_rank_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rank.ncu-rep --page source --csv > gpurun_out/prof_rank_source.csv 2>/dev/null
python scripts/ncu_source_summary.py gpurun_out/prof_rank_source.csv 30 all > gpurun_out/prof_rank_source_summary.txt 2>&1
find gpurun_out -name "prof_rank.ncu-rep" -size +20M -delete
tail -2 gpurun_out/prof_rank.log
grep -n "^kernel\|^stall" gpurun_out/prof_rank_source_summary.txt | cut -c1-300
