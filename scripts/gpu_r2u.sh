#!/bin/bash
# round 2, session 5, call A: span tests on the packed matcher tail, micro bench (with CUDA-graph device times of the small calls),
# ncu --set full of the gIoU and matcher-cost launches (16384 x 16384)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "span or giou or matcher or iou or postproc" 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_new.json 2> gpurun_out/mb_new.err
tail -3 gpurun_out/mb_new.err
python - <<'P'
import json
d = json.load(open("gpurun_out/micro_bench_new.json"))
for r in d["kernels"]:
    print(f"{r['kernel']:40s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:span_pair_kernel -s 4 -c 2 -f -o gpurun_out/prof_span python scripts/diag_span_ncu.py > gpurun_out/prof_span.log 2>&1
ncu -i gpurun_out/prof_span.ncu-rep --page raw --csv > gpurun_out/prof_span_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_span.ncu-rep --page source --csv > gpurun_out/prof_span_source.csv 2>/dev/null
python scripts/ncu_source_summary.py gpurun_out/prof_span_source.csv 30 all > gpurun_out/prof_span_source_summary.txt 2>&1
find gpurun_out -name "prof_span.ncu-rep" -size +20M -delete
tail -3 gpurun_out/prof_span.log
python - <<'P'
import csv
rows = list(csv.reader(open("gpurun_out/prof_span_raw.csv")))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:40] if "Kernel Name" in hdr else "")
    for w in want:
        if w in hdr:
            print("   ", w, r[hdr.index(w)])
P
