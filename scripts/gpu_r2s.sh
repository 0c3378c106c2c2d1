#!/bin/bash
# round 2, session 4, call B: full gpu suite, ncu --set full of rank_topk (8192 x 16384) and of the packed span kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
prof() {  # name, kernel regex, skip, script
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python $4 > gpurun_out/prof_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
  python scripts/ncu_source_summary.py gpurun_out/prof_$1_source.csv 30 > gpurun_out/prof_$1_source_summary.txt 2>&1
  find gpurun_out -name "prof_$1.ncu-rep" -size +20M -delete
}
prof rank rank_topk_staged 1 scripts/diag_rank_ncu.py
prof giou "span_pair_kernel<0>" 1 scripts/diag_span_ncu.py
prof cost "span_pair_kernel<2>" 1 scripts/diag_span_ncu.py
head -5 gpurun_out/prof_rank_source_summary.txt
