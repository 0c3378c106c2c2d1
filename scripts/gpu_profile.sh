#!/bin/bash
# ncu --set full captures of the hot kernels (one GPU).  The .ncu-rep files stay small (few
# launches) and the raw/source pages are exported to CSV on the box; gpurun_out/ is capped at 64 MiB.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
WHAT="${@:-xpool gemm mha}"
for w in $WHAT; do
  case $w in
    xpool) K="regex:xpool_score"; S=4; C=1;;
    gemm)  K="regex:gemm_tc_kernel"; S=${GEMM_SKIP:-150}; C=${GEMM_COUNT:-8};;
    mha)   K="regex:mha_core"; S=10; C=2;;
    dec)   K="regex:dec_attn_folded"; S=6; C=1;;
    rank)  K="regex:rank_topk"; S=2; C=1;;
    *) echo "unknown $w"; continue;;
  esac
  timeout 900 ncu --set full --clock-control none --import-source on -k $K -s $S -c $C \
    -o gpurun_out/prof_$w -f $B > gpurun_out/prof_$w.log 2>&1
  ncu -i gpurun_out/prof_$w.ncu-rep --page raw --csv > gpurun_out/prof_${w}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$w.ncu-rep --page source --csv > gpurun_out/prof_${w}_source.csv 2>/dev/null
  ls -la gpurun_out/prof_$w*
  # keep the transfer small: drop reports above 20 MB (the CSV pages carry what is read here)
  find gpurun_out -name "prof_$w.ncu-rep" -size +20M -delete
done
du -sh gpurun_out
