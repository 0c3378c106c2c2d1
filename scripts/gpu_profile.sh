#!/bin/bash
# ncu --set full captures of the hot kernels (one GPU).  Reports land in gpurun_out/*.ncu-rep.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# launches per step ~304; skip the first step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xpool_score -s 1 -c 1 \
  -o gpurun_out/prof_xpool -f $B > gpurun_out/prof_xpool.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 253 -c 40 \
  -o gpurun_out/prof_gemm -f $B > gpurun_out/prof_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mha_core -s 17 -c 3 \
  -o gpurun_out/prof_mha -f $B > gpurun_out/prof_mha.log 2>&1
ls -la gpurun_out/*.ncu-rep
