#!/bin/bash
# round 2, session 5: sanitizers on the kernels changed in this session (rank / top-k counting order and wide staging, packed matcher tail)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 \
  -k "(span or moment_postproc or rank_and_topk or topk_value or topk_group or topk_merge or hungarian) and not full_size" > gpurun_out/r02_v_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_v_memcheck.log
tail -4 gpurun_out/r02_v_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 \
  -k "(span_fast or span_kernels_bit_exact_vs_oracle or rank_and_topk or topk_value or topk_group) and not full_size" > gpurun_out/r02_v_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_v_racecheck.log
tail -4 gpurun_out/r02_v_racecheck.log
