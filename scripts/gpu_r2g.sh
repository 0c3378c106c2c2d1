#!/bin/bash
# round 2, call g: all gpu tests (fp32 mode, compat mirrors, checkpoint), smoke, memcheck of the whole suite, racecheck of the
# hand-rolled mbarrier / named-barrier kernels at small shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -25 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests -m gpu -q -x --timeout 1200 \
  -k "not full_size and not 16384" > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_memcheck.log
tail -12 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 \
  -k "tcgen05_gemm or fused_ffn or xpool_scoring or mha_core or topk_value" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck.log
tail -12 gpurun_out/r02_racecheck.log
