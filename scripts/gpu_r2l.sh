#!/bin/bash
# round 2, call l: new tests, memcheck + racecheck of the final kernels, ncu full of the split-mode GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests -m gpu -q -x --timeout 1200 \
  -k "not full_size and not 16384" > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_memcheck.log
tail -5 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 \
  -k "tcgen05_gemm or fused_ffn or xpool_scoring or mha_core or topk_value or ca_fusion" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck.log
tail -5 gpurun_out/r02_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 30 -c 16 \
  -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_gemm.ncu-rep --page source --csv > gpurun_out/prof_gemm_source.csv 2>/dev/null
find gpurun_out -name "prof_gemm.ncu-rep" -size +20M -delete
ls -la gpurun_out/prof_gemm_raw.csv
