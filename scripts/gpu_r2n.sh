#!/bin/bash
# round 2, call n: suite after the GEMM / span-kernel changes, sanitizers on the changed kernels, launch list with DRAM bytes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 \
  -k "(tcgen05 or span or giou or moment_postproc or hungarian or temporal_encoders or detr_detection or fused_ffn) and not full_size" > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_memcheck.log
tail -4 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 \
  -k "(tcgen05 or span_fast or span_kernels_bit_exact_vs_oracle) and not weight_stationary" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck.log
tail -4 gpurun_out/r02_racecheck.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --detect-topk 0 \
  > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out/launches.csv
