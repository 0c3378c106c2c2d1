#!/bin/bash
# round 2, first GPU session: parity tests, window errors per precision mode, bench per precision mode, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python tests/tools/gpu_window_error.py 96 256 > gpurun_out/window_error.log 2>&1
echo "window exit $?" >> gpurun_out/window_error.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
MADE_PRECISION=fp16 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 420 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/window_error.log; cat gpurun_out/bench_split.json; tail -3 gpurun_out/bench_split.err; cat gpurun_out/bench_fp16.json
