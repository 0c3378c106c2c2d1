#!/bin/bash
# One GPU session: parity tests, smoke, bench, launch list.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -400 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 420 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/bench_under_ncu.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
