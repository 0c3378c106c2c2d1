#!/bin/bash
# round 2: parity tests (per-test timeout), window errors, bench, launch list, ncu full capture of the fused FFN
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python tests/tools/gpu_window_error.py 96 256 > gpurun_out/window_error.log 2>&1
echo "window exit $?" >> gpurun_out/window_error.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
MADE_FUSED_FFN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 420 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_fused -s 12 -c 3 \
  -o gpurun_out/prof_ffn -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_ffn.log 2>&1
ncu -i gpurun_out/prof_ffn.ncu-rep --page raw --csv > gpurun_out/prof_ffn_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_ffn.ncu-rep --page source --csv > gpurun_out/prof_ffn_source.csv 2>/dev/null
find gpurun_out -name "prof_ffn.ncu-rep" -size +20M -delete
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/window_error.log; cat gpurun_out/bench_split.json; tail -3 gpurun_out/bench_split.err; cat gpurun_out/bench_unfused.json
