#!/bin/bash
# tests, bench (new bench.py), launch list, ncu full of the fused FFN
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python tests/tools/gpu_window_error.py 96 256 > gpurun_out/window_error.log 2>&1
for i in 1 2; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 420 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_fused -s 12 -c 3 \
  -o gpurun_out/prof_ffn -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_ffn.log 2>&1
ncu -i gpurun_out/prof_ffn.ncu-rep --page raw --csv > gpurun_out/prof_ffn_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_ffn.ncu-rep --page source --csv > gpurun_out/prof_ffn_source.csv 2>/dev/null
find gpurun_out -name "prof_ffn.ncu-rep" -size +20M -delete
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/window_error.log; cat gpurun_out/bench_1.json; tail -3 gpurun_out/bench_1.err; cat gpurun_out/bench_2.json
