#!/bin/bash
# round 2, call j: full gpu suite, smoke, bench (all legs), launch list, ncu --set full of the new X-Pool kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 420 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xpool_score -s 4 -c 2 \
  -o gpurun_out/prof_xpool -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_xpool.log 2>&1
ncu -i gpurun_out/prof_xpool.ncu-rep --page raw --csv > gpurun_out/prof_xpool_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_xpool.ncu-rep --page source --csv > gpurun_out/prof_xpool_source.csv 2>/dev/null
find gpurun_out -name "prof_xpool.ncu-rep" -size +20M -delete
tail -22 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step']), 'launches', d['gpu_launches_per_step'])
print('gemm ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'exec frac', r['executed_frac'], 'xpool', r['xpool']['kernel_ms_per_step'], r['xpool']['executed_frac'])
print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'])
P
tail -3 gpurun_out/bench.err
