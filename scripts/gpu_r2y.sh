#!/bin/bash
# round 2, session 5, final: gpu suite, smoke, micro bench, bench (all legs), launch list with DRAM bytes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_final.json 2> gpurun_out/mb.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --detect-topk 0 \
  > gpurun_out/bench_under_ncu.log 2>&1
python scripts/step_dram.py gpurun_out/launches.csv > gpurun_out/step_dram.json 2> gpurun_out/step_dram.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), round(d['e2e']['h2d_bound_ms'],2)), 'launches', d['gpu_launches_per_step'])
print('gemm ms', round(r['kernel_ms_per_step'],3), 'frac', round(r['frac'],3), 'exec frac', round(r['executed_frac'],3), 'xpool', round(r['xpool']['kernel_ms_per_step'],3), round(r['xpool']['executed_frac'],3), 'serial', round(r['serial_step_ms'],3))
print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1), d['clocks'], 'rtd', d['retrieve_then_detect']['ms_per_step'])
for r in json.load(open("gpurun_out/micro_bench_final.json"))["kernels"]:
    print(f"{r['kernel']:40s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
s=json.load(open('gpurun_out/step_dram.json')); s.pop('per_kernel'); print(s)
P
tail -2 gpurun_out/bench.err
