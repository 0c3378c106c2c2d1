#!/bin/bash
# round 2, call o: racecheck of the GEMM forms after the CTA-pair barrier fix, smoke, micro bench (with CUDA-graph replays of
# configs[2]), bench (all legs)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 \
  -k "(tcgen05 or span_fast or span_kernels_bit_exact_vs_oracle) and not weight_stationary" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck.log
tail -4 gpurun_out/r02_racecheck.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_r02o.json 2> gpurun_out/mb.err
tail -3 gpurun_out/mb.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), round(d['e2e']['h2d_bound_ms'],2)), 'launches', d['gpu_launches_per_step'])
print('gemm ms', round(r['kernel_ms_per_step'],3), 'frac', round(r['frac'],3), 'exec frac', round(r['executed_frac'],3), 'xpool', round(r['xpool']['kernel_ms_per_step'],3), round(r['xpool']['executed_frac'],3), 'serial', round(r['serial_step_ms'],3))
print('hbm view', r.get('hbm_view') and {k: r['hbm_view'][k] for k in ('achieved_gbs','frac','whole_step_frac')})
print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1), d['clocks'], 'rtd', d['retrieve_then_detect']['ms_per_step'])
for r in json.load(open("gpurun_out/micro_bench_r02o.json"))["kernels"]:
    print(f"{r['kernel']:34s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
P
tail -2 gpurun_out/bench.err
