#!/bin/bash
# round 2, session 5, last call: the whole gpu suite + smoke + the bench line on the final commit
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), round(d['e2e']['h2d_bound_ms'],2), d['e2e']['chunk']), 'launches', d['gpu_launches_per_step'])
print('gemm ms', round(r['kernel_ms_per_step'],3), 'frac', round(r['frac'],3), 'xpool', round(r['xpool']['kernel_ms_per_step'],3), 'rank', r['other_families_ms_per_step'])
print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1), d['clocks'])
P
tail -2 gpurun_out/bench.err
