#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/diag_ffn.py > gpurun_out/diag_ffn.log 2>&1
timeout 300 python scripts/diag_stages.py > gpurun_out/diag_stages.log 2>&1
timeout 600 python tests/tools/gpu_window_error.py 96 256 > gpurun_out/window_error.log 2>&1
for i in 1 2 3; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err
done
timeout 300 python scripts/gallery_sweep.py --tracks 4096,32768 --steps 3 > gpurun_out/sweep_1gpu.jsonl 2> gpurun_out/sweep_1gpu.err
tail -12 gpurun_out/pytest_gpu.log; grep "debug=0\|debug=2" gpurun_out/diag_ffn.log; cat gpurun_out/diag_stages.log; grep split gpurun_out/window_error.log
for i in 1 2 3; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$i.json')); r=d['roofline']
print('bench $i', d['ms_per_step'], 'gemm ms', r['kernel_ms_per_step'], 'ffn', r['ffn_fused_ms_per_step'], 'xpool', r['xpool']['kernel_ms_per_step'], 'frac', r['frac'])"; done
cat gpurun_out/sweep_1gpu.jsonl; tail -3 gpurun_out/sweep_1gpu.err
