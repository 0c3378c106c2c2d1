#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 300 -k "xpool or cfg1 or full_size or index or forward_vs or retrieve" 2>&1 | tail -30 > gpurun_out/pytest_xpool.log
cat gpurun_out/pytest_xpool.log
timeout 300 python scripts/diag_stages.py > gpurun_out/diag_stages.log 2>&1
cat gpurun_out/diag_stages.log
