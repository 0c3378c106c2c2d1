#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_ffn.py > gpurun_out/diag_ffn.log 2>&1
cat gpurun_out/diag_ffn.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "ffn or temporal or full_size or index or detr" 2>&1 | tail -15
