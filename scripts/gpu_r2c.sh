#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_stages.py > gpurun_out/diag_fused.log 2>&1
MADE_FUSED_FFN=0 timeout 300 python scripts/diag_stages.py > gpurun_out/diag_unfused.log 2>&1
cat gpurun_out/diag_fused.log gpurun_out/diag_unfused.log
