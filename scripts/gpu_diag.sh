#!/bin/bash
# Run each diagnostic stage in its own process under a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/diag_gpu.txt
for s in "${@:-span rank gemm attn encode xpool detr}"; do
  for stage in $s; do
    echo "===== $stage =====" | tee -a gpurun_out/diag.log
    timeout 240 python tests/tools/gpu_diag.py $stage 2>&1 | tail -60 | tee -a gpurun_out/diag.log
    echo "exit: ${PIPESTATUS[0]}" | tee -a gpurun_out/diag.log
  done
done
