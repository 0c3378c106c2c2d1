#!/bin/bash
# round 2, call p: racecheck, product kernels and the opt-in CTA-pair form apart
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 \
  -k "(tcgen05 or span_fast or span_kernels_bit_exact_vs_oracle or fused_ffn) and not weight_stationary and not cta_pair" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck.log
tail -4 gpurun_out/r02_racecheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 \
  -k "cta_pair" > gpurun_out/r02_racecheck_pair.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck_pair.log
grep -c "Race reported" gpurun_out/r02_racecheck_pair.log; grep "Race reported\|access at" gpurun_out/r02_racecheck_pair.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head
tail -3 gpurun_out/r02_racecheck_pair.log
