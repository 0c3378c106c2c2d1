#!/bin/bash
# round 2, session 5: rank kernel after the single-barrier count + deferred key fetch: tests, racecheck, micro bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "rank or topk or recall or sharded or index or full_size" 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 \
  -k "(rank_and_topk or topk_value or topk_group) and not full_size" > gpurun_out/r02_w_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_w_racecheck.log
tail -3 gpurun_out/r02_w_racecheck.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_new.json 2> gpurun_out/mb_new.err
python - <<'P'
import json
d = json.load(open("gpurun_out/micro_bench_new.json"))
for r in d["kernels"]:
    if "rank" in r["kernel"]:
        print(f"{r['kernel']:40s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
P
