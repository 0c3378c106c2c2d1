#!/bin/bash
# round 2, session 4, call A: gpu suite on the new rank / span kernels, micro bench A/B (new default vs. switches off), short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_new.json 2> gpurun_out/mb_new.err
MADE_RANK_GROUP_MAXIMA=0 MADE_SPAN_PACKED=0 timeout 300 python scripts/micro_bench.py > gpurun_out/micro_bench_old.json 2> gpurun_out/mb_old.err
python - <<'P'
import json
for tag in ("new", "old"):
    try:
        d = json.load(open(f"gpurun_out/micro_bench_{tag}.json"))
    except Exception as e:
        print(tag, "failed", e); continue
    for r in d["kernels"]:
        if "16384" in r["kernel"] or "rank" in r["kernel"] or "postproc" in r["kernel"]:
            print(f"{tag} {r['kernel']:34s} {r['ms']:.4f} ms {r['gbs']:.0f} GB/s {100*r['frac_of_hbm_peak']:.1f} %")
P
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench exit $?" >> gpurun_out/bench_a.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_a.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2)), 'launches', d['gpu_launches_per_step'])
print('gemm ms', round(r['kernel_ms_per_step'],3), 'xpool', round(r['xpool']['kernel_ms_per_step'],3), 'serial', round(r['serial_step_ms'],3), r.get('other_families_ms_per_step'))
P
tail -2 gpurun_out/bench_a.err
