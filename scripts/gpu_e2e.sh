for ch in 500 1000; do
timeout 300 python bench.py --no-cpu-baseline --detect-topk 0 --steps 20 --warmup 3 --e2e-chunk $ch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('e2e chunk $ch: e2e ms', round(e['ms_per_step'],2), 'median', round(e['ms_per_step_median'],2), 'bound', round(e['h2d_bound_ms'],2), 'dev ms', round(d['ms_per_step'],3))"
done
