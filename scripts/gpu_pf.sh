for pf in 0 2; do
echo "== L2 prefetch mode $pf"
MADE_GEMM_L2_PREFETCH=$pf timeout 120 python scripts/diag_gemm_pair.py 2>&1 | grep -E "M=303104|M=160000|N=768"
MADE_GEMM_L2_PREFETCH=$pf timeout 200 python bench.py --no-e2e --no-cpu-baseline --detect-topk 0 --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench ms', d['ms_per_step'], 'gemm ms', d['roofline']['kernel_ms_per_step'])"
done
MADE_GEMM_L2_PREFETCH=2 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tcgen05 or encoders" 2>&1 | tail -2
