#!/bin/bash
# configs[4]: gallery sweep on N GPUs (N = $1)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  scripts/gallery_sweep.py --tracks ${2:-4096,65536,262144,1048576} --steps 3 > gpurun_out/sweep_${N}gpu.jsonl 2> gpurun_out/sweep_${N}gpu.err
echo "exit $?" >> gpurun_out/sweep_${N}gpu.err
cat gpurun_out/sweep_${N}gpu.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['n_gpus'], d['n_tracks'], 'ms', round(d['ms_per_query_batch'],2), 'q/s', round(d['value']), 'pairs/s %.3g' % d['pairs_per_s'], 'xpool share', round(d['xpool_share'],2), 'exec frac', round(d['xpool_frac_of_peak_executed'],3), 'build tracks/s/gpu', round(d['index_build_tracks_per_s_per_gpu']), 'GB/gpu', round(d['resident_gb_per_gpu'],1), 'ok', d['rank_consistent_with_topk'])
"
tail -5 gpurun_out/sweep_${N}gpu.err
