#!/usr/bin/env python
"""BASELINE.json configs[4]: synthetic gallery scale sweep, 4k -> 1M music tracks sharded over the GPUs of one box,
2000 queries, top-100 with exact ranks, NCCL candidate exchange.

    python scripts/gallery_sweep.py --tracks 4096,65536 [--steps 5]                     # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/gallery_sweep.py \
        --tracks 4096,65536,262144,1048576

Per gallery size: the shard is built ON the device (synthetic AST features drawn per 1000-track batch, encoded,
X-Pool operands kept resident: 219 KB per track), then `--steps` searches of the same 2000-query batch are timed
with CUDA events (barrier + synchronize on both sides, max over ranks).  One JSON line per size on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

F_XPOOL_PAIR, F_XPOOL_PAIR_EXEC = 360_960.0, 2.0 * (96 * 256 + 96 * 112 + 96 * 256)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", default="4096,16384,65536")
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--score-chunk", type=int, default=8192)
    args = ap.parse_args()
    import torch.distributed as dist
    from mgsv_b200 import _lib, synth
    from mgsv_b200 import config as cfg
    from mgsv_b200.engine import Engine
    from mgsv_b200.index import GalleryIndex, ShardedIndex
    from mgsv_b200.parallel import shard_bounds
    from mgsv_b200.pipeline import GalleryEvaluator

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")):
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)

    eng = Engine(dev)
    eng.load_state_dict(synth.make_state_dict(0))
    ev = GalleryEvaluator(eng, k=args.k)
    nq = args.queries
    q0, q1 = shard_bounds(nq, rank, world)
    v = synth.make_videos(nq, synth.BASE_SEED + 2)
    _, vf_all, _ = ev.encode_queries(v["frame_feats"].to(dev), v["frame_mask"].to(dev))
    vf_local = vf_all[q0:q1].contiguous()
    q_sizes = [shard_bounds(nq, r, world)[1] - shard_bounds(nq, r, world)[0] for r in range(world)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for nm in [int(x) for x in args.tracks.split(",")]:
        m0, m1 = shard_bounds(nm, rank, world)
        n_loc = m1 - m0
        idx = GalleryIndex(ev, capacity=n_loc, score_chunk=args.score_chunk, col_offset=m0)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        rng = np.random.Generator(np.random.PCG64([77, rank, nm]))
        barrier()
        t0 = time.perf_counter()
        for s in range(0, n_loc, 1000):
            b = min(1000, n_loc - s)
            _, m_dur, _, _, _, n_seg = synth._lengths(rng, b)
            mask = (torch.arange(cfg.L_M)[None, :] < torch.from_numpy(n_seg)[:, None]).float().to(dev)
            feats = torch.randn((b, cfg.L_M, cfg.D_AST), generator=gen, device=dev) * mask[:, :, None]
            idx.add(feats, mask)
        barrier()
        build_s = time.perf_counter() - t0
        sh = ShardedIndex(idx, rank, world)
        gt_col = (torch.arange(nq, dtype=torch.int64) * 7919) % nm          # a paired track per query, spread over shards
        for _ in range(2):
            out = sh.search(vf_local, args.k, gt_col=gt_col, q_sizes=q_sizes)
        barrier()
        _lib.prof_collect()
        _lib.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = sh.search(vf_local, args.k, gt_col=gt_col, q_sizes=q_sizes)
        e1.record()
        barrier()
        _lib.prof_enable(False)
        prof = _lib.prof_collect()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        # sanity: the paired track of a query must be ranked (count >= 0) and, if it is in the top-k, at that position
        rk = out["rank"]
        hit = (out["topk_idx"].long() == gt_col[q0:q1].to(dev)[:, None])
        pos = torch.where(hit.any(1), hit.float().argmax(1), torch.full_like(rk.long(), -1))
        ok = bool(((pos < 0) | (pos == rk.long())).all())
        if rank == 0:
            pairs_gpu = float(nq) * n_loc
            xp_ms = prof["xpool"][0] / args.steps
            line = {
                "metric": "queries/sec (match, top-100 + exact rank) vs gallery size", "n_gpus": world, "n_tracks": nm,
                "tracks_per_gpu": n_loc, "n_queries": nq, "k": args.k, "ms_per_query_batch": ms,
                "value": nq / (ms / 1e3), "unit": "queries/s", "pairs_per_s": nq * float(nm) / (ms / 1e3),
                "index_build_s": build_s, "index_build_tracks_per_s_per_gpu": n_loc / build_s,
                "resident_bytes_per_track": idx.bytes_per_track, "resident_gb_per_gpu": idx.bytes_per_track * n_loc / 1e9,
                "score_chunk": args.score_chunk, "steps": args.steps,
                "xpool_ms_per_batch": xp_ms, "xpool_share": xp_ms / ms,
                "xpool_tflops_algorithmic": F_XPOOL_PAIR * pairs_gpu / (xp_ms / 1e3) / 1e12 if xp_ms > 0 else None,
                "xpool_frac_of_peak_executed": F_XPOOL_PAIR_EXEC * pairs_gpu / (xp_ms / 1e3) / 1e12 / peak_tf if xp_ms > 0 else None,
                "rank_topk_ms_per_batch": prof["rank"][0] / args.steps,
                "rank_topk_gbs": 8.0 * pairs_gpu / (prof["rank"][0] / args.steps / 1e3) / 1e9 if prof["rank"][0] > 0 else None,
                "rank_consistent_with_topk": ok, "scaling": "weak-in-gallery (queries fixed, tracks per GPU = n_tracks / n_gpus)",
                "data": "synthetic (features drawn on the device)", "precision": eng.precision,
            }
            print(json.dumps(line), flush=True)
        del idx, sh, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
