"""On-hardware parity of the sharded path: 2 NCCL ranks (one process per GPU) must return the single-GPU path's
ranks, top-k, scores and spans bit for bit — with distinct music ids and with repeated ids (dedup ranks,
utils/util_test.py:46-60).  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NQ, NM, K = 300, 520, 50


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ids_with_repeats():
    ids = [f"m{i}" for i in range(NM)]
    rng = np.random.default_rng(3)
    for c in rng.choice(np.arange(1, NM), size=60, replace=False):
        ids[c] = ids[c - 1]                 # runs of repeated ids, some across the uniform shard cut
    ids[NM // 2] = ids[NM // 2 - 1]
    return ids


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from mgsv_b200 import ops, parallel, synth
    from mgsv_b200.engine import Engine
    from mgsv_b200.parallel import ShardedEvaluator, shard_bounds
    from mgsv_b200.pipeline import GalleryEvaluator
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        eng = Engine(dev)
        eng.load_state_dict(synth.make_state_dict(0))
        v, m, _ = synth.make_eval_set(NQ, NM, synth.BASE_SEED + 41)
        vkeys, mkeys = ("frame_feats", "frame_mask"), ("segment_feats", "segment_mask", "gt_moment", "m_duration")
        full_v = {k: v[k].to(dev) for k in vkeys}
        full_m = {k: m[k].to(dev) for k in mkeys}
        # pair query i with a track on the far side of the gallery: forces cross-shard pairs
        gt_col = torch.tensor([(NM - 1 - 2 * i) % NM for i in range(NQ)], dtype=torch.int32)
        ev = GalleryEvaluator(eng, k=K, music_chunk=200, video_chunk=128)
        keys = ("rank", "topk_idx", "topk_score", "pred_st", "pred_ed", "iou", "score")
        result = {}
        for name, ids in (("distinct", None), ("repeats", _ids_with_repeats())):
            prev = gt_last = None
            if ids is not None:
                prev_np, _, _ = ops.dedup_tables(ids)
                last = {mid: c for c, mid in enumerate(ids)}
                gt_last = torch.tensor([last[ids[int(g)]] for g in gt_col], dtype=torch.int32)
                prev = torch.from_numpy(prev_np).to(dev)
            one = ev.run(full_v, full_m, (gt_col if gt_last is None else gt_last).to(dev), prev_same=prev)
            # the moment is detected on the PAIRED track (gt_col); with repeated ids the rank's GT column is the last
            # column of the id, so run the detection pairing separately for the single-GPU reference
            one_det = ev.run(full_v, full_m, gt_col.to(dev)) if ids is not None else one
            one = {k: (one if k in ("rank", "topk_idx", "topk_score") else one_det)[k].clone() for k in keys}
            bounds = parallel.plan_track_shards(ids, NM, world)
            q0, q1 = shard_bounds(NQ, rank, world)
            m0, m1 = bounds[rank]
            sh = ShardedEvaluator(ev, rank, world, track_bounds=bounds)
            out = sh.run({k: t[q0:q1].contiguous() for k, t in full_v.items()},
                         {k: t[m0:m1].contiguous() for k, t in full_m.items()}, gt_col, NQ, NM, music_ids=ids)
            torch.cuda.synchronize()
            bad = [k for k in ("rank", "topk_idx", "topk_score") if not torch.equal(out[k], one[k])]
            bad += [k for k in ("pred_st", "pred_ed", "iou", "score") if not torch.equal(out[k], one[k][q0:q1])]
            own = sh.run({k: t[q0:q1].contiguous() for k, t in full_v.items()},
                         {k: t[m0:m1].contiguous() for k, t in full_m.items()}, gt_col, NQ, NM, music_ids=ids,
                         gather_results=False)
            torch.cuda.synchronize()
            bad += [k + "(own)" for k in ("rank", "topk_idx", "topk_score") if not torch.equal(own[k], one[k][q0:q1])]
            result[name] = bad
        q.put((rank, result))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_two_gpus_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _q
    import time as _t
    deadline = _t.time() + 500
    while len(res) < world and _t.time() < deadline:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a rank died"
    for p in procs:
        p.join(60)
    assert len(res) == world
    for rank, result in res:
        for name, bad in result.items():
            assert bad == [], f"rank {rank}, {name} ids: sharded path differs from the single-GPU path in {bad}"
