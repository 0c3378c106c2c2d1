"""world_size-2 gloo test of the gallery-sharding host logic (mgsv_b200/parallel.py) on CPU.

The CUDA kernels cannot run here, so the per-rank compute is a CPU stand-in built from the oracle
with the same interface as GalleryEvaluator / ops.rank_topk / ops.topk_merge; what is under test is
the sharding, the five exchange steps and the candidate merge, which must reproduce the
single-process result exactly.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgsv_b200 import ops, parallel, synth
from mgsv_b200.parallel import ShardedEvaluator, shard_bounds
from oracle import made_oracle as O

NQ, NM, K = 12, 20, 5


def _cpu_rank_topk(single, dual, gt_col=None, prev_same=None, k=0, col_offset=0, gt_score_in=None, n_cols=None):
    """CPU stand-in of ops.rank_topk with its dedup semantics (made_b200.h made_rank_topk): prev_same chains
    group the columns of one music id; the GT score is the best column of the chain that ends at gt_col; the
    rank counts the chains whose best column beats it."""
    tot = single.double() + dual.double()
    n, m = tot.shape
    out = dict(topk_idx=None, topk_score=None, rank=None, gt_score=None)
    group = np.arange(m)
    if prev_same is not None:
        prev = prev_same.numpy()
        for c in range(m):
            if prev[c] >= 0:
                group[c] = group[prev[c]]
    if gt_col is not None or gt_score_in is not None:
        if gt_score_in is not None:
            gs = gt_score_in.clone()
        else:
            gs = torch.full((n,), float("-inf"), dtype=torch.float64)
            for i in range(n):
                if gt_col[i] >= 0:
                    gs[i] = tot[i, torch.from_numpy(group == group[int(gt_col[i])])].max()
        out["gt_score"] = gs
        best = torch.full((n, m), float("-inf"), dtype=torch.float64)
        for g in np.unique(group):
            best[:, g] = tot[:, torch.from_numpy(group == g)].max(1).values
        out["rank"] = (best > gs[:, None]).sum(1).to(torch.int32)
    if k > 0:
        kk = min(k, tot.shape[1])
        order = np.lexsort((np.broadcast_to(np.arange(tot.shape[1]), tot.shape), -tot.numpy()), axis=1)[:, :kk]
        idx = torch.full((n, k), -1, dtype=torch.int32)
        sc = torch.full((n, k), float("-inf"), dtype=torch.float64)
        idx[:, :kk] = torch.from_numpy(order.astype(np.int32)) + col_offset
        sc[:, :kk] = torch.gather(tot, 1, torch.from_numpy(order.astype(np.int64)))
        out["topk_idx"], out["topk_score"] = idx, sc
    return out


def _cpu_topk_merge(cs, ci, k):
    cs = cs.clone()
    cs[ci < 0] = float("-inf")
    order = np.lexsort((ci.numpy(), -cs.numpy()), axis=1)[:, :k]
    o = torch.from_numpy(order.astype(np.int64))
    return torch.gather(ci, 1, o), torch.gather(cs, 1, o)


class _CpuEvaluator:
    """Oracle-backed stand-in with GalleryEvaluator's interface."""

    def __init__(self, sd, k):
        self.sd, self.k, self.dev, self.launches = sd, k, torch.device("cpu"), 0

    def _count(self, *a, **k):
        pass

    def encode_queries(self, feats, mask):
        seq, pooled = O.encode_video(self.sd, feats, mask)
        return seq, pooled, mask

    def encode_gallery(self, feats, mask, on_chunk=None):
        seq, pooled = O.encode_music(self.sd, feats, mask)
        return dict(seq=seq, pooled=pooled, mask=mask)

    def score(self, video_feats, gal):
        single, dual, _ = O.gallery_similarity(self.sd, video_feats, gal["pooled"], gal["seq"], gal["mask"])
        return single, dual

    def detect(self, frame_seq, frame_mask, gal, video_feats, track_idx, gt_moment, m_duration):
        idx = track_idx.long()
        seg, smask = gal["seq"][idx].float(), gal["mask"][idx]
        src = torch.cat([frame_seq, seg], 1)
        mask = torch.cat([frame_mask, smask], 1)
        hs, _ = O.detr_forward(self.sd, src, mask, O.position_embedding_sine(mask), video_feats.unsqueeze(1))
        om = O.calc_output(self.sd, hs, frame_seq)
        st, ed, sc = O.moment_postproc(om["pred_logits"], om["pred_spans"])
        iou = O.detr_iou(st, ed, gt_moment.reshape(-1, 1, 2), m_duration)
        return dict(pred_st=st, pred_ed=ed, score=sc, iou=iou)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dup_ids():
    """20 columns, 14 distinct ids: ids 3, 9 (straddling the uniform cut at column 10) and 15 repeat."""
    ids = [f"m{i}" for i in range(NM)]
    ids[4] = ids[3]
    ids[10] = ids[9]
    ids[11] = ids[9]
    ids[16] = ids[15]
    ids[17] = ids[15]
    ids[18] = ids[15]
    return ids


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops.rank_topk, ops.topk_merge = _cpu_rank_topk, _cpu_topk_merge
        parallel.ops = ops
        sd = synth.make_state_dict(0)
        v, m, ids = synth.make_eval_set(NQ, NM, 77)
        # pair query i with track (NM - 1 - i): forces cross-shard pairs
        gt_col = torch.tensor([NM - 1 - i for i in range(NQ)], dtype=torch.int32)
        q0, q1 = shard_bounds(NQ, rank, world)
        m0, m1 = shard_bounds(NM, rank, world)
        videos = {k: v[k][q0:q1] for k in ("frame_feats", "frame_mask")}
        tracks = {k: m[k][m0:m1] for k in ("segment_feats", "segment_mask", "gt_moment", "m_duration")}
        sh = ShardedEvaluator(_CpuEvaluator(sd, K), rank, world)
        out = sh.run(videos, tracks, gt_col, NQ, NM)
        # the same job with repeated music ids: shards must not cut through an id, ranks are dedup ranks
        ids2 = _dup_ids()
        bounds = parallel.plan_track_shards(ids2, NM, world)
        b0, b1 = bounds[rank]
        tracks2 = {k: m[k][b0:b1] for k in ("segment_feats", "segment_mask", "gt_moment", "m_duration")}
        sh2 = ShardedEvaluator(_CpuEvaluator(sd, K), rank, world, track_bounds=bounds)
        out2 = sh2.run(videos, tracks2, gt_col, NQ, NM, music_ids=ids2, gather_results=False)
        q.put((rank, out["rank"].numpy(), out["topk_idx"].numpy(), out["topk_score"].numpy(),
               out["iou"].numpy(), out["pred_st"].numpy(), out["q_range"], out2["rank"].numpy(), bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_equals_single_process():
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
    deadline = _t.time() + 400
    while len(res) < world and _t.time() < deadline:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a rank died"
    assert len(res) == world
    res.sort(key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process oracle on the same job
    sd = synth.make_state_dict(0)
    v, m, ids = synth.make_eval_set(NQ, NM, 77)
    gt_col = np.array([NM - 1 - i for i in range(NQ)])
    fo, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
    so, mf = O.encode_music(sd, m["segment_feats"], m["segment_mask"])
    single, dual, total = O.gallery_similarity(sd, vf, mf, so, m["segment_mask"])
    gs = total[np.arange(NQ), gt_col]
    ref_rank = (total > gs[:, None]).sum(1)
    ref_top = np.lexsort((np.broadcast_to(np.arange(NM), total.shape), -total), axis=1)[:, :K]
    src = torch.cat([fo, so[gt_col]], 1)
    mask = torch.cat([v["frame_mask"], m["segment_mask"][gt_col]], 1)
    hs, _ = O.detr_forward(sd, src, mask, O.position_embedding_sine(mask), vf.unsqueeze(1))
    om = O.calc_output(sd, hs, fo)
    st, ed, sc = O.moment_postproc(om["pred_logits"], om["pred_spans"])
    ref_iou = O.detr_iou(st, ed, m["gt_moment"][gt_col], m["m_duration"][gt_col]).numpy()
    # dedup ranks of the job with repeated ids (Recall_metrics(dedup=True) through the oracle)
    ids2 = _dup_ids()
    _, ref_ind2, _ = O.recall_metrics(total, ids2, gt_col)
    assert parallel.plan_track_shards(ids2, NM, 2) == [(0, 12), (12, 20)]       # the cut moved past id m9
    assert parallel.plan_track_shards(["a", "b", "a", "b"], 4, 2) == [(0, 4), (4, 4)]   # interleaved ids: one shard
    assert list(parallel.order_tracks_by_id(["a", "b", "a", "b"])) == [0, 2, 1, 3]
    for rank, rk, ti, ts, iou, pst, (q0, q1), rk2, bounds in res:
        assert bounds == [(0, 12), (12, 20)]
        assert np.array_equal(rk2, ref_ind2[q0:q1])               # owner-only results (gather_results=False)
        assert np.array_equal(rk, ref_rank)                       # ranks need both shards
        assert np.array_equal(ti, ref_top)
        np.testing.assert_allclose(ts, np.take_along_axis(total, ref_top, 1), rtol=0, atol=1e-6)  # fp32 BLAS blocking differs per shard size
        np.testing.assert_allclose(iou, ref_iou[q0:q1], atol=1e-5)  # detection shards by query
        np.testing.assert_allclose(pst, st.numpy()[q0:q1], atol=1e-3)
