"""Resident gallery index for retrieval at scale (BASELINE.json configs[4]: 4k -> 1M tracks sharded over the GPUs).

The reference scores the whole gallery in one shot and materialises sim[N_v, N_m] (and, before it, the
[N_m, N_v, 256] pooled tensor) on the CPU (test-MaDe.py:386-413).  At 1M tracks that matrix alone is 16 GB in
fp64, so here the gallery lives on the GPU as the ENCODED operands of its shard (219 KB per
track -> 27 GB per GPU for 1M tracks on 8 GPUs) and a query batch streams over it in score chunks:

    for every chunk of `score_chunk` tracks:
        single, dual [N_v, chunk]  <- fused X-Pool scoring + tensor-core cosine      (made_xpool_score, made_cosine_sim)
        count of ids ahead of the ground truth, exact top-k of the chunk               (made_rank_topk)
        running top-k <- merge(running, chunk top-k)                                   (made_topk_merge)

Only two [N_v, chunk] fp32 tiles and the [N_v, 2k] candidate lists exist at any time.  Ranks are exact: the
ground-truth score of a query is computed first, by scoring it against its paired track alone through the same
kernels (pair scores do not depend on what else is in a launch, so this is bit-identical to the main pass).
Sharding over ranks: `ShardedIndex` all-gathers the query embeddings, every rank searches its shard, and the
packed [count | k columns | k scores] rows travel to the rank that owns each query (one all_to_all, as in
`parallel.ShardedEvaluator`).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib, ops
from . import config as cfg
from .pipeline import GalleryEvaluator


class GalleryIndex:
    def __init__(self, ev: GalleryEvaluator, capacity: int, score_chunk: int = 8192, col_offset: int = 0):
        """capacity: tracks this shard will hold; col_offset: global column of local track 0."""
        self.ev, self.eng, self.dev = ev, ev.eng, ev.dev
        self.capacity, self.score_chunk, self.col_offset = int(capacity), (int(score_chunk) + 3) // 4 * 4, int(col_offset)
        self.gal = ev.new_gallery(self.capacity)
        self.gal["mask"] = torch.zeros((self.capacity, cfg.L_M), dtype=torch.float32, device=self.dev)
        self.n = 0

    @property
    def bytes_per_track(self) -> int:
        g = self.gal
        return sum(g[k].numel() * g[k].element_size() for k in ("seq", "pooled", "kz", "gram", "bits", "mask")) // max(self.capacity, 1)

    @torch.no_grad()
    def add(self, segment_feats: torch.Tensor, segment_mask: torch.Tensor) -> None:
        """Encode a batch of tracks (device or pinned-host features, the encoder's input schema) and append their
        operands to the resident shard."""
        b = segment_feats.shape[0]
        if self.n + b > self.capacity:
            raise ValueError(f"index capacity {self.capacity} exceeded")
        L, s, e = cfg.L_M, self.n, self.n + b
        mask_d = segment_mask.to(self.dev, non_blocking=True).to(torch.float32)
        self.gal["mask"][s:e] = mask_d
        self.eng.encode(_lib.MUSIC, segment_feats.to(self.dev, non_blocking=True), mask_d, want_f32=False,
                        out=(self.gal["seq"][s:e], self.gal["pooled"][s:e]))
        self.eng.gallery_prepare(self.gal["seq"][s:e], mask_d,
                                 out=(self.gal["kz"][s * L:e * L], self.gal["gram"][s * L:e * L], self.gal["bits"][s:e]))
        self.n = e

    def _subgallery(self, idx: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Operands of the tracks `idx` (local indices) gathered into a small gallery."""
        L, g = cfg.L_M, self.gal
        rows = (idx.long()[:, None] * L + torch.arange(L, device=self.dev)[None, :]).reshape(-1)
        return dict(kz=g["kz"][rows], gram=g["gram"][rows], bits=g["bits"][idx.long()], pooled=g["pooled"][idx.long()])

    @torch.no_grad()
    def gt_scores(self, video_feats: torch.Tensor, gt_col: torch.Tensor, qprep=None, pair_chunk: int = 256) -> torch.Tensor:
        """fp64 score of every query against its paired track; -inf where the pair lives on another shard.
        gt_col [N_v] are GLOBAL columns.  Queries are grouped `pair_chunk` at a time and each group is scored
        against just its own paired tracks (the diagonal of a [chunk, chunk] launch)."""
        n_q = video_feats.shape[0]
        out = torch.full((n_q,), float("-inf"), dtype=torch.float64, device=self.dev)
        loc = gt_col.to(self.dev).long() - self.col_offset
        mine = torch.nonzero((loc >= 0) & (loc < self.n)).reshape(-1)      # host sync: planning step of a search
        if mine.numel() == 0:
            return out
        if qprep is None:
            qprep = self.eng.query_prepare(video_feats)
        q, vhat = qprep
        # 16-byte aligned rows whatever the group size: the cosine then takes the same (tcgen05) route as the main pass
        dual_buf = torch.empty((pair_chunk, (pair_chunk + 3) // 4 * 4), dtype=torch.float32, device=self.dev)
        for s in range(0, mine.numel(), pair_chunk):
            qi = mine[s:s + pair_chunk]
            sub = self._subgallery(loc[qi])
            single = self.eng.xpool_score(q[qi].contiguous(), vhat[qi].contiguous(), sub["kz"], sub["gram"], sub["bits"])
            dual = ops.cal_distance(video_feats[qi].contiguous(), sub["pooled"], out=dual_buf[:qi.numel()])
            out[qi] = single.diagonal().double() + dual[:, :qi.numel()].diagonal().double()
        return out

    @torch.no_grad()
    def search(self, video_feats: torch.Tensor, k: int, gt_score: Optional[torch.Tensor] = None, qprep=None):
        """→ dict(topk_idx [N_v,k] int32 GLOBAL columns, topk_score [N_v,k] f64, count [N_v] int32 = local tracks
        whose score beats gt_score (None without it)).  Music ids are taken as distinct (one column per id)."""
        n_q = video_feats.shape[0]
        if qprep is None:
            qprep = self.eng.query_prepare(video_feats)
        q, vhat = qprep
        L, g, C = cfg.L_M, self.gal, self.score_chunk
        single = torch.empty((n_q, min(C, (max(self.n, 1) + 3) // 4 * 4)), dtype=torch.float32, device=self.dev)
        dual = torch.empty_like(single)
        run_i = run_s = None
        count = torch.zeros(n_q, dtype=torch.int32, device=self.dev) if gt_score is not None else None
        self.pairs_scored = 0
        for s in range(0, self.n, C):
            e = min(self.n, s + C)
            w = e - s
            self.eng.xpool_score(q, vhat, g["kz"][s * L:e * L], g["gram"][s * L:e * L], g["bits"][s:e], out=single, col_offset=0)
            ops.cal_distance(video_feats, g["pooled"][s:e], out=dual, col_offset=0)
            r = ops.rank_topk(single, dual, None, None, k=min(k, w), col_offset=self.col_offset + s,
                              gt_score_in=gt_score, n_cols=w)
            self.pairs_scored += n_q * w
            if count is not None:
                count += r["rank"]
            ci, cs = r["topk_idx"], r["topk_score"]
            if ci.shape[1] < k:      # a chunk narrower than k: pad with empty candidates
                pad = k - ci.shape[1]
                ci = torch.cat([ci, torch.full((n_q, pad), -1, dtype=torch.int32, device=self.dev)], 1)
                cs = torch.cat([cs, torch.full((n_q, pad), float("-inf"), dtype=torch.float64, device=self.dev)], 1)
            if run_i is None:
                run_i, run_s = ci, cs
            else:
                run_i, run_s = ops.topk_merge(torch.cat([run_s, cs], 1), torch.cat([run_i, ci], 1), k)
        if run_i is None:
            run_i = torch.full((n_q, k), -1, dtype=torch.int32, device=self.dev)
            run_s = torch.full((n_q, k), float("-inf"), dtype=torch.float64, device=self.dev)
        return dict(topk_idx=run_i, topk_score=run_s, count=count)


class ShardedIndex:
    """The gallery index sharded over the ranks of a torch.distributed group (one process per GPU)."""

    def __init__(self, index: GalleryIndex, rank: int, world: int, group=None):
        self.index, self.rank, self.world, self.group = index, rank, world, group

    @torch.no_grad()
    def search(self, video_feats_local: torch.Tensor, k: int, gt_col: Optional[torch.Tensor] = None, q_sizes=None):
        """video_feats_local [n_local, 256]: THIS rank's query embeddings; gt_col [N_v] GLOBAL columns of every
        query (optional: turns on exact ranks).  → results for this rank's queries: dict(rank, topk_idx, topk_score)."""
        import torch.distributed as dist
        W, dev, idx = self.world, self.index.dev, self.index
        nl = video_feats_local.shape[0]
        if q_sizes is None:
            q_sizes = [nl] * W
        if W > 1:
            outs = [torch.empty((s, cfg.D_MODEL), dtype=torch.float32, device=dev) for s in q_sizes]
            dist.all_gather(outs, video_feats_local.contiguous(), group=self.group)
            vf = torch.cat(outs, 0)
        else:
            vf = video_feats_local
        qprep = idx.eng.query_prepare(vf)
        gt_score = None
        if gt_col is not None:
            gt_score = idx.gt_scores(vf, gt_col, qprep)
            if W > 1:
                dist.all_reduce(gt_score, op=dist.ReduceOp.MAX, group=self.group)
        r = idx.search(vf, k, gt_score, qprep)
        if W == 1:
            return dict(rank=r["count"], topk_idx=r["topk_idx"], topk_score=r["topk_score"])
        cnt = r["count"] if r["count"] is not None else torch.zeros(vf.shape[0], dtype=torch.int32, device=dev)
        packed = torch.cat([cnt.reshape(-1, 1), r["topk_idx"], r["topk_score"].view(torch.int32)], 1)
        recv = torch.empty((W, nl, 1 + 3 * k), dtype=torch.int32, device=dev)
        dist.all_to_all_single(recv.view(W * nl, -1), packed, [nl] * W, list(q_sizes), group=self.group)
        rank_cnt = recv[:, :, 0].sum(0, dtype=torch.int32)
        cand_i = recv[:, :, 1:1 + k].permute(1, 0, 2).reshape(nl, W * k)
        cand_s = recv[:, :, 1 + k:].permute(1, 0, 2).reshape(nl, W * 2 * k).contiguous().view(torch.float64)
        topk_idx, topk_score = ops.topk_merge(cand_s, cand_i, k)
        return dict(rank=rank_cnt if gt_col is not None else None, topk_idx=topk_idx, topk_score=topk_score)
