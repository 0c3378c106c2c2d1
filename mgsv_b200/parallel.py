"""Gallery sharding across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Tracks are independent units, so the
gallery is split into contiguous shards; the only exchange steps are tiny:
  1. all_gather of the query embeddings [N_v/G, 256] (each rank encodes its slice of the queries),
  2. all_reduce(MAX) of the ground-truth scores [N_v] f64 (the GT track lives on one shard),
  3. all_reduce(SUM) of the "ids ahead of the GT" counts [N_v] i32,
  4. all_gather of the local top-k candidates [N_v, k] (score f64, global index i32) + merge kernel.
Moment detection shards by query; a query's paired track may live on another shard, so the encoded
segments of the paired tracks are exchanged with one all_gather of [N_v/G, 96, 256] fp16 slices.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from . import config as cfg
from .pipeline import GalleryEvaluator


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous split with the remainder spread over the first ranks."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def owner_of(col: torch.Tensor, n: int, world: int) -> torch.Tensor:
    """Rank that owns gallery column `col` under shard_bounds."""
    base, rem = divmod(n, world)
    big = (base + 1) * rem
    return torch.where(col < big, col // max(base + 1, 1), rem + (col - big) // max(base, 1))


def deliver_rows(send: torch.Tensor, sizes, rank: int, group=None) -> torch.Tensor:
    """Each row of `send` [sum(sizes), ...] is non-zero on exactly one rank (its owner) and zero
    elsewhere; rank r must end up with rows offs[r]:offs[r+1].  A SUM reduce-scatter does that in
    one collective (NCCL); backends without reduce_scatter (gloo, CPU tests) all_reduce and slice."""
    offs = [0]
    for s_ in sizes:
        offs.append(offs[-1] + s_)
    if dist.get_backend(group) == "nccl":
        recv = torch.empty((sizes[rank],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.reduce_scatter(recv, [send[offs[r]:offs[r + 1]] for r in range(len(sizes))], group=group)
        return recv
    buf = send.clone()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[offs[rank]:offs[rank + 1]].contiguous()


class ShardedEvaluator:
    """Strong-scaling evaluation of one (N_v queries x N_m tracks) job on `world` GPUs."""

    def __init__(self, ev: GalleryEvaluator, rank: int, world: int, group=None):
        self.ev, self.rank, self.world, self.group = ev, rank, world, group

    def _all_gather_cat(self, t: torch.Tensor, sizes) -> torch.Tensor:
        outs = [torch.empty((s,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for s in sizes]
        dist.all_gather(outs, t.contiguous(), group=self.group)
        return torch.cat(outs, 0)

    @torch.no_grad()
    def run(self, videos: Dict[str, torch.Tensor], tracks: Dict[str, torch.Tensor], gt_col: torch.Tensor,
            n_queries: int, n_tracks: int, on_host: bool = False):
        """`videos` holds THIS rank's slice of the queries, `tracks` THIS rank's gallery shard
        (features, masks, gt_moment, m_duration); gt_col [n_queries] are GLOBAL column indices."""
        ev, dev, W, R = self.ev, self.ev.dev, self.world, self.rank
        ev.launches = 0
        q0, q1 = shard_bounds(n_queries, R, W)
        m0, m1 = shard_bounds(n_tracks, R, W)
        q_sizes = [shard_bounds(n_queries, r, W)[1] - shard_bounds(n_queries, r, W)[0] for r in range(W)]
        frame_seq, vf_local, frame_mask = ev.encode_queries(videos["frame_feats"], videos["frame_mask"])
        gal = ev.encode_gallery(tracks["segment_feats"], tracks["segment_mask"])
        video_feats = self._all_gather_cat(vf_local, q_sizes)                      # exchange 1
        single, dual = ev.score(video_feats, gal)
        gt = gt_col.to(dev).to(torch.int32)
        local_gt = torch.where((gt >= m0) & (gt < m1), gt - m0, torch.full_like(gt, -1))
        r1 = ops.rank_topk(single, dual, local_gt, None, k=0)
        ev._count("rank")
        gt_score = r1["gt_score"]
        dist.all_reduce(gt_score, op=dist.ReduceOp.MAX, group=self.group)          # exchange 2
        r2 = ops.rank_topk(single, dual, None, None, k=ev.k, col_offset=m0, gt_score_in=gt_score)
        ev._count("rank")
        rank_cnt = r2["rank"]
        dist.all_reduce(rank_cnt, op=dist.ReduceOp.SUM, group=self.group)          # exchange 3
        cand_s = [torch.empty_like(r2["topk_score"]) for _ in range(W)]
        cand_i = [torch.empty_like(r2["topk_idx"]) for _ in range(W)]
        dist.all_gather(cand_s, r2["topk_score"], group=self.group)                # exchange 4
        dist.all_gather(cand_i, r2["topk_idx"], group=self.group)
        topk_idx, topk_score = ops.topk_merge(torch.cat(cand_s, 1), torch.cat(cand_i, 1), ev.k)
        ev.launches += 1
        # ---- detection for this rank's queries; fetch the paired tracks' encoded segments ----
        all_need = gt.long()                       # every rank knows the global pairing
        sel = (all_need >= m0) & (all_need < m1)
        loc = all_need[sel] - m0
        send = torch.zeros((n_queries, cfg.L_M, cfg.D_MODEL), dtype=gal["seq"].dtype, device=dev)
        send_mask = torch.zeros((n_queries, cfg.L_M), dtype=torch.float32, device=dev)
        send_meta = torch.zeros((n_queries, 3), dtype=torch.float32, device=dev)
        send[sel] = gal["seq"][loc]
        send_mask[sel] = gal["mask"][loc]
        gtm = tracks["gt_moment"].to(dev).reshape(-1, 2)
        send_meta[sel] = torch.cat([gtm[loc], tracks["m_duration"].to(dev)[loc].unsqueeze(1)], 1)
        recv = deliver_rows(send, q_sizes, R, self.group)                          # exchange 5
        recv_mask = deliver_rows(send_mask, q_sizes, R, self.group)
        recv_meta = deliver_rows(send_meta, q_sizes, R, self.group)
        pair = dict(seq=recv, mask=recv_mask)
        det = ev.detect(frame_seq, frame_mask, pair, vf_local,
                        torch.arange(q1 - q0, dtype=torch.int32, device=dev),
                        recv_meta[:, :2].contiguous(), recv_meta[:, 2].contiguous())
        return dict(rank=rank_cnt, topk_idx=topk_idx, topk_score=topk_score, q_range=(q0, q1), **det)
