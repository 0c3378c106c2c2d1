"""Gallery sharding across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Tracks are independent units, so the
gallery is split into contiguous shards; the only exchange steps are tiny:
  1. all_gather of the query embeddings [N_v/G, 256] (each rank encodes its slice of the queries),
  2. all_reduce(MAX) of the ground-truth scores [N_v] f64 (the GT track lives on one shard),
  3. all_reduce(SUM) of the "ids ahead of the GT" counts [N_v] i32,
  4. all_gather of the local top-k candidates [N_v, k] (score f64, global index i32) + merge kernel,
  5. all_to_all of the paired tracks' encoded segments (planned on the host, no device sync).
Moment detection shards by query; a query's paired track may live on another shard, so the encoded
segments of the paired tracks travel in one all_to_all of [*, 96, 256] fp16 rows (+ one of their
masks / ground-truth moments) whose split sizes are planned on the host from the pairing.
"""
from __future__ import annotations

import os
import time
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


class ShardedEvaluator:
    """Strong-scaling evaluation of one (N_v queries x N_m tracks) job on `world` GPUs."""

    def __init__(self, ev: GalleryEvaluator, rank: int, world: int, group=None):
        self.ev, self.rank, self.world, self.group = ev, rank, world, group

    def _all_gather_cat(self, t: torch.Tensor, sizes) -> torch.Tensor:
        outs = [torch.empty((s,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for s in sizes]
        dist.all_gather(outs, t.contiguous(), group=self.group)
        return torch.cat(outs, 0)

    def exchange_plan(self, gt_col, n_queries: int, n_tracks: int):
        """Host-side plan of the paired-track exchange (no device sync): detection of query q runs on
        the rank that owns q, its paired track gt_col[q] lives on the rank that owns that gallery
        column.  Returns (send_loc [sum], in_splits [W], out_splits [W], perm [q1-q0]) where
        send_loc are local track indices grouped by destination rank (query order inside a group),
        and perm[i] is the position of my i-th query's row in the received buffer."""
        import numpy as np
        W, R = self.world, self.rank
        gt = np.asarray(gt_col.cpu() if isinstance(gt_col, torch.Tensor) else gt_col, dtype=np.int64)
        q_b = [shard_bounds(n_queries, r, W) for r in range(W)]
        m_b = [shard_bounds(n_tracks, r, W) for r in range(W)]
        m0, m1 = m_b[R]
        send_loc, in_splits = [], []
        for d in range(W):
            g = gt[q_b[d][0]:q_b[d][1]]
            mine = g[(g >= m0) & (g < m1)] - m0
            send_loc.append(mine)
            in_splits.append(int(mine.shape[0]))
        g = gt[q_b[R][0]:q_b[R][1]]
        starts = np.array([b[0] for b in m_b] + [n_tracks])
        owner = np.searchsorted(starts, g, side="right") - 1
        out_splits = [int((owner == s_).sum()) for s_ in range(W)]
        # received rows are grouped by source rank, query order inside a group
        order = np.argsort(owner, kind="stable")          # received position j holds query order[j]
        perm = np.empty_like(order)
        perm[order] = np.arange(order.shape[0])
        return np.concatenate(send_loc) if send_loc else np.zeros(0, np.int64), in_splits, out_splits, perm

    def _all_to_all_rows(self, send: torch.Tensor, in_splits, out_splits) -> torch.Tensor:
        recv = torch.empty((sum(out_splits),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send.contiguous(), out_splits, in_splits, group=self.group)
        return recv

    @torch.no_grad()
    def run(self, videos: Dict[str, torch.Tensor], tracks: Dict[str, torch.Tensor], gt_col: torch.Tensor,
            n_queries: int, n_tracks: int, on_host: bool = False):
        """`videos` holds THIS rank's slice of the queries, `tracks` THIS rank's gallery shard
        (features, masks, gt_moment, m_duration); gt_col [n_queries] are GLOBAL column indices
        (pass a CPU tensor: the exchange plan is made on the host without a device sync)."""
        ev, dev, W, R = self.ev, self.ev.dev, self.world, self.rank
        ev.launches = 0
        ev._keep = []
        trace = os.environ.get("MADE_TRACE_PHASES")     # diagnostics: per-phase wall times (adds syncs)
        marks = []

        def mark(name):
            if trace:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        mark("start")
        q0, q1 = shard_bounds(n_queries, R, W)
        m0, m1 = shard_bounds(n_tracks, R, W)
        q_sizes = [shard_bounds(n_queries, r, W)[1] - shard_bounds(n_queries, r, W)[0] for r in range(W)]
        send_loc, in_splits, out_splits, perm = self.exchange_plan(gt_col, n_queries, n_tracks)
        send_loc_d = torch.from_numpy(send_loc).to(dev, non_blocking=True)
        perm_d = torch.from_numpy(perm).to(dev, non_blocking=True)
        frame_seq, vf_local, frame_mask = ev.encode_queries(videos["frame_feats"], videos["frame_mask"])
        mark("plan+encode_queries")
        gal = ev.encode_gallery(tracks["segment_feats"], tracks["segment_mask"])
        mark("encode_gallery")
        # ---- exchange 5 (issued early): every query's paired track -> the rank that detects it ----
        gtm = tracks["gt_moment"].to(dev, non_blocking=True).reshape(-1, 2).to(torch.float32)
        mdur = tracks["m_duration"].to(dev, non_blocking=True).to(torch.float32)
        aux = torch.cat([gal["mask"], gtm, mdur.unsqueeze(1)], 1)                  # [n_local, 96 + 3]
        recv_seq = self._all_to_all_rows(gal["seq"][send_loc_d], in_splits, out_splits)[perm_d]
        recv_aux = self._all_to_all_rows(aux[send_loc_d], in_splits, out_splits)[perm_d]
        mark("all_to_all")
        # ---- detection for this rank's queries on the received tracks: enqueued now on the detection
        # stream, it overlaps the scoring / ranking / collectives below ----
        pair = dict(seq=recv_seq, mask=recv_aux[:, :cfg.L_M].contiguous())
        det = ev.detect(frame_seq, frame_mask, pair, vf_local,
                        torch.arange(q1 - q0, dtype=torch.int32, device=dev),
                        recv_aux[:, cfg.L_M:cfg.L_M + 2].contiguous(), recv_aux[:, cfg.L_M + 2].contiguous())
        video_feats = self._all_gather_cat(vf_local, q_sizes)                      # exchange 1
        mark("all_gather_q")
        single, dual = ev.score(video_feats, gal)
        mark("score")
        gt = (gt_col if gt_col.device == dev else gt_col.to(dev, non_blocking=True)).to(torch.int32)
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
        mark("rank+topk+collectives")
        if hasattr(ev, "join_detect"):
            ev.join_detect()
        mark("detect")
        if trace and R == 0:
            print("[phases rank 0] " + " ".join(f"{n}={1e3 * (t - marks[i][1]):.2f}" for i, (n, t) in
                                                 enumerate(marks[1:])), flush=True)
        return dict(rank=rank_cnt, topk_idx=topk_idx, topk_score=topk_score, q_range=(q0, q1), **det)
