"""Gallery sharding across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Tracks are independent units, so the
gallery is split into contiguous shards (never through a music id: `plan_track_shards`); queries are
sharded for encoding, detection and for OWNING their results.  Four exchange steps per job:
  1. all_gather of the query embeddings [N_v/G, 256] (each rank encodes its slice of the queries),
  2. all_to_all of the paired tracks: detection of query q runs on the rank that owns q, its paired
     track lives on the rank that owns that gallery column; encoded segments, mask, ground-truth
     moment and duration of a pair travel as ONE packed byte row (planned on the host from the
     pairing: no device sync, exact bytes),
  3. all_reduce(MAX) of the ground-truth scores [N_v] f64 (the GT id lives on one shard; every
     shard needs its score to count the ids that beat it),
  4. all_to_all of ONE packed int32 row per query — [count of ids ahead | k candidate columns |
     k candidate scores (f64 as two words)] — to the rank that owns the query, which sums the
     counts and merges the candidates (`made_topk_merge`).
Dedup (`Recall_metrics(dedup=True)`, utils/util_test.py:46-60): all columns that carry one music id
sit on one shard, so "distinct ids ahead of the ground truth" is a sum of per-shard distinct counts
and the ground-truth score is the MAX over the shards of the per-shard best column of that id.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
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


def plan_track_shards(music_ids: Optional[Sequence[str]], n_tracks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous gallery shards [(m0, m1)] * world whose boundaries never separate two columns that carry
    the same music id (the dedup rank is then a plain sum over shards).  Starts from the uniform split and
    moves each boundary forward to the first column whose id has no earlier occurrence at or after the
    previous boundary.  Raises ValueError when the ids interleave so that no contiguous split exists
    (sort the gallery by music id first: `order_tracks_by_id`)."""
    bounds = [shard_bounds(n_tracks, r, world) for r in range(world)]
    if music_ids is None:
        return bounds
    if len(music_ids) != n_tracks:
        raise ValueError(f"music_ids has {len(music_ids)} entries for {n_tracks} tracks")
    first, last = {}, {}
    for c, mid in enumerate(music_ids):
        first.setdefault(mid, c)
        last[mid] = c
    cuts = [0]
    for r in range(1, world):
        c = max(bounds[r][0], cuts[-1])
        # a cut at c is legal iff no id has occurrences on both sides: advance past every id that straddles it
        moved = True
        while moved and c < n_tracks:
            moved = False
            for j in range(cuts[-1], c):
                if last[music_ids[j]] >= c:
                    c = last[music_ids[j]] + 1
                    moved = True
        cuts.append(min(c, n_tracks))
    cuts.append(n_tracks)
    out = [(cuts[r], cuts[r + 1]) for r in range(world)]
    for (a, b) in out:
        for j in range(a, b):
            if first[music_ids[j]] < a or last[music_ids[j]] >= b:
                raise ValueError("music ids interleave across shards; order the gallery by id (order_tracks_by_id)")
    return out


def order_tracks_by_id(music_ids: Sequence[str]) -> np.ndarray:
    """Stable permutation that makes the columns of every music id adjacent (first-occurrence order)."""
    slot = {}
    for mid in music_ids:
        slot.setdefault(mid, len(slot))
    return np.argsort(np.array([slot[m] for m in music_ids]), kind="stable")


class ShardedEvaluator:
    """One (N_v queries x N_m tracks) job on `world` GPUs: gallery sharded by track, queries sharded for
    encoding / detection / result ownership."""

    ROW_SEQ = cfg.L_M * cfg.D_MODEL * 2                    # fp16 encoded segments of one track, bytes
    ROW_AUX = (cfg.L_M + 3) * 4                            # mask [96] + gt moment [2] + duration [1], fp32

    def __init__(self, ev: GalleryEvaluator, rank: int, world: int, group=None,
                 track_bounds: Optional[List[Tuple[int, int]]] = None):
        self.ev, self.rank, self.world, self.group = ev, rank, world, group
        self.track_bounds = track_bounds

    # ---- host-side planning ------------------------------------------------------------------------
    def _bounds(self, n_queries: int, n_tracks: int):
        W = self.world
        q_b = [shard_bounds(n_queries, r, W) for r in range(W)]
        m_b = self.track_bounds if self.track_bounds is not None else [shard_bounds(n_tracks, r, W) for r in range(W)]
        if len(m_b) != W or m_b[0][0] != 0 or m_b[-1][1] != n_tracks or any(m_b[r][1] != m_b[r + 1][0] for r in range(W - 1)):
            raise ValueError("track_bounds must be a contiguous cover of the gallery, one interval per rank")
        return q_b, m_b

    def exchange_plan(self, gt_col, n_queries: int, n_tracks: int):
        """Host-side plan of the paired-track exchange (no device sync).  Returns (send_loc [sum],
        in_splits [W], out_splits [W], perm [q1-q0]) where send_loc are local track indices grouped by
        destination rank (query order inside a group), and perm[i] is the position of my i-th query's
        row in the received buffer."""
        W, R = self.world, self.rank
        gt = np.asarray(gt_col.cpu() if isinstance(gt_col, torch.Tensor) else gt_col, dtype=np.int64)
        q_b, m_b = self._bounds(n_queries, n_tracks)
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
        # empty shards share a start with their successor: searchsorted then names the LAST of them — the
        # one that is not empty is the one whose interval contains g
        for i, gg in enumerate(g):
            while not (m_b[owner[i]][0] <= gg < m_b[owner[i]][1]):
                owner[i] -= 1
        out_splits = [int((owner == s_).sum()) for s_ in range(W)]
        # received rows are grouped by source rank, query order inside a group
        order = np.argsort(owner, kind="stable")          # received position j holds query order[j]
        perm = np.empty_like(order)
        perm[order] = np.arange(order.shape[0])
        return np.concatenate(send_loc) if send_loc else np.zeros(0, np.int64), in_splits, out_splits, perm

    def dedup_plan(self, music_ids: Optional[Sequence[str]], gt_col, n_tracks: int):
        """→ (gt_col' [N_v] int64: the LAST column that carries each query's ground-truth id,
        prev_local [m1-m0] int32 or None: previous LOCAL column with the same id, -1 if none)."""
        gt = np.asarray(gt_col.cpu() if isinstance(gt_col, torch.Tensor) else gt_col, dtype=np.int64)
        if music_ids is None:
            return gt, None
        _, m_b = self._bounds(1, n_tracks)
        prev, _, has_dups = ops.dedup_tables(music_ids)
        if not has_dups:
            return gt, None
        last = {}
        for c, mid in enumerate(music_ids):
            last[mid] = c
        gt2 = np.array([last[music_ids[g]] for g in gt], dtype=np.int64)
        m0, m1 = m_b[self.rank]
        loc = prev[m0:m1].astype(np.int64)
        if ((loc >= 0) & (loc < m0)).any():
            raise ValueError("a music id spans two gallery shards: build track_bounds with plan_track_shards()")
        loc = np.where(loc >= 0, loc - m0, -1).astype(np.int32)
        return gt2, loc

    # ---- collectives ---------------------------------------------------------------------------------
    def _all_gather_cat(self, t: torch.Tensor, sizes) -> torch.Tensor:
        outs = [torch.empty((s,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for s in sizes]
        dist.all_gather(outs, t.contiguous(), group=self.group)
        return torch.cat(outs, 0)

    def _all_to_all_rows(self, send: torch.Tensor, in_splits, out_splits) -> torch.Tensor:
        recv = torch.empty((sum(out_splits),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send.contiguous(), out_splits, in_splits, group=self.group)
        return recv

    @torch.no_grad()
    def run(self, videos: Dict[str, torch.Tensor], tracks: Dict[str, torch.Tensor], gt_col: torch.Tensor,
            n_queries: int, n_tracks: int, on_host: bool = False, music_ids: Optional[Sequence[str]] = None,
            gather_results: bool = True, want_sims: bool = False):
        """`videos` holds THIS rank's slice of the queries, `tracks` THIS rank's gallery shard
        (features, masks, gt_moment, m_duration); gt_col [n_queries] are GLOBAL column indices
        (pass a CPU tensor: the exchange plan is made on the host without a device sync);
        music_ids (optional, GLOBAL list in column order) turns on the dedup rank.
        → rank / topk_idx / topk_score for every query (gather_results=True: one more all_gather) or
        for this rank's queries only, plus the detection outputs of this rank's queries."""
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
        q_b, m_b = self._bounds(n_queries, n_tracks)
        q0, q1 = q_b[R]
        m0, m1 = m_b[R]
        q_sizes = [b - a for a, b in q_b]
        gt_last, prev_local = self.dedup_plan(music_ids, gt_col, n_tracks)
        send_loc, in_splits, out_splits, perm = self.exchange_plan(gt_col, n_queries, n_tracks)
        send_loc_d = torch.from_numpy(send_loc).to(dev, non_blocking=True)
        perm_d = torch.from_numpy(perm).to(dev, non_blocking=True)
        prev_d = None if prev_local is None else torch.from_numpy(prev_local).to(dev, non_blocking=True)
        gt = torch.from_numpy(gt_last).to(dev, non_blocking=True).to(torch.int32)
        # host inputs: the head of the gallery shard goes to the copy engine before the queries' many small copies
        # (GalleryEvaluator.run, same switch)
        seg = tracks["segment_feats"]
        prime = 3 if (on_host and dev.type == "cuda" and not seg.is_cuda and ev.h2d_mode in ("dma", "dma16") and
                      os.environ.get("MADE_PRIME_GALLERY", "1") != "0") else 0
        started = ev.start_gallery(seg, tracks["segment_mask"], prime) if prime else None
        frame_seq, vf_local, frame_mask = ev.encode_queries(videos["frame_feats"], videos["frame_mask"])
        video_feats = self._all_gather_cat(vf_local, q_sizes)                      # exchange 1
        mark("plan+encode_queries+all_gather_q")
        gal = ev.encode_gallery(seg, tracks["segment_mask"], started=started) if started is not None else \
            ev.encode_gallery(seg, tracks["segment_mask"])
        mark("encode_gallery")
        # ---- exchange 2: every query's paired track -> the rank that detects it, one packed byte row per pair ----
        gtm = tracks["gt_moment"].to(dev, non_blocking=True).reshape(-1, 2).to(torch.float32)
        mdur = tracks["m_duration"].to(dev, non_blocking=True).to(torch.float32)
        aux = torch.cat([gal["mask"], gtm, mdur.unsqueeze(1)], 1)                  # [n_local, 96 + 3] fp32
        seq_rows = gal["seq"].reshape(gal["seq"].shape[0], -1)                     # [n_local, 96*256] fp16
        if dev.type == "cuda":
            row = torch.cat([seq_rows[send_loc_d].view(torch.uint8), aux[send_loc_d].view(torch.uint8)], 1)
            recv = self._all_to_all_rows(row, in_splits, out_splits)[perm_d]
            recv_seq = recv[:, :self.ROW_SEQ].contiguous().view(torch.float16).reshape(-1, cfg.L_M, cfg.D_MODEL)
            recv_aux = recv[:, self.ROW_SEQ:].contiguous().view(torch.float32)
        else:   # the CPU stand-in of the gloo tests keeps fp32 segments: two exchanges
            recv_seq = self._all_to_all_rows(gal["seq"][send_loc_d], in_splits, out_splits)[perm_d]
            recv_aux = self._all_to_all_rows(aux[send_loc_d], in_splits, out_splits)[perm_d]
        mark("all_to_all")
        # ---- detection for this rank's queries on the received tracks: enqueued now on the detection
        # stream, it overlaps the scoring / ranking / collectives below ----
        pair = dict(seq=recv_seq, mask=recv_aux[:, :cfg.L_M].contiguous())
        det = ev.detect(frame_seq, frame_mask, pair, vf_local,
                        torch.arange(q1 - q0, dtype=torch.int32, device=dev),
                        recv_aux[:, cfg.L_M:cfg.L_M + 2].contiguous(), recv_aux[:, cfg.L_M + 2].contiguous())
        single, dual = ev.score(video_feats, gal)
        mark("score")
        k = ev.k
        local_gt = torch.where((gt >= m0) & (gt < m1), gt - m0, torch.full_like(gt, -1))
        r1 = ops.rank_topk(single, dual, local_gt, prev_d, k=0)
        ev._count("rank")
        gt_score = r1["gt_score"]
        dist.all_reduce(gt_score, op=dist.ReduceOp.MAX, group=self.group)          # exchange 3
        r2 = ops.rank_topk(single, dual, None, prev_d, k=k, col_offset=m0, gt_score_in=gt_score)
        ev._count("rank")
        # ---- exchange 4: [count | k columns | k scores] of query q -> the rank that owns q ----
        packed = torch.cat([r2["rank"].reshape(-1, 1).to(torch.int32), r2["topk_idx"],
                            r2["topk_score"].view(torch.int32)], 1)                 # [N_v, 1 + 3k] int32
        nl = q1 - q0
        recv = torch.empty((W, nl, 1 + 3 * k), dtype=torch.int32, device=dev)
        dist.all_to_all_single(recv.view(W * nl, -1), packed, [nl] * W, q_sizes, group=self.group)
        rank_cnt = recv[:, :, 0].sum(0, dtype=torch.int32)
        cand_i = recv[:, :, 1:1 + k].permute(1, 0, 2).reshape(nl, W * k)
        cand_s = recv[:, :, 1 + k:].permute(1, 0, 2).reshape(nl, W * 2 * k).contiguous().view(torch.float64)
        topk_idx, topk_score = ops.topk_merge(cand_s, cand_i, k)
        ev.launches += 1
        mark("rank+topk+collectives")
        if gather_results and W > 1:
            res = torch.cat([rank_cnt.reshape(-1, 1), topk_idx, topk_score.view(torch.int32)], 1)
            res = self._all_gather_cat(res, q_sizes)
            rank_cnt, topk_idx = res[:, 0].contiguous(), res[:, 1:1 + k].contiguous()
            topk_score = res[:, 1 + k:].contiguous().view(torch.float64)
        if hasattr(ev, "join_detect"):
            ev.join_detect()
        mark("detect")
        if trace and R == 0:
            print("[phases rank 0] " + " ".join(f"{n}={1e3 * (t - marks[i][1]):.2f}" for i, (n, t) in
                                                 enumerate(marks[1:])), flush=True)
        out = dict(rank=rank_cnt, topk_idx=topk_idx, topk_score=topk_score, q_range=(q0, q1),
                   gathered=bool(gather_results), **det)
        if want_sims:
            out.update(single=single, dual=dual, gt_score=gt_score)
        return out
