"""Whole-job evaluation: match + moment-detect for N_v query videos against an N_m-track gallery.

This is the B200 replacement of the body of test-MaDe.py:eval_epoch (243-447): encode queries and
gallery, fused full-gallery X-Pool scoring (never materialising [N_m,N_v,256]), dual-tower cosine,
fp64-sum ranking/top-k, DETR moment detection for the paired track, IoU — all as C-ABI kernel
launches on one stream; host<->device copies of the e2e path run on a second stream and overlap
the kernels chunk by chunk.  Multi-GPU: the gallery is sharded by track, queries are replicated
for scoring and sharded for detection (`shard=(rank, world)`), see `parallel.py`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from . import config as cfg
from .engine import Engine


class GalleryEvaluator:
    def __init__(self, engine: Engine, k: int = 100, music_chunk: int = 1024, video_chunk: int = 1024,
                 detr_chunk: int = 500):
        self.eng = engine
        self.dev = engine.device
        self.k = k
        self.music_chunk = music_chunk
        self.video_chunk = video_chunk
        self.detr_chunk = detr_chunk
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.launches = 0   # kernels launched by the last run (counted per C-ABI op, see _count)
        self.xpool_events = None   # bench.py: list that receives (start, end) CUDA events per xpool launch

    # ---- launch accounting (bench.py reports gpu_launches) --------------------------------------
    # kernels per C-ABI call: encode = cast + 6 GEMM + attn + pool = 9; gallery_prepare = LN + GEMM +
    # Gram GEMM + maskbits = 4; query_prepare = LN + GEMM + vhat = 3; xpool_score = 1; cosine = 1;
    # rank_topk = 1; detr_detect = prep 1 + enc 2*6 + KV 2 + cast 1 + dec 6*6 (5 GEMM + CA) + 6 D2D copies
    # (not kernels) + LN 1 + heads 3 = 56; moment_postproc = 1.
    _K = dict(encode=9, gallery_prepare=4, query_prepare=3, xpool=1, cosine=1, rank=1, detr=56, postproc=1)

    def _count(self, what: str, n: int = 1):
        self.launches += self._K[what] * n

    # ---- stages -----------------------------------------------------------------------------------
    def encode_queries(self, frame_feats, frame_mask, on_host: bool):
        n = frame_feats.shape[0]
        seq = torch.empty((n, cfg.L_V, cfg.D_MODEL), dtype=torch.float16, device=self.dev)
        pooled = torch.empty((n, cfg.D_MODEL), dtype=torch.float32, device=self.dev)
        mask_d = self._to_dev(frame_mask, on_host)
        for s, e, feats_d in self._chunks(frame_feats, self.video_chunk, on_host):
            sq, _, pl = self.eng.encode(_lib.VIDEO, feats_d, mask_d[s:e], want_f32=False)
            seq[s:e] = sq
            pooled[s:e] = pl
            self._count("encode")
        return seq, pooled, mask_d

    def encode_gallery(self, segment_feats, segment_mask, on_host: bool):
        n = segment_feats.shape[0]
        seq = torch.empty((n, cfg.L_M, cfg.D_MODEL), dtype=torch.float16, device=self.dev)
        pooled = torch.empty((n, cfg.D_MODEL), dtype=torch.float32, device=self.dev)
        kz = torch.empty((n * cfg.L_M, 3 * cfg.D_MODEL), dtype=torch.float16, device=self.dev)
        gram = torch.empty((n * cfg.L_M, cfg.L_M), dtype=torch.float16, device=self.dev)
        bits = torch.empty((n, 4), dtype=torch.int32, device=self.dev)
        mask_d = self._to_dev(segment_mask, on_host)
        for s, e, feats_d in self._chunks(segment_feats, self.music_chunk, on_host):
            sq, _, pl = self.eng.encode(_lib.MUSIC, feats_d, mask_d[s:e], want_f32=False)
            seq[s:e] = sq
            pooled[s:e] = pl
            k_, g_, b_ = self.eng.gallery_prepare(sq, mask_d[s:e])
            kz[s * cfg.L_M:e * cfg.L_M] = k_
            gram[s * cfg.L_M:e * cfg.L_M] = g_
            bits[s:e] = b_
            self._count("encode")
            self._count("gallery_prepare")
        return dict(seq=seq, pooled=pooled, kz=kz, gram=gram, bits=bits, mask=mask_d)

    def score(self, video_feats, gal, out=None, col_offset: int = 0):
        """single/dual similarity of every query against this gallery (shard)."""
        n_q, n_m = video_feats.shape[0], gal["bits"].shape[0]
        if out is None:
            single = torch.empty((n_q, n_m), dtype=torch.float32, device=self.dev)
            dual = torch.empty((n_q, n_m), dtype=torch.float32, device=self.dev)
        else:
            single, dual = out
        q, vhat = self.eng.query_prepare(video_feats)
        if self.xpool_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.eng.xpool_score(q, vhat, gal["kz"], gal["gram"], gal["bits"], out=single, col_offset=col_offset)
        if self.xpool_events is not None:
            e1.record()
            self.xpool_events.append((e0, e1))
        ops.cal_distance(video_feats, gal["pooled"], out=dual, col_offset=col_offset)
        self._count("query_prepare"), self._count("xpool"), self._count("cosine")
        return single, dual

    def detect(self, frame_seq, frame_mask, gal, video_feats, track_idx, gt_moment, m_duration):
        """DETR moment detection for (query b, track track_idx[b]) pairs + post-processing + IoU."""
        n = video_feats.shape[0]
        st = torch.empty(n, dtype=torch.float32, device=self.dev)
        ed, sc, iou = torch.empty_like(st), torch.empty_like(st), torch.empty_like(st)
        spans = torch.empty((n, 2), dtype=torch.float32, device=self.dev)
        for s in range(0, n, self.detr_chunk):
            e = min(n, s + self.detr_chunk)
            r = self.eng.detr_detect(frame_seq[s:e], frame_mask[s:e], gal["seq"], gal["mask"], video_feats[s:e],
                                     track_idx=track_idx[s:e])
            spans[s:e] = r["pred_spans"][-1]
            a, b, c, d = ops.moment_postproc(r["pred_logits"][-1], r["pred_spans"][-1], gt_moment[s:e], m_duration[s:e])
            st[s:e], ed[s:e], sc[s:e], iou[s:e] = a, b, c, d
            self._count("detr"), self._count("postproc")
        return dict(pred_st=st, pred_ed=ed, score=sc, iou=iou, pred_spans=spans)

    # ---- whole job --------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, videos: Dict[str, torch.Tensor], tracks: Dict[str, torch.Tensor], gt_col: torch.Tensor,
            prev_same: Optional[torch.Tensor] = None, on_host: bool = False, want_sims: bool = False):
        """One step of the hot path.  `videos`/`tracks` are the dicts of `synth.make_*` (device
        resident, or pinned host tensors with on_host=True).  Query i is paired with track gt_col[i]
        for both the rank and the moment detection (test-MaDe.py:280 evaluates the paired track)."""
        self.launches = 0
        frame_seq, video_feats, frame_mask = self.encode_queries(videos["frame_feats"], videos["frame_mask"], on_host)
        gal = self.encode_gallery(tracks["segment_feats"], tracks["segment_mask"], on_host)
        single, dual = self.score(video_feats, gal)
        gt_col_d = self._to_dev(gt_col, on_host).to(torch.int32)
        prev_d = None if prev_same is None else self._to_dev(prev_same, on_host).to(torch.int32)
        rk = ops.rank_topk(single, dual, gt_col_d, prev_d, k=self.k)
        self._count("rank")
        gt_moment = self._to_dev(tracks["gt_moment"], on_host)
        m_dur = self._to_dev(tracks["m_duration"], on_host)
        idx64 = gt_col_d.long()
        det = self.detect(frame_seq, frame_mask, gal, video_feats, gt_col_d, gt_moment[idx64], m_dur[idx64])
        out = dict(rank=rk["rank"], topk_idx=rk["topk_idx"], topk_score=rk["topk_score"], gt_score=rk["gt_score"],
                   video_feats=video_feats, music_feats=gal["pooled"], **det)
        if want_sims:
            out.update(single=single, dual=dual)
        return out

    def to_host(self, out: Dict[str, torch.Tensor], keys=("rank", "topk_idx", "iou", "pred_st", "pred_ed", "score")):
        """Device→host read of the step's result (what eval_epoch consumes on the CPU)."""
        host = {k: out[k].cpu() for k in keys}
        return host

    # ---- helpers ----------------------------------------------------------------------------------
    def _to_dev(self, t: torch.Tensor, on_host: bool):
        if t.device == self.dev:
            return t
        return t.to(self.dev, non_blocking=True)

    def _chunks(self, feats: torch.Tensor, chunk: int, on_host: bool):
        """Yield (start, end, device chunk).  Host inputs are copied on the copy stream one chunk
        ahead of the kernels that consume them."""
        n = feats.shape[0]
        bounds = [(s, min(n, s + chunk)) for s in range(0, n, chunk)]
        if not on_host:
            for s, e in bounds:
                yield s, e, feats[s:e]
            return
        cur = torch.cuda.current_stream(self.dev)
        pending = []

        def issue(i):
            s, e = bounds[i]
            with torch.cuda.stream(self.copy_stream):
                d = feats[s:e].to(self.dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            pending.append((d, ev))

        issue(0)
        for i, (s, e) in enumerate(bounds):
            if i + 1 < len(bounds):
                issue(i + 1)
            d, ev = pending.pop(0)
            cur.wait_event(ev)
            d.record_stream(cur)
            yield s, e, d
