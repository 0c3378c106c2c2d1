"""Torch-tensor front ends of the stateless C-ABI kernels (same names and argument meaning as the
reference's free functions).  Tensors must live on the CUDA device; outputs are fresh tensors
allocated by torch and filled by the kernels on the current stream."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MAX_M_DURATION = 240.0


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


def _default_device() -> torch.device:
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(t: torch.Tensor) -> torch.Tensor:
    """Host tensors handed to a reference-named free function: moved to the current CUDA device (there is
    no CPU path; without a device this raises)."""
    if t.is_cuda:
        return t
    _lib.require_cuda()
    return t.to(torch.device("cuda", torch.cuda.current_device()))


# ---------------------------------------------------------------------------------------------
# music_detr/span_utils.py
# ---------------------------------------------------------------------------------------------
def span_cw_to_se(cw_spans: torch.Tensor) -> torch.Tensor:
    """span_utils.py:15-24."""
    cw = _f32c(cw_spans)
    out = torch.empty_like(cw)
    _lib.check(_lib.load().made_span_cw_to_se(_lib.ptr(cw), _lib.ptr(out), cw.shape[0], _lib.stream_ptr()))
    return out


def span_se_to_cw(se_spans: torch.Tensor) -> torch.Tensor:
    """span_utils.py:4-13."""
    se = _f32c(se_spans)
    out = torch.empty_like(se)
    _lib.check(_lib.load().made_span_se_to_cw(_lib.ptr(se), _lib.ptr(out), se.shape[0], _lib.stream_ptr()))
    return out


def span_iou(pred_st: torch.Tensor, pred_ed: torch.Tensor, gt_moment: torch.Tensor, m_duration: torch.Tensor,
             max_m_duration: float = MAX_M_DURATION) -> torch.Tensor:
    """detr_iou + individual_IoU_tensor (span_utils.py:119-170), batched: spans in seconds → IoU [n]."""
    st, ed = _f32c(pred_st.reshape(-1)), _f32c(pred_ed.reshape(-1))
    gt, md = _f32c(gt_moment.reshape(-1, 2)), _f32c(m_duration.reshape(-1))
    if not (st.shape[0] == ed.shape[0] == gt.shape[0] == md.shape[0]):
        raise ValueError("span_iou: inputs disagree on the number of spans")
    out = torch.empty_like(st)
    _lib.check(_lib.load().made_span_iou(_lib.ptr(st), _lib.ptr(ed), _lib.ptr(gt), _lib.ptr(md), max_m_duration,
                                         st.shape[0], _lib.ptr(out), _lib.stream_ptr()))
    return out


def detr_iou(args, mr_results_list, device=None) -> List[torch.Tensor]:
    """span_utils.py:147-170 with the reference's argument (list of dicts with "gt_moment" [1,2],
    "m_duration", "ranked_preds" [#preds,3]); one kernel launch for the whole list.  Returns the
    reference's list of 0-d tensors (on the host)."""
    if not mr_results_list:
        return []
    dev = torch.device(device) if device is not None else _default_device()
    st = torch.tensor([float(d["ranked_preds"][0][0]) for d in mr_results_list], dtype=torch.float32)
    ed = torch.tensor([float(d["ranked_preds"][0][1]) for d in mr_results_list], dtype=torch.float32)
    gt = torch.stack([torch.as_tensor(d["gt_moment"], dtype=torch.float32).reshape(-1, 2)[0] for d in mr_results_list])
    md = torch.tensor([float(d["m_duration"]) for d in mr_results_list], dtype=torch.float32)
    iou = span_iou(st.to(dev), ed.to(dev), gt.to(dev), md.to(dev), float(getattr(args, "max_m_duration", MAX_M_DURATION)))
    return list(iou.cpu().unbind(0))


def _check_spans(s: torch.Tensor, name: str) -> torch.Tensor:
    if s.dim() != 2 or s.shape[1] != 2:
        raise ValueError(f"{name} must be [n,2], got {tuple(s.shape)}")
    return _f32c(s)


def generalized_temporal_iou(spans1: torch.Tensor, spans2: torch.Tensor, check: bool = True) -> torch.Tensor:
    """span_utils.py:86-115.  `check=True` keeps the reference's `assert e >= s` (which costs a
    device sync, as it does in the reference)."""
    s1, s2 = _check_spans(spans1, "spans1"), _check_spans(spans2, "spans2")
    if check:
        assert (s1[:, 1] >= s1[:, 0]).all()
        assert (s2[:, 1] >= s2[:, 0]).all()
    out = torch.empty((s1.shape[0], s2.shape[0]), dtype=torch.float32, device=s1.device)
    _lib.check(_lib.load().made_giou(_lib.ptr(s1), s1.shape[0], _lib.ptr(s2), s2.shape[0], _lib.ptr(out),
                                     _lib.stream_ptr()))
    return out


def temporal_iou(spans1: torch.Tensor, spans2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """span_utils.py:39-66 → (iou, union)."""
    s1, s2 = _check_spans(spans1, "spans1"), _check_spans(spans2, "spans2")
    iou = torch.empty((s1.shape[0], s2.shape[0]), dtype=torch.float32, device=s1.device)
    uni = torch.empty_like(iou)
    _lib.check(_lib.load().made_temporal_iou(_lib.ptr(s1), s1.shape[0], _lib.ptr(s2), s2.shape[0], _lib.ptr(iou),
                                             _lib.ptr(uni), _lib.stream_ptr()))
    return iou, uni


def matcher_cost(prob_fg: torch.Tensor, out_spans_cw: torch.Tensor, tgt_spans_cw: torch.Tensor,
                 cost_span: float = 10.0, cost_giou: float = 1.0, cost_class: float = 4.0) -> torch.Tensor:
    """Cost matrix of HungarianMatcher.forward (matcher.py:66-88) with build_matcher's weights."""
    p = _f32c(prob_fg)
    a, b = _check_spans(out_spans_cw, "out_spans"), _check_spans(tgt_spans_cw, "tgt_spans")
    if p.shape[0] != a.shape[0]:
        raise ValueError("prob_fg and out_spans disagree on the number of predictions")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().made_matcher_cost(_lib.ptr(p), _lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0],
                                             cost_span, cost_giou, cost_class, _lib.ptr(out), _lib.stream_ptr()))
    return out


def moment_postproc(pred_logits: torch.Tensor, pred_spans: torch.Tensor, gt_moment: Optional[torch.Tensor] = None,
                    m_duration: Optional[torch.Tensor] = None, max_m_duration: float = MAX_M_DURATION):
    """test-MaDe.py:306-316 + detr_iou (span_utils.py:147-170): logits [n,1,2] or [n,2], spans
    (c,w) likewise, gt_moment [n,1,2] seconds → (pred_st, pred_ed, score, iou or None)."""
    lg = _f32c(pred_logits.reshape(-1, 2))
    sp = _f32c(pred_spans.reshape(-1, 2))
    n = lg.shape[0]
    st = torch.empty(n, dtype=torch.float32, device=lg.device)
    ed, sc = torch.empty_like(st), torch.empty_like(st)
    iou = gt = md = None
    if gt_moment is not None:
        gt = _f32c(gt_moment.reshape(-1, 2))
        md = _f32c(m_duration.reshape(-1))
        iou = torch.empty_like(st)
    _lib.check(_lib.load().made_moment_postproc(_lib.ptr(lg), _lib.ptr(sp), _lib.ptr(gt), _lib.ptr(md),
                                                max_m_duration, n, _lib.ptr(st), _lib.ptr(ed), _lib.ptr(sc),
                                                _lib.ptr(iou), _lib.stream_ptr()))
    return st, ed, sc, iou


# ---------------------------------------------------------------------------------------------
# ranking (utils/util_test.py:32-97) and similarities (modules/loss.py:52-56)
# ---------------------------------------------------------------------------------------------
def dedup_tables(music_ids: Sequence[str], gt_ids: Optional[Sequence[str]] = None):
    """Host-side id bookkeeping for Recall_metrics(dedup=True): prev_same[c] = previous column with
    the same music id (-1 if none); gt_col[r] = LAST column carrying row r's ground-truth id
    (row r's GT id defaults to music_ids[r], the reference's square layout, Q13)."""
    last = {}
    prev = np.full(len(music_ids), -1, dtype=np.int32)
    for c, mid in enumerate(music_ids):
        if mid in last:
            prev[c] = last[mid]
        last[mid] = c
    if gt_ids is None:
        gt_ids = music_ids
    gt_col = np.array([last.get(g, -1) for g in gt_ids], dtype=np.int32)
    has_dups = bool((prev >= 0).any())
    return prev, gt_col, has_dups


def rank_topk(single: torch.Tensor, dual: Optional[torch.Tensor], gt_col: Optional[torch.Tensor] = None,
              prev_same: Optional[torch.Tensor] = None, k: int = 0, col_offset: int = 0,
              gt_score_in: Optional[torch.Tensor] = None, n_cols: Optional[int] = None):
    """Exact top-k and dedup-aware ground-truth rank of rows of (double(single) + double(dual)).
    Returns dict(topk_idx [n,k] int32, topk_score [n,k] f64, rank [n] int32, gt_score [n] f64)."""
    if single.dtype != torch.float32 or (dual is not None and dual.dtype != torch.float32):
        raise ValueError("rank_topk takes fp32 similarity matrices")
    n_rows, ld = single.shape[0], single.stride(0)
    n_cols = single.shape[1] if n_cols is None else n_cols
    if dual is not None and (dual.stride(0) != ld or dual.shape[0] != n_rows):
        raise ValueError("single and dual must share their layout")
    dev = single.device
    out = {}
    topk_idx = topk_score = rank = gt_score = None
    if k > 0:
        topk_idx = torch.empty((n_rows, k), dtype=torch.int32, device=dev)
        topk_score = torch.empty((n_rows, k), dtype=torch.float64, device=dev)
    if gt_col is not None or gt_score_in is not None:
        rank = torch.empty(n_rows, dtype=torch.int32, device=dev)
        gt_score = torch.empty(n_rows, dtype=torch.float64, device=dev)
    _lib.check(_lib.load().made_rank_topk(
        single.data_ptr(), None if dual is None else dual.data_ptr(), ld, n_rows, n_cols, _lib.ptr(gt_col),
        _lib.ptr(gt_score_in), _lib.ptr(prev_same), col_offset, k, _lib.ptr(topk_idx), _lib.ptr(topk_score),
        _lib.ptr(rank), _lib.ptr(gt_score), _lib.stream_ptr()))
    out.update(topk_idx=topk_idx, topk_score=topk_score, rank=rank, gt_score=gt_score)
    return out


def topk_merge(cand_score: torch.Tensor, cand_idx: torch.Tensor, k: int):
    """Merge per-shard candidate lists [n_rows, n_cand] into the global top-k."""
    cs, ci = cand_score.to(torch.float64).contiguous(), cand_idx.to(torch.int32).contiguous()
    n_rows, n_cand = cs.shape
    oi = torch.empty((n_rows, k), dtype=torch.int32, device=cs.device)
    os_ = torch.empty((n_rows, k), dtype=torch.float64, device=cs.device)
    _lib.check(_lib.load().made_topk_merge(_lib.ptr(cs), _lib.ptr(ci), n_rows, n_cand, k, _lib.ptr(oi),
                                           _lib.ptr(os_), _lib.stream_ptr()))
    return oi, os_


def cal_distance(x, y, distance_type: str = "COS", out: Optional[torch.Tensor] = None, col_offset: int = 0):
    """modules/loss.py:30-62, COS branch only (the shipped config; L2 raises ValueError).
    Tensors in -> fp32 CUDA tensor out.  numpy arrays in (the branch calc_similarity feeds,
    loss.py:57-61) -> numpy float64 out, computed on the current CUDA device."""
    if distance_type != "COS":
        raise ValueError(f"distance_type={distance_type!r} is not supported by made_b200 (COS only)")
    as_numpy = isinstance(x, np.ndarray)
    if as_numpy:
        x, y = _to_cuda(torch.from_numpy(np.ascontiguousarray(x))), _to_cuda(torch.from_numpy(np.ascontiguousarray(y)))
    assert x.shape[1] == y.shape[1], "The second dimension of x and y must be the same."
    x, y = _f32c(x), _f32c(y)
    if out is None:
        out = torch.empty((x.shape[0], y.shape[0]), dtype=torch.float32, device=x.device)
        col_offset = 0
    _lib.check(_lib.load().made_cosine_sim(_lib.ptr(x), x.shape[0], _lib.ptr(y), y.shape[0], x.shape[1],
                                           out.data_ptr() + 4 * col_offset, out.stride(0), _lib.stream_ptr()))
    return out.cpu().numpy().astype(np.float64) if as_numpy else out


def sim_matrix_music_pooling(video_embeds: torch.Tensor, music_embeds_pooled: torch.Tensor,
                             out: Optional[torch.Tensor] = None, col_offset: int = 0) -> torch.Tensor:
    """modules/metrics.py:10-24 on a MATERIALISED pooled tensor: video_embeds [N_v,256], music_embeds_pooled
    [N_m, N_v, 256] → sims [N_v, N_m] fp32 (row-wise cosine of every video with its conditioned pooled music
    embedding).  Tensors are moved to the current CUDA device; the fused scoring path (`Engine.xpool_score`) never
    needs this function."""
    v = _f32c(_to_cuda(video_embeds))
    p = _f32c(_to_cuda(music_embeds_pooled))
    if p.dim() != 3 or p.shape[1] != v.shape[0] or p.shape[2] != v.shape[1]:
        raise ValueError(f"expected pooled [N_m, {v.shape[0]}, {v.shape[1]}], got {tuple(p.shape)}")
    if v.shape[1] != 256:
        raise ValueError("sim_matrix_music_pooling: made_b200 is built for 256-d embeddings")
    n_q, n_m = v.shape[0], p.shape[0]
    if out is None:
        out = torch.empty((n_q, n_m), dtype=torch.float32, device=v.device)
        col_offset = 0
    _lib.check(_lib.load().made_pooled_cosine(_lib.ptr(v), _lib.ptr(p), n_q, n_m, _lib.ptr(out), out.stride(0), col_offset,
                                              _lib.stream_ptr()))
    return out


def sim_matrix_video_pooling(video_embeds_pooled: torch.Tensor, music_embeds: torch.Tensor) -> torch.Tensor:
    """modules/metrics.py:26-41 (the `XA-*video*` fusions): video_embeds_pooled [N_v, N_m, 256] (every video pooled under
    the guidance of every track), music_embeds [N_m, 256] → sims [N_v, N_m] fp32.  Same kernel as
    `sim_matrix_music_pooling` with the roles of the two sides exchanged."""
    return sim_matrix_music_pooling(music_embeds, video_embeds_pooled).t()


# ---------------------------------------------------------------------------------------------
# building blocks exported for tests
# ---------------------------------------------------------------------------------------------
def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, act: int = 0, ln=None,
              out_dtype=torch.float16) -> torch.Tensor:
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    g, b = (ln if ln is not None else (None, None))
    _lib.check(_lib.load().made_gemm_f16(
        _lib.ptr(a), _lib.ptr(w), M, N, K, _lib.ptr(bias), _lib.ptr(residual), act, _lib.ptr(g), _lib.ptr(b),
        _lib.ptr(out) if out_dtype == torch.float16 else None,
        _lib.ptr(out) if out_dtype == torch.float32 else None, _lib.stream_ptr()))
    return out


def split_pair(x: torch.Tensor) -> torch.Tensor:
    """fp32 [M, K] -> fp16 [M, 2K] rows of (hi | lo): hi = fp16(x), lo = fp16(x - hi) (what the split-precision
    GEMMs consume; host-side helper for tests and packing, plain torch casts)."""
    hi = x.to(torch.float16)
    lo = (x.float() - hi.float()).to(torch.float16)
    return torch.cat([hi, lo], dim=1).contiguous()


def gemm_f16_split(a: torch.Tensor, w_pair: torch.Tensor, split: int, bias=None, residual_pair=None, act: int = 0,
                   ln=None, out_pair: bool = False) -> torch.Tensor:
    """a: [M, K] fp16 (split=1) or [M, 2K] (hi | lo) pairs (split=2); w_pair [N, 2K] pairs.
    → fp32 [M, N], or the (hi | lo) fp16 pair [M, 2N] of the result with out_pair=True."""
    N, K = w_pair.shape[0], w_pair.shape[1] // 2
    M = a.shape[0]
    out = torch.empty((M, 2 * N), dtype=torch.float16, device=a.device) if out_pair else \
        torch.empty((M, N), dtype=torch.float32, device=a.device)
    g, b = (ln if ln is not None else (None, None))
    _lib.check(_lib.load().made_gemm_f16_split(
        _lib.ptr(a), _lib.ptr(w_pair), M, N, K, split, _lib.ptr(bias), _lib.ptr(residual_pair), act, _lib.ptr(g),
        _lib.ptr(b), _lib.ptr(out) if out_pair else None, None if out_pair else _lib.ptr(out), _lib.stream_ptr()))
    return out


def gemm_f16_split_h(a: torch.Tensor, w_pair: torch.Tensor, split: int, bias=None, act: int = 0) -> torch.Tensor:
    """Split-precision GEMM with one plain fp16 output [M, N] (no row-wide epilogue): weight-stationary when it can."""
    N, K = w_pair.shape[0], w_pair.shape[1] // 2
    M = a.shape[0]
    out = torch.empty((M, N), dtype=torch.float16, device=a.device)
    _lib.check(_lib.load().made_gemm_f16_split_h(_lib.ptr(a), _lib.ptr(w_pair), M, N, K, split, _lib.ptr(bias), act,
                                                 _lib.ptr(out), _lib.stream_ptr()))
    return out


def ffn_fused(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, act: int,
              residual: Optional[torch.Tensor] = None, ln=None, pair: bool = False) -> torch.Tensor:
    """Fused 256 -> 1024 -> 256 feed-forward block.  x [M, >=256] fp16 (first 256 columns are used), residual
    [M,512] (hi | lo) pairs when pair else [M,256] fp16 → [M,512] pairs / [M,256] fp16."""
    M = x.shape[0]
    out = torch.empty((M, 512 if pair else 256), dtype=torch.float16, device=x.device)
    g, b = (ln if ln is not None else (None, None))
    _lib.check(_lib.load().made_ffn_fused(
        _lib.ptr(x), x.stride(0), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), act, _lib.ptr(residual),
        0 if residual is None else residual.stride(0), _lib.ptr(g), _lib.ptr(b), _lib.ptr(out), out.stride(0),
        1 if pair else 0, M, _lib.stream_ptr()))
    return out


def mha_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_mask: torch.Tensor) -> torch.Tensor:
    """q,k,v [B,L,256] fp16, key_mask [B,L] float (1 = valid) → [B,L,256] fp16."""
    B, L, _ = q.shape
    out = torch.empty_like(q)
    _lib.check(_lib.load().made_mha_core(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(key_mask), B, L,
                                         _lib.ptr(out), _lib.stream_ptr()))
    return out
