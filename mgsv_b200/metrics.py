"""Host-side metric summaries with the reference's names and semantics (utils/util_test.py).

The O(N_v x N_m) work (fp64 score sum, dedup-aware rank, top-k, IoU) runs in CUDA kernels
(`ops.rank_topk`, `ops.moment_postproc`); what is left here is the reference's own O(N_v) numpy
bookkeeping over per-query integers/floats that were copied back from the device.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops


def summarize_ranks(ind: np.ndarray) -> Dict[str, float]:
    """util_test.py:81-96: R@k (% of ind < k), MedianR/MeanR (+1), MRR, cols."""
    ind = np.asarray(ind)
    m = {}
    for k in (1, 3, 5, 10, 20, 25, 50, 100):
        m[f"R{k}"] = float(np.sum(ind < k)) * 100 / len(ind)
    m["MedianR"] = np.median(ind) + 1
    m["MeanR"] = np.mean(ind) + 1
    m["cols"] = [int(i) for i in list(ind)]
    m["MRR"] = np.mean(1.0 / (ind + 1))
    return m


def calc_similarity(video_feat_list, audio_feat_list, distance_type: str = "COS") -> np.ndarray:
    """utils/util_test.py:10-29 — same arguments (lists of [bs, dim] blocks, tensors or numpy arrays) and
    return value: the [val_len_v, val_len_m] numpy matrix, float32 for tensor blocks and float64 for
    numpy blocks as modules/loss.py:52-61 produces them.  The reference loops over block pairs; rows
    are normalised independently, so ONE cosine launch over the concatenated blocks gives the same
    matrix."""
    if len(video_feat_list) == 0 or len(audio_feat_list) == 0:
        raise ValueError("need at least one array to concatenate")     # np.concatenate(()) in the reference
    as_numpy = isinstance(video_feat_list[0], np.ndarray)
    if as_numpy:
        x = np.concatenate([np.asarray(b) for b in video_feat_list], axis=0)
        y = np.concatenate([np.asarray(b) for b in audio_feat_list], axis=0)
        return ops.cal_distance(x, y, distance_type)
    x = torch.cat([b.detach() for b in video_feat_list], dim=0)
    y = torch.cat([b.detach() for b in audio_feat_list], dim=0)
    x = ops._to_cuda(x)
    return ops.cal_distance(x, y.to(x.device), distance_type).cpu().numpy()


def Recall_metrics(sim_single: torch.Tensor, sim_dual: Optional[torch.Tensor] = None, distance_type: str = "COS",
                   dedup: bool = True, all_music_ids_list: Optional[Sequence[str]] = None,
                   gt_music_ids: Optional[Sequence[str]] = None):
    """utils/util_test.py:32-97 on device similarity matrices.  `sim_single` (+ optional `sim_dual`,
    summed in float64 as test-MaDe.py:403) is [N_v, N_m] fp32 on the GPU.  Returns
    (metrics, ind, ret_results_list) like the reference."""
    if distance_type != "COS":
        raise ValueError("only COS similarities are supported")
    n_rows, n_cols = sim_single.shape
    ids = list(all_music_ids_list) if all_music_ids_list is not None else [str(i) for i in range(n_cols)]
    gt_ids = list(gt_music_ids) if gt_music_ids is not None else ids[:n_rows]
    prev, gt_col, has_dups = ops.dedup_tables(ids, gt_ids)
    if (gt_col < 0).any():
        raise ValueError("a ground-truth music id is missing from the gallery id list")
    dev = sim_single.device
    r = ops.rank_topk(sim_single, sim_dual, torch.from_numpy(gt_col).to(dev),
                      torch.from_numpy(prev).to(dev) if (dedup and has_dups) else None, k=1)
    ind = r["rank"].cpu().numpy().astype(np.int64)
    top1 = r["topk_idx"][:, 0].cpu().numpy()
    results = [dict(music_id=gt_ids[i], rank=int(ind[i]) + 1, topk_music_ids=[ids[top1[i]]]) for i in range(n_rows)]
    return summarize_ranks(ind), ind, results


def Recall_metrics_matrix(sim_matrix, distance_type: str = "COS", dedup: bool = False, all_music_ids_list=None):
    """utils/util_test.py:32-97 with the reference's own signature: `sim_matrix` is the [val_len, val_len] host
    matrix test-MaDe.py:403 builds (float64 numpy = f64(single) + f64(dual); float32 numpy or a tensor work too).
    The matrix is moved to the current CUDA device as an (fp32 hi, fp32 lo) pair whose float64 sum the rank kernel
    forms — exact for float32 input, and within 2^-48 relative of a float64 entry (far below the spacing of any two
    scores that differ at all in their float32 parts), so the ordering is the reference's argsort ordering up to
    exact ties.  Returns (metrics, ind, ret_results_list) like the reference; `dedup=False` ranks by column."""
    if isinstance(sim_matrix, torch.Tensor):
        x = sim_matrix.detach()
    else:
        x = torch.from_numpy(np.ascontiguousarray(sim_matrix))
    if x.dim() != 2:
        raise ValueError("sim_matrix must be [val_len, val_len]")
    x = ops._to_cuda(x)
    if x.dtype == torch.float64:
        hi = x.to(torch.float32)
        lo = (x - hi.to(torch.float64)).to(torch.float32)
    else:
        hi, lo = x.to(torch.float32).contiguous(), None
    ids = all_music_ids_list if (all_music_ids_list is not None and len(all_music_ids_list) > 0) else None
    return Recall_metrics(hi.contiguous(), None if lo is None else lo.contiguous(), distance_type,
                          dedup=bool(dedup and ids is not None), all_music_ids_list=ids)


def _f32(values) -> np.ndarray:
    if isinstance(values, torch.Tensor):
        return values.detach().to("cpu", torch.float32).numpy().reshape(-1)
    return np.asarray([float(i) for i in values], dtype=np.float32)


def IoU_metrics(IoU_list) -> Dict[str, float]:
    """util_test.py:101-111.  The reference holds 0-d fp32 tensors: thresholds compare in fp32
    (strict >) and `sum()` accumulates sequentially in fp32."""
    v = _f32(IoU_list)
    n = len(v)
    th = lambda t: float((v > np.float32(t)).sum()) * 100 / n
    return {"mIoU": float(np.cumsum(v, dtype=np.float32)[-1] / np.float32(n)), "IoU@0.3": th(0.3),
            "IoU@0.5": th(0.5), "IoU@0.7": th(0.7)}


def Composite_metrics(ret_rank_list, IoU_list, mr_results_list=None, all_video_ids_list=None,
                      all_music_ids_list=None) -> Dict[str, float]:
    """util_test.py:140-199 including the double division of R*_miou."""
    ranks = np.asarray(ret_rank_list).astype(np.int64) + 1
    iou = _f32(IoU_list)
    n = len(ranks)
    m = {}
    for r in (1, 10, 50, 100):
        sel = ranks <= r
        m[f"R{r}_iou0.5"] = float((iou[sel] > np.float32(0.5)).sum()) / n * 100
        m[f"R{r}_iou0.7"] = float((iou[sel] > np.float32(0.7)).sum()) / n * 100
        cnt = int(sel.sum())
        tot = float(np.cumsum(iou[sel], dtype=np.float32)[-1]) if cnt else 0.0
        m[f"R{r}_miou"] = (tot / n) / cnt if cnt > 0 else 0.0
    order = [f"R{r}_{s}" for s in ("iou0.5", "iou0.7", "miou") for r in (1, 10, 50, 100)]
    return {k: m[k] for k in order}
