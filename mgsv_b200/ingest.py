"""Feature-file ingest and result export around the hot path.

`FeatureStore` is the host side of the step just before the path: it reads the reference's
pre-extracted feature files (dataloaders/dataloader_MGSV_EC_feature.py:46-75 — one `{id}.pt` tensor
per video under `vit_feature/` + `vit_mask/`, per track under `ast_feature/` + `ast_mask/`) and the
evaluation CSV (:12, :31-52) ONCE into pinned host tensors with the reference's schema, so that
`GalleryEvaluator.run(..., on_host=True)` can move the valid rows to the GPU with the copy engines.
By default the gallery carries one column per CSV row (the reference's square layout,
util_test.py:50-52) with `prev_same` marking repeated tracks for the dedup-aware rank;
`dedup_tracks=True` stores every distinct track once instead.

`save_results_json` writes the reference's `uni_save_results_json` schema (utils/util_test.py:202-226).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import config as cfg
from . import ops


def get_cw_proportion(gt_spans: torch.Tensor, max_m_duration: float = cfg.MAX_M_DURATION) -> torch.Tensor:
    """dataloader_MGSV_EC_feature.py:18-27: [n,2] (start, end) seconds → (centre, width) / max duration."""
    gt = gt_spans.clone().to(torch.float32)
    gt[:, 1] = torch.clamp(gt[:, 1], max=max_m_duration)
    return torch.stack([(gt[:, 0] + gt[:, 1]) / 2.0 / max_m_duration, (gt[:, 1] - gt[:, 0]) / max_m_duration], dim=-1)


def _load_pt(path: str) -> torch.Tensor:
    if not os.path.exists(path):
        raise FileNotFoundError(f"feature file missing: {path}")
    return torch.load(path, map_location="cpu")


class FeatureStore:
    """Pinned host copies of the query and gallery features of one evaluation CSV."""

    def __init__(self, videos: Dict[str, torch.Tensor], tracks: Dict[str, torch.Tensor], meta: Dict[str, list],
                 gt_col: torch.Tensor, prev_same: Optional[torch.Tensor]):
        self.videos, self.tracks, self.meta, self.gt_col, self.prev_same = videos, tracks, meta, gt_col, prev_same

    @staticmethod
    def _stack(paths_feat: Sequence[str], paths_mask: Sequence[str], L: int, dim: int, pin: bool):
        n = len(paths_feat)
        feats = torch.zeros((n, L, dim), dtype=torch.float32)
        masks = torch.zeros((n, L), dtype=torch.float32)
        for i, (pf, pm) in enumerate(zip(paths_feat, paths_mask)):
            f, m = _load_pt(pf), _load_pt(pm)
            if f.dim() == 3:          # some extractors save a leading batch dimension of 1
                f, m = f[0], m.reshape(-1)
            if f.shape[0] > L or f.shape[1] != dim or m.shape[0] != f.shape[0]:
                raise ValueError(f"{pf}: expected at most [{L},{dim}] features with a matching mask, "
                                 f"got {tuple(f.shape)} / {tuple(m.shape)}")
            feats[i, :f.shape[0]] = f.to(torch.float32)
            masks[i, :m.shape[0]] = m.to(torch.float32)
        if pin and torch.cuda.is_available():
            feats, masks = feats.pin_memory(), masks.pin_memory()
        return feats, masks

    @classmethod
    def from_csv(cls, csv_path: str, frame_root: str, music_root: str, max_m_duration: float = cfg.MAX_M_DURATION,
                 pin: bool = True, dedup_tracks: bool = False) -> "FeatureStore":
        """`frame_root` / `music_root` = args.frame_frozen_feature_path / args.music_frozen_feature_path.
        dedup_tracks=False keeps the reference's gallery (one column per CSV row, repeated tracks
        included); True stores every distinct track once (gt_col then indexes the distinct list)."""
        import pandas as pd
        df = pd.read_csv(csv_path)
        video_ids = [str(x) for x in df["video_id"].to_numpy()]
        music_ids_rows = [str(x) for x in df["music_id"].to_numpy()]
        if dedup_tracks:
            music_ids = list(dict.fromkeys(music_ids_rows))
            col_of = {m: i for i, m in enumerate(music_ids)}
            gt_col = np.array([col_of[m] for m in music_ids_rows], dtype=np.int32)
            prev = None
        else:
            music_ids = music_ids_rows
            prev_np, gt_col, has_dups = ops.dedup_tables(music_ids)
            prev = torch.from_numpy(prev_np) if has_dups else None
        first_row = {}
        for r, m in enumerate(music_ids_rows):
            first_row.setdefault(m, r)
        ff, fm = cls._stack([os.path.join(frame_root, "vit_feature", f"{v}.pt") for v in video_ids],
                            [os.path.join(frame_root, "vit_mask", f"{v}.pt") for v in video_ids],
                            cfg.L_V, cfg.D_VIT, pin)
        sf, sm = cls._stack([os.path.join(music_root, "ast_feature", f"{m}.pt") for m in music_ids],
                            [os.path.join(music_root, "ast_mask", f"{m}.pt") for m in music_ids],
                            cfg.L_M, cfg.D_AST, pin)
        # per-gallery-column metadata (row of the CSV that first mentions the track)
        rows = [first_row[m] for m in music_ids]
        m_dur = torch.tensor(df["music_total_duration"].to_numpy()[rows].astype(np.float32))
        gt_moment = torch.tensor(np.stack([df["music_start"].to_numpy()[rows], df["music_end"].to_numpy()[rows]],
                                          1).astype(np.float32)).reshape(-1, 1, 2)
        v_dur = torch.tensor((df["video_end"].to_numpy() - df["video_start"].to_numpy()).astype(np.float32))
        # the ground-truth moment belongs to the (video, track) ROW, not to the track
        gt_moment_rows = torch.tensor(np.stack([df["music_start"].to_numpy(), df["music_end"].to_numpy()],
                                               1).astype(np.float32)).reshape(-1, 1, 2)
        spans_target = get_cw_proportion(gt_moment_rows[:, 0], max_m_duration).reshape(-1, 1, 2)
        # per-query ground truth (the moment belongs to the CSV row): `GalleryEvaluator.run` prefers these
        videos = dict(frame_feats=ff, frame_mask=fm, v_duration=v_dur, gt_moment=gt_moment_rows,
                      m_duration=torch.tensor(df["music_total_duration"].to_numpy().astype(np.float32)))
        tracks = dict(segment_feats=sf, segment_mask=sm, m_duration=m_dur, gt_moment=gt_moment)
        meta = dict(video_ids=video_ids, music_ids=music_ids, row_music_ids=music_ids_rows,
                    gt_moment_rows=gt_moment_rows, spans_target=spans_target,
                    m_duration_rows=torch.tensor(df["music_total_duration"].to_numpy().astype(np.float32)))
        return cls(videos, tracks, meta, torch.from_numpy(np.asarray(gt_col, dtype=np.int32)), prev)


def save_results_json(ret_results_list: List[dict], loc_results_list: List[dict], iou_list, save_path: str) -> None:
    """utils/util_test.py:202-226 — same keys, rounding and clamps."""
    out = []
    for ret, loc, iou in zip(ret_results_list, loc_results_list, iou_list):
        assert ret["music_id"] == loc["music_id"]
        out.append(dict(
            video_id=loc["video_id"],
            music_id=ret["music_id"],
            topk_mids=ret["topk_music_ids"],
            gt_mid_rank=ret["rank"],
            iou=round(float(iou), 4),
            m_duration=loc["m_duration"],
            gt_st=round(loc["gt_moment"][0][0], 3),
            gt_ed=round(loc["gt_moment"][0][1], 3),
            pred_st=round(max(loc["pred_st"], 0), 3),
            pred_ed=round(min(loc["pred_ed"], 240), 3),
        ))
    with open(save_path, "w") as f:
        json.dump(out, f, indent=4)
