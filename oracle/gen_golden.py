"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Runs only in the build container (the GPU box has no /root/reference); the fixtures it writes
are committed.  Inputs and weights are NOT stored: both sides regenerate them from
`mgsv_b200.synth` seeds.  Recipe = SURVEY.md §8(c): stub `clip`/`wget`/`timm`, encoder types
other than "AST"/"ViT" and hand-attached vit_proj/ast_proj, gloo process group for the driver.

    python oracle/gen_golden.py            # writes tests/golden/{forward_b8,cfg1_256,span_pairs}.npz
"""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MADE_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from mgsv_b200 import config as C  # noqa: E402
from mgsv_b200 import synth  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def _stub_modules():
    for name in ["clip", "wget", "timm", "timm.models", "timm.models.layers"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["timm"].__version__ = "0.4.5"
    sys.modules["timm.models.layers"].to_2tuple = lambda x: (x, x)
    sys.modules["timm.models.layers"].trunc_normal_ = lambda *a, **k: None
    if REF not in sys.path:
        sys.path.insert(0, REF)


def build_reference_model(sd, **overrides):
    _stub_modules()
    from model.model_Uni import Uni_model
    args = C.default_args(name="oracle", **overrides)
    logger = logging.getLogger("gen_golden")
    model = Uni_model(args, torch.device("cpu"), logger)
    model.vit_proj = nn.Linear(C.D_VIT, C.D_MODEL)   # model_Base.py:289 (not built for feature input)
    model.ast_proj = nn.Linear(C.D_AST, C.D_MODEL)   # model_Base.py:282
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.float().eval()
    return model, args


def dup_tracks(tracks, ids, n_dup, src0=0, dst0=None):
    """Make tracks[dst0+j] an exact copy (features + id) of tracks[src0+j]: the reference gallery
    repeats a track once per video that uses it (util_test.py:40-60 dedups them)."""
    n = tracks["segment_feats"].shape[0]
    dst0 = n - n_dup if dst0 is None else dst0
    for j in range(n_dup):
        for k in ("segment_feats", "segment_mask", "m_duration", "gt_moment", "spans_target", "n_segments"):
            tracks[k][dst0 + j] = tracks[k][src0 + j]
        ids["music_ids"][dst0 + j] = ids["music_ids"][src0 + j]


def t2n(x):
    return x.detach().cpu().numpy()


@torch.no_grad()
def gen_forward_b8(model, sd):
    B = 8
    v, m, ids = synth.make_eval_set(B, B, synth.BASE_SEED + 100)
    out, loss, feat, mask, _ = model(
        v["frame_feats"].clone(), m["segment_feats"].clone(), v["frame_mask"], m["segment_mask"],
        m["spans_target"], v_duration=v["v_duration"], video_ids=ids["video_ids"],
        music_ids=ids["music_ids"], is_train=False)
    g = {
        "pred_logits": t2n(out["pred_logits"]), "pred_spans": t2n(out["pred_spans"]),
        "proj_queries": t2n(out["proj_queries"]), "proj_vid_mem": t2n(out["proj_vid_mem"]),
        "video_feats": t2n(feat["video_feats"]), "music_feats": t2n(feat["music_feats"]),
        "frame_feats": t2n(feat["frame_feats"]), "segment_feats": t2n(feat["segment_feats"]),
        "retrieval_loss": t2n(loss["retrieval_loss"]),
        "localization_loss": t2n(loss["localization_loss"]),
    }
    for i, aux in enumerate(out["aux_outputs"]):
        g[f"aux{i}_pred_logits"] = t2n(aux["pred_logits"])
        g[f"aux{i}_pred_spans"] = t2n(aux["pred_spans"])
    names = sorted(loss["localization_loss_dict"].keys())
    g["loss_names"] = np.array(names)
    g["loss_values"] = np.array([float(loss["localization_loss_dict"][k]) for k in names], dtype=np.float64)
    # pieces used by kernel-level parity tests
    pooled = model.video_guided_to_music_pooling_cross_transformer(
        feat["video_feats"], feat["segment_feats"], mask["segment_masks"])
    g["xpool_pooled"] = t2n(pooled)
    pos = model.music_position_embedding(
        torch.cat([feat["frame_feats"], feat["segment_feats"]], 1),
        torch.cat([mask["frame_masks"], mask["segment_masks"]], 1))
    g["detr_pos"] = t2n(pos[:2])
    np.savez_compressed(os.path.join(GOLD, "forward_b8.npz"), **g)
    print("forward_b8:", {k: getattr(v_, "shape", None) for k, v_ in g.items()})


@torch.no_grad()
def gen_forward_b8_ca():
    """mml_fusion "CA": the unmodified reference with its CrossTransformer fusion, same B = 8 batch."""
    sd = synth.make_state_dict(0, ca=True)
    model, args = build_reference_model(sd, mml_fusion="CA")
    B = 8
    v, m, ids = synth.make_eval_set(B, B, synth.BASE_SEED + 100)
    out, loss, feat, mask, _ = model(
        v["frame_feats"].clone(), m["segment_feats"].clone(), v["frame_mask"], m["segment_mask"],
        m["spans_target"], v_duration=v["v_duration"], video_ids=ids["video_ids"],
        music_ids=ids["music_ids"], is_train=False)
    fused, _ = model.video_music_fusion_cross_transformer(feat["segment_feats"], feat["frame_feats"],
                                                          q_mask=mask["segment_masks"], kv_mask=mask["frame_masks"])
    fused = fused.masked_fill(mask["segment_masks"].unsqueeze(-1) == 0, 0)       # model_Uni.py:210
    g = {"pred_logits": t2n(out["pred_logits"]), "pred_spans": t2n(out["pred_spans"]),
         "proj_queries": t2n(out["proj_queries"]), "fused": t2n(fused),
         "retrieval_loss": t2n(loss["retrieval_loss"]), "localization_loss": t2n(loss["localization_loss"])}
    for i, aux in enumerate(out["aux_outputs"]):
        g[f"aux{i}_pred_spans"] = t2n(aux["pred_spans"])
    np.savez_compressed(os.path.join(GOLD, "forward_b8_ca.npz"), **g)
    print("forward_b8_ca:", {k: getattr(v_, "shape", None) for k, v_ in g.items()})


def _load_driver():
    """Import test-MaDe.py unmodified; it calls init_process_group('nccl') at import (:25)."""
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    real_init = dist.init_process_group

    def gloo_init(backend=None, *a, **k):
        return real_init("gloo", *a, **k)
    dist.init_process_group = gloo_init
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        spec = importlib.util.spec_from_file_location("test_made_driver", os.path.join(REF, "test-MaDe.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
        dist.init_process_group = real_init
    mod.logger = logging.getLogger("gen_golden.driver")
    return mod


@torch.no_grad()
def gen_cfg1(model, args, sd):
    """cfg 1: 64 queries x 256 tracks, run square 256x256 as the reference requires (Q13)."""
    N = 256
    v, m, ids = synth.make_eval_set(N, N, synth.BASE_SEED + 1)
    dup_tracks(m, ids, n_dup=16)
    bs = 32
    batches = []
    for s in range(0, N, bs):
        e = s + bs
        data_map = dict(frame_feats=v["frame_feats"][s:e].clone(), frame_mask=v["frame_mask"][s:e],
                        segment_feats=m["segment_feats"][s:e].clone(), segment_mask=m["segment_mask"][s:e])
        meta_map = dict(video_id=ids["video_ids"][s:e], music_id=ids["music_ids"][s:e],
                        gt_moment=m["gt_moment"][s:e], m_duration=m["m_duration"][s:e],
                        v_duration=v["v_duration"][s:e])
        batches.append((data_map, meta_map, m["spans_target"][s:e]))

    # (i) the unmodified driver
    drv = _load_driver()
    loss_avg, ret_m, loc_m, com_m = drv.eval_epoch(1, args, model, batches, torch.device("cpu"))

    # (ii) the same stage through the reference's functions, keeping the tensors
    from modules.metrics import sim_matrix_music_pooling
    from utils.util_test import calc_similarity, Recall_metrics, IoU_metrics, Composite_metrics
    from music_detr.span_utils import span_cw_to_se, detr_iou
    import torch.nn.functional as F
    vfl, afl, seg, segm, mr = [], [], [], [], []
    pst, ped, psc = [], [], []
    for data_map, meta_map, spans_target in batches:
        out, loss, feat, mask, _ = model(
            data_map["frame_feats"].clone(), data_map["segment_feats"].clone(), data_map["frame_mask"],
            data_map["segment_mask"], spans_target, v_duration=meta_map["v_duration"],
            video_ids=meta_map["video_id"], music_ids=meta_map["music_id"], is_train=False)
        vfl.append(feat["video_feats"]), afl.append(feat["music_feats"])
        seg.append(feat["segment_feats"]), segm.append(mask["segment_masks"])
        prob = F.softmax(out["pred_logits"], dim=-1)[:, :, model.criterion.foreground_label]
        for i in range(prob.shape[0]):
            spans = span_cw_to_se(out["pred_spans"][i]) * args.max_m_duration
            ranked = torch.cat((spans, prob[i].unsqueeze(-1)), dim=-1)
            ranked = sorted(ranked, key=lambda x: x[2], reverse=True)[:1]
            mr.append(dict(gt_moment=meta_map["gt_moment"][i], m_duration=meta_map["m_duration"][i],
                           ranked_preds=ranked))
            pst.append(float(ranked[0][0])), ped.append(float(ranked[0][1])), psc.append(float(ranked[0][2]))
    video_embeds = torch.cat(vfl)
    pooled = model.video_guided_to_music_pooling_cross_transformer(video_embeds, torch.cat(seg), torch.cat(segm))
    single = sim_matrix_music_pooling(video_embeds, pooled).numpy()
    dual = calc_similarity([x.numpy() for x in vfl], [x.numpy() for x in afl], distance_type="COS")
    total = single * 1.0 + dual * 1.0
    ret2, ind, res = Recall_metrics(total, dedup=True, all_music_ids_list=ids["music_ids"])
    iou_list = detr_iou(args, mr)
    loc2 = IoU_metrics(iou_list)
    com2 = Composite_metrics(ind, iou_list, mr, ids["video_ids"], ids["music_ids"])
    assert abs(ret2["MeanR"] - ret_m["MeanR"]) < 1e-9 and abs(float(loc2["mIoU"]) - float(loc_m["mIoU"])) < 1e-7

    def md(d):
        keys = sorted(k for k in d if k != "cols")
        return np.array(keys), np.array([float(d[k]) for k in keys], dtype=np.float64)
    g = dict(
        single=single[:64].astype(np.float32), dual=dual[:64].astype(np.float64), total=total[:64],
        ind=np.asarray(ind, dtype=np.int64), top1=np.array([r["topk_music_ids"][0] for r in res]),
        pred_st=np.array(pst, dtype=np.float32), pred_ed=np.array(ped, dtype=np.float32),
        pred_score=np.array(psc, dtype=np.float32),
        iou=np.array([float(i) for i in iou_list], dtype=np.float32),
        video_feats=video_embeds.numpy(), music_feats=torch.cat(afl).numpy(),
        driver_loss_avg=np.float64(float(loss_avg)),
    )
    for name, d in (("ret", ret_m), ("loc", loc_m), ("com", com_m)):
        g[f"{name}_keys"], g[f"{name}_vals"] = md(d)
    np.savez_compressed(os.path.join(GOLD, "cfg1_256.npz"), **g)
    print("cfg1_256: R1=%.2f MeanR=%.2f mIoU=%.4f" % (ret_m["R1"], ret_m["MeanR"], float(loc_m["mIoU"])))


def gen_span_pairs():
    _stub_modules()
    from music_detr.span_utils import generalized_temporal_iou, temporal_iou, span_cw_to_se
    t1 = torch.Tensor([[0, 0.2], [0.5, 1.0]])
    t2 = torch.Tensor([[0, 0.3], [0., 1.0]])
    iou, union = temporal_iou(t1, t2)
    a, b, logits = synth.make_span_pairs(64, 48, synth.BASE_SEED + 3)
    giou = generalized_temporal_iou(span_cw_to_se(a), span_cw_to_se(b))
    prob = logits.softmax(-1)
    keep = b[:, 1] != 0                                   # matcher.py:59
    tgt = b[keep]
    cost_class = -prob[:, torch.full([len(tgt)], 0)]
    cost_span = torch.cdist(a.float(), tgt.float(), p=1)
    cost_giou = -generalized_temporal_iou(span_cw_to_se(a), span_cw_to_se(tgt))
    Cm = 10 * cost_span + 1 * cost_giou + 4 * cost_class   # matcher.py:88 with build_matcher weights
    np.savez_compressed(
        os.path.join(GOLD, "span_pairs.npz"),
        doctest_iou=iou.numpy(), doctest_union=union.numpy(),
        doctest_giou=generalized_temporal_iou(t1, t2).numpy(),
        giou=giou.numpy(), prob_fg=prob[:, 0].numpy(), cost=Cm.numpy(), l1=cost_span.numpy())
    print("span_pairs: giou", giou.shape, "cost", Cm.shape, "nan count", int(torch.isnan(giou).sum()))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLD, exist_ok=True)
    sd = synth.make_state_dict(0)
    model, args = build_reference_model(sd)
    gen_span_pairs()
    gen_forward_b8(model, sd)
    gen_cfg1(model, args, sd)
    gen_forward_b8_ca()


if __name__ == "__main__":
    main()
