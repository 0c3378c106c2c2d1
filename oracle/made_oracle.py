"""CPU ORACLE — test infrastructure only, never shipped, never on the product path.

A functional fp32 restatement (torch CPU / numpy) of the MaDe inference + scoring hot path of
xxayt/MGSV for the shipped config, each function citing the reference file:line
it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this package.

Pinned: `tests/golden/*.npz` hold outputs of the UNMODIFIED reference source run in the build
container (torch 2.11 CPU fp32) through `oracle/gen_golden.py`; `tests/test_oracle_golden.py`
checks this restatement against them (plus the two doctest vectors of
music_detr/span_utils.py:48-54,99-103).

Weights arrive as a flat `sd` dict with the reference's state_dict key names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

D = 256
N_HEADS = 8
MAX_M_DURATION = 240.0
XPOOL = "video_guided_to_music_pooling_cross_transformer"


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def _linear(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[f"{prefix}.weight"], sd[f"{prefix}.bias"])


def _ln(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (D,), sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], 1e-5)


def mha(sd: SD, prefix: str, query, key, value, key_padding_mask=None, n_heads=N_HEADS):
    """torch.nn.MultiheadAttention forward (batch_first=False layout restated batch-first),
    as called at model_Base.py:87 and music_detr/transformer.py:196,284,289.
    query [B,Lq,D], key/value [B,Lk,D], key_padding_mask [B,Lk] bool True = ignore.
    """
    w, b = sd[f"{prefix}.in_proj_weight"], sd[f"{prefix}.in_proj_bias"]
    B, Lq, _ = query.shape
    Lk = key.shape[1]
    dh = D // n_heads
    q = F.linear(query, w[:D], b[:D]).view(B, Lq, n_heads, dh).transpose(1, 2)
    k = F.linear(key, w[D:2 * D], b[D:2 * D]).view(B, Lk, n_heads, dh).transpose(1, 2)
    v = F.linear(value, w[2 * D:], b[2 * D:]).view(B, Lk, n_heads, dh).transpose(1, 2)
    scores = torch.matmul(q * (1.0 / math.sqrt(dh)), k.transpose(-1, -2))  # [B,H,Lq,Lk]
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    attn = torch.softmax(scores, dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).reshape(B, Lq, D)
    return F.linear(o, sd[f"{prefix}.out_proj.weight"], sd[f"{prefix}.out_proj.bias"])


# ---------------------------------------------------------------------------------------------
# temporal encoders  (model_Base.py:544-617, 520-542, 82-91, 48-60)
# ---------------------------------------------------------------------------------------------
def temporal_encoder(sd: SD, feats, masks, proj: str, tr: str, pe_key: str):
    """forward_{video,audio}_encoder_feature.  feats [B,L,Din], masks [B,L] float {0,1}.
    Returns (seq_feats [B,L,256], pooled [B,256])."""
    L = feats.shape[1]
    x = feats.masked_fill(masks.unsqueeze(-1) == 0, 0)            # :556 / :595
    x = _linear(sd, proj, x)                                      # :559 / :598
    x = x + sd[pe_key][:, :L]                                     # :533 (all positions, Q3)
    # Transformer_enhancement.forward :82-91, depth 1; residuals from the NORMED tensor (Q2)
    x = _ln(sd, f"{tr}.layers.0.0", x)
    x = mha(sd, f"{tr}.layers.0.1", x, x, x, key_padding_mask=~(masks.bool())) + x
    x = _ln(sd, f"{tr}.layers.0.2", x)
    h = F.gelu(_linear(sd, f"{tr}.layers.0.3.0", x))              # nn.GELU() = erf form
    x = _linear(sd, f"{tr}.layers.0.3.3", h) + x
    x = _linear(sd, f"{tr}.final_linear", x)                      # :91
    x = x.masked_fill(masks.unsqueeze(-1) == 0, 0)                # :541
    pooled = x.sum(dim=1) / masks.sum(dim=1).unsqueeze(-1)        # :579 / :615
    pooled = F.normalize(pooled, p=2, dim=-1)                     # :580 / :616
    return x, pooled


def encode_video(sd: SD, frame_feats, frame_masks):
    return temporal_encoder(sd, frame_feats, frame_masks, "vit_proj", "video_transformer",
                            "video_position_embedding.pe")


def encode_music(sd: SD, segment_feats, segment_masks):
    return temporal_encoder(sd, segment_feats, segment_masks, "ast_proj", "audio_transformer",
                            "audio_position_embedding.pe")


# ---------------------------------------------------------------------------------------------
# X-Pool  (modules/transformer.py:156-180, 87-123) and similarities
# ---------------------------------------------------------------------------------------------
def xpool(sd: SD, video_embeds, music_embeds, music_mask):
    """Transformer_XA.forward → [N_m, N_v, 256] (materialised, like the reference)."""
    v = _ln(sd, f"{XPOOL}.layer_norm1", video_embeds)             # :164
    s = _ln(sd, f"{XPOOL}.layer_norm1", music_embeds)             # :165 (shared LN1, Q4)
    q = _linear(sd, f"{XPOOL}.cross_attn.q_proj", v)              # [N_v,D]
    k = _linear(sd, f"{XPOOL}.cross_attn.k_proj", s)              # [N_m,L,D]
    val = _linear(sd, f"{XPOOL}.cross_attn.v_proj", s)
    logits = torch.matmul(q.unsqueeze(0), k.transpose(-1, -2))    # [N_m,N_v,L]  (1 head)
    logits = logits / math.sqrt(D)                                # :111  (head_dim = 256)
    logits = logits.masked_fill(music_mask[:, None, :] == 0, float("-inf"))  # :116
    w = torch.softmax(logits, dim=-1)
    a = torch.matmul(w, val)                                      # [N_m,N_v,D]
    o = _linear(sd, f"{XPOOL}.cross_attn.out_proj", a)            # :122
    o = _ln(sd, f"{XPOOL}.layer_norm2", o)                        # :174
    out = o + _linear(sd, f"{XPOOL}.linear_proj", o)              # :176-177 (dropout off)
    return _ln(sd, f"{XPOOL}.layer_norm3", out)                   # :178


def sim_matrix_music_pooling(video_embeds, music_embeds_pooled):
    """modules/metrics.py:10-24 → [N_v, N_m]."""
    v = video_embeds / video_embeds.norm(dim=-1, keepdim=True)
    p = music_embeds_pooled / music_embeds_pooled.norm(dim=-1, keepdim=True)
    return torch.einsum("vd,mvd->vm", v, p)


def cal_distance_cos(x, y):
    """modules/loss.py:52-56 (torch branch)."""
    x = x / x.norm(p=2, dim=1, keepdim=True)
    y = y / y.norm(p=2, dim=1, keepdim=True)
    return torch.matmul(x, y.t())


def gallery_similarity(sd: SD, video_feats, music_feats, segment_feats, segment_masks,
                       track_chunk: int = 64):
    """test-MaDe.py:386-403: single (X-Pool cosine) + dual (tower cosine), summed in float64.
    Chunked over tracks so the [N_m,N_v,256] intermediate stays small; per-(v,m) arithmetic is
    unchanged.  Returns (single f32 [N_v,N_m], dual f32, total f64 numpy)."""
    singles = []
    for s in range(0, segment_feats.shape[0], track_chunk):
        pooled = xpool(sd, video_feats, segment_feats[s:s + track_chunk],
                       segment_masks[s:s + track_chunk])
        singles.append(sim_matrix_music_pooling(video_feats, pooled))
    single = torch.cat(singles, dim=1)
    # calc_similarity feeds numpy arrays → numpy branch of cal_distance (loss.py:57-61)
    x = video_feats.numpy()
    y = music_feats.numpy()
    x = x / np.linalg.norm(x, axis=1, keepdims=True)
    y = y / np.linalg.norm(y, axis=1, keepdims=True)
    dual = np.matmul(x, y.T)
    total = single.numpy() * 1.0 + dual.astype(np.float64) * 1.0   # test-MaDe.py:403
    return single, torch.from_numpy(dual), total


# ---------------------------------------------------------------------------------------------
# DETR  (music_detr/position_encoding.py:51-71, music_detr/transformer.py:51-81,191-210,273-307)
# ---------------------------------------------------------------------------------------------
def position_embedding_sine(mask, num_pos_feats: int = D, temperature: float = 10000.0):
    x_embed = mask.cumsum(1, dtype=torch.float32)
    x_embed = x_embed / (x_embed[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, None] / dim_t
    return torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)


def detr_encoder_layer(sd: SD, p: str, src, pad_mask, pos):
    q = k = src + pos
    src2 = mha(sd, f"{p}.self_attn", q, k, src, key_padding_mask=pad_mask)
    src = _ln(sd, f"{p}.norm1", src + src2)
    src2 = _linear(sd, f"{p}.linear2", F.relu(_linear(sd, f"{p}.linear1", src)))
    return _ln(sd, f"{p}.norm2", src + src2)


def detr_decoder_layer(sd: SD, p: str, tgt, memory, pad_mask, pos, query_pos):
    # self-attention always runs (Q1: build_transformer drops `args`, transformer.py:325-335)
    q = k = tgt + query_pos
    tgt2 = mha(sd, f"{p}.self_attn", q, k, tgt)
    tgt = _ln(sd, f"{p}.norm1", tgt + tgt2)
    tgt2 = mha(sd, f"{p}.multihead_attn", tgt + query_pos, memory + pos, memory,
               key_padding_mask=pad_mask)
    tgt = _ln(sd, f"{p}.norm2", tgt + tgt2)
    tgt2 = _linear(sd, f"{p}.linear2", F.relu(_linear(sd, f"{p}.linear1", tgt)))
    return _ln(sd, f"{p}.norm3", tgt + tgt2)


def detr_forward(sd: SD, src, valid_mask, pos, target, n_enc=2, n_dec=6):
    """Transformer.forward, batch-first.  src [B,146,256]; valid_mask [B,146] float (1 = valid);
    target [B,1,256].  Returns hs [n_dec,B,1,256] (each through decoder.norm), memory."""
    pad = ~(valid_mask.bool())
    memory = src
    for i in range(n_enc):
        memory = detr_encoder_layer(sd, f"detr_transformer.encoder.layers.{i}", memory, pad, pos)
    query_pos = sd["decoder_query_embed.weight"].unsqueeze(0).expand(src.shape[0], -1, -1)
    out = target
    hs = []
    for i in range(n_dec):
        out = detr_decoder_layer(sd, f"detr_transformer.decoder.layers.{i}", out, memory, pad,
                                 pos, query_pos)
        hs.append(_ln(sd, "detr_transformer.decoder.norm", out))
    return torch.stack(hs), memory


def span_mlp(sd: SD, x):
    """MLP(256,256,2,3) music_detr/transformer.py:348-360."""
    x = F.relu(_linear(sd, "span_embed.layers.0", x))
    x = F.relu(_linear(sd, "span_embed.layers.1", x))
    return _linear(sd, "span_embed.layers.2", x)


def calc_output(sd: SD, hs, frame_feats):
    """model_Uni.py:117-173 for the shipped flags."""
    outputs_class = _linear(sd, "class_embed", hs)
    outputs_coord = span_mlp(sd, hs).sigmoid()
    proj_queries = F.normalize(_linear(sd, "contrastive_align_projection_query", hs), p=2, dim=-1)
    proj_vid_mem = F.normalize(_linear(sd, "contrastive_align_projection_vid", frame_feats),
                               p=2, dim=-1)
    out = {
        "pred_logits": outputs_class[-1], "pred_spans": outputs_coord[-1],
        "proj_queries": proj_queries[-1], "proj_vid_mem": proj_vid_mem,
        "aux_outputs": [
            {"pred_logits": a, "pred_spans": b, "proj_queries": c, "proj_vid_mem": proj_vid_mem}
            for a, b, c in zip(outputs_class[:-1], outputs_coord[:-1], proj_queries[:-1])],
    }
    return out


# ---------------------------------------------------------------------------------------------
# span utils / matcher / criterion (music_detr/span_utils.py, matcher.py, loss_detr.py)
# ---------------------------------------------------------------------------------------------
def span_cw_to_se(cw):
    """span_utils.py:15-24."""
    return torch.stack([cw[:, 0] - 0.5 * cw[:, 1], cw[:, 0] + 0.5 * cw[:, 1]], dim=-1)


def temporal_iou(s1, s2):
    """span_utils.py:39-66."""
    a1 = s1[:, 1] - s1[:, 0]
    a2 = s2[:, 1] - s2[:, 0]
    left = torch.max(s1[:, None, 0], s2[:, 0])
    right = torch.min(s1[:, None, 1], s2[:, 1])
    inter = (right - left).clamp(min=0)
    union = a1[:, None] + a2 - inter
    return inter / union, union


def generalized_temporal_iou(s1, s2):
    """span_utils.py:86-115."""
    s1 = s1.float()
    s2 = s2.float()
    assert (s1[:, 1] >= s1[:, 0]).all()
    assert (s2[:, 1] >= s2[:, 0]).all()
    iou, union = temporal_iou(s1, s2)
    left = torch.min(s1[:, None, 0], s2[:, 0])
    right = torch.max(s1[:, None, 1], s2[:, 1])
    enc = (right - left).clamp(min=0)
    return iou - (enc - union) / enc


def matcher_cost(prob_fg, out_spans_cw, tgt_spans_cw, w_span=10.0, w_giou=1.0, w_class=4.0):
    """matcher.py:66-88 given the foreground probabilities.
    C = 10*L1(cw) + 1*(-gIoU(se)) + 4*(-p_fg), evaluated left to right."""
    cost_class = -prob_fg[:, None].expand(-1, tgt_spans_cw.shape[0])
    cost_span = torch.cdist(out_spans_cw.float(), tgt_spans_cw.float(), p=1)
    cost_giou = -generalized_temporal_iou(span_cw_to_se(out_spans_cw), span_cw_to_se(tgt_spans_cw))
    return w_span * cost_span + w_giou * cost_giou + w_class * cost_class


def hungarian_indices(pred_logits, pred_spans, targets):
    """HungarianMatcher.forward matcher.py:36-92 (scipy LSAP per sample)."""
    from scipy.optimize import linear_sum_assignment
    bs, nq = pred_spans.shape[:2]
    prob = pred_logits.flatten(0, 1).softmax(-1)
    moment_mask = targets[:, :, 1] != 0
    tgt = targets[moment_mask]
    sizes = moment_mask.sum(dim=1).tolist()
    Cm = matcher_cost(prob[:, 0], pred_spans.flatten(0, 1), tgt).view(bs, nq, -1)
    idx = [linear_sum_assignment(c[i]) for i, c in enumerate(Cm.split(sizes, -1))]
    return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64))
            for i, j in idx]


def _src_idx(indices):
    b = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)])
    s = torch.cat([s for (s, _) in indices])
    return b, s


def criterion_losses(outputs, targets, empty_weight, temperature=0.07):
    """SetCriterion.forward for one output dict (loss_detr.py:74-128), fb_label '01'."""
    indices = hungarian_indices(outputs["pred_logits"], outputs["pred_spans"], targets)
    idx = _src_idx(indices)
    losses = {}
    src_spans = outputs["pred_spans"][idx]
    tgt_spans = torch.cat([t[i] for t, (_, i) in zip(targets, indices)], dim=0)
    losses["loss_span"] = F.l1_loss(src_spans, tgt_spans, reduction="none").mean()
    losses["loss_giou"] = (1 - torch.diag(generalized_temporal_iou(
        span_cw_to_se(src_spans), span_cw_to_se(tgt_spans)))).mean()
    logits = outputs["pred_logits"]
    target_classes = torch.full(logits.shape[:2], 1, dtype=torch.int64)
    target_classes[idx] = 0
    losses["loss_label"] = F.cross_entropy(logits.transpose(1, 2), target_classes, empty_weight,
                                           reduction="none").mean()
    sel = logits[idx]
    pred = sel.topk(1, 1, True, True)[1].t()
    losses["class_error"] = 100 - pred.eq(0).view(-1).float().sum(0) * (100.0 / sel.size(0))
    lg = torch.einsum("bmd,bnd->bmn", outputs["proj_queries"], outputs["proj_vid_mem"])
    lg = lg.sum(2) / temperature
    pos_map = torch.zeros_like(lg, dtype=torch.bool)
    pos_map[idx] = True
    pos_term = lg.masked_fill(~pos_map, 0).sum(1)
    losses["loss_contrastive_align"] = (-pos_term / pos_map.sum(1) + lg.logsumexp(1)).mean()
    return losses


WEIGHT_DICT = {"loss_span": 4, "loss_giou": 1, "loss_label": 0.8, "loss_contrastive_align": 0.2}


def set_criterion(sd: SD, output_map, targets):
    loss_map = criterion_losses(output_map, targets, sd["criterion.empty_weight"])
    for i, aux in enumerate(output_map["aux_outputs"]):
        loss_map.update({f"{k}_{i}": v for k, v in
                         criterion_losses(aux, targets, sd["criterion.empty_weight"]).items()})
    return loss_map


def localization_loss(loss_dict):
    """model_Uni.py:288-289 with loss_detr.py:36-45 weights (aux copies share weights)."""
    tot = 0
    for k, v in loss_dict.items():
        base = k.rsplit("_", 1)[0] if k.rsplit("_", 1)[-1].isdigit() else k
        if base in WEIGHT_DICT:
            tot = tot + v * WEIGHT_DICT[base]
    return tot


def clip_loss(sims, logit_scale):
    """modules/loss.py:5-24."""
    logits = sims * logit_scale.exp()
    t2v = -torch.diag(F.log_softmax(logits, dim=1)).mean()
    v2t = -torch.diag(F.log_softmax(logits, dim=0)).mean()
    return (t2v + v2t) / 2.0


def info_nce_loss(sims, logit_scale):
    """modules/loss.py:66-123 with audio_id=None (Q7)."""
    lv = sims * logit_scale.exp()
    lab = torch.arange(lv.shape[0])
    return (F.cross_entropy(lv, lab) + F.cross_entropy(lv.t(), lab)) / 2


# ---------------------------------------------------------------------------------------------
# mml_fusion "CA": CrossTransformer (model/model_Base.py:169-213), CrossAttention (:93-165), FeedForward (:22-46)
# ---------------------------------------------------------------------------------------------
CA = "video_music_fusion_cross_transformer"


def cross_transformer(sd: SD, query, context, q_mask, kv_mask, heads: int = 8):
    """depth 1.  query [B,Lq,256], context [B,Lk,256], masks float {0,1} → [B,Lq,256] (before the caller's masked_fill)."""
    x = query
    nx = _ln(sd, f"{CA}.attention_query_layer_norms.0", x)                     # :203
    nc = _ln(sd, f"{CA}.attention_context_layer_norms.0", context)            # :204
    q = F.linear(nx, sd[f"{CA}.layers.0.0.to_q.weight"])                       # :135 (no bias)
    k, v = F.linear(nc, sd[f"{CA}.layers.0.0.to_kv.weight"]).chunk(2, dim=-1)  # :136
    B, Lq, inner = q.shape
    dh = inner // heads
    sp = lambda t: t.reshape(B, t.shape[1], heads, dh).permute(0, 2, 1, 3)     # 'b n (h d) -> b h n d'
    q, k, v = sp(q), sp(k), sp(v)
    dots = torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5                   # :140
    dots = dots.masked_fill(kv_mask[:, None, None, :] == 0, float("-inf"))     # :158 (kv_mask before the softmax)
    attn = torch.softmax(dots, dim=-1)
    attn = attn.masked_fill(q_mask[:, None, :, None] == 0, 0)                  # :160 (q_mask after the softmax)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, Lq, inner)      # :162-163
    x_res = _linear(sd, f"{CA}.layers.0.0.to_out.0", out)                      # :164
    attn_x = x_res + x                                                         # :206
    nf = _ln(sd, f"{CA}.ff_layer_norms.0", attn_x)                             # :207
    ff = _linear(sd, f"{CA}.layers.0.1.net.3", F.gelu(_linear(sd, f"{CA}.layers.0.1.net.0", nf)))
    x = ff + attn_x                                                            # :208
    return _linear(sd, f"{CA}.final_linear", x)                                # :210


# ---------------------------------------------------------------------------------------------
# Uni_model.forward  (model_Uni.py:177-322)
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def uni_forward(sd: SD, frame_feats, segment_feats, frame_masks, segment_masks, spans_target,
                video_ids=None, music_ids=None, with_losses: bool = True, mml_fusion: str = "concat"):
    frame_out, video_feats = encode_video(sd, frame_feats, frame_masks)
    segment_out, music_feats = encode_music(sd, segment_feats, segment_masks)
    pooled = xpool(sd, video_feats, segment_out, segment_masks)              # :201
    if mml_fusion == "CA":                                                   # :209-211
        src = cross_transformer(sd, segment_out, frame_out, segment_masks, frame_masks)
        src = src.masked_fill(segment_masks.unsqueeze(-1) == 0, 0)
        mask = segment_masks
    else:
        src = torch.cat([frame_out, segment_out], dim=1)                     # :207
        mask = torch.cat([frame_masks, segment_masks], dim=1)
    pos = position_embedding_sine(mask)                                      # :216
    hs, memory = detr_forward(sd, src, mask, pos, video_feats.unsqueeze(1))  # :218-227
    output_map = calc_output(sd, hs, frame_out)
    loss_map = {}
    if with_losses:
        dual = cal_distance_cos(video_feats, music_feats)
        single = sim_matrix_music_pooling(video_feats, pooled)
        loss_map["retrieval_loss"] = info_nce_loss(dual, sd["logit_scale"]) * 1.0 + \
            clip_loss(single, sd["logit_scale"]) * 1.0                       # :255-262
        ld = set_criterion(sd, output_map, spans_target)
        loss_map["localization_loss_dict"] = ld
        loss_map["localization_loss"] = localization_loss(ld)
    feat_map = dict(video_feats=video_feats, music_feats=music_feats, frame_feats=frame_out,
                    segment_feats=segment_out)
    mask_map = dict(frame_masks=frame_masks, segment_masks=segment_masks)
    id_map = dict(video_ids=video_ids, music_ids=music_ids)
    output_map["hs"] = hs
    output_map["memory"] = memory
    return output_map, loss_map, feat_map, mask_map, id_map


# ---------------------------------------------------------------------------------------------
# driver post-processing + metrics  (test-MaDe.py:306-330; span_utils.py:119-170;
# utils/util_test.py:32-199)
# ---------------------------------------------------------------------------------------------
def moment_postproc(pred_logits, pred_spans):
    """test-MaDe.py:306-316 with 1 query: fg score, (st, ed) seconds."""
    prob = F.softmax(pred_logits, dim=-1)
    score = prob[:, 0, 0]
    se = span_cw_to_se(pred_spans[:, 0, :]) * MAX_M_DURATION
    return se[:, 0], se[:, 1], score


def detr_iou(pred_st, pred_ed, gt_moment, m_duration):
    """span_utils.py:147-170 + individual_IoU_tensor :119-145, vectorised (fp32)."""
    pred_st = torch.clamp(pred_st, min=0)
    pred_ed = torch.clamp(pred_ed, max=MAX_M_DURATION)
    gt_st, gt_ed = gt_moment[:, 0, 0], gt_moment[:, 0, 1]
    pred_st = torch.clamp(pred_st, min=0)
    pred_ed = torch.minimum(pred_ed, m_duration)
    inter = torch.clamp(torch.min(gt_ed, pred_ed) - torch.max(gt_st, pred_st), min=0)
    union = (pred_ed - pred_st) + (gt_ed - gt_st) - inter
    iou = inter / union
    bad = (gt_st >= gt_ed) | (union <= 0)
    return torch.where(bad, torch.zeros_like(iou), iou)


def recall_metrics(sim_matrix: np.ndarray, music_ids: List[str], gt_cols: Optional[np.ndarray] = None):
    """Recall_metrics(dedup=True) util_test.py:32-97.  Row i's GT id is music_ids[gt_cols[i]]
    (gt_cols defaults to arange → the reference's square case, Q13)."""
    n_rows = sim_matrix.shape[0]
    if gt_cols is None:
        gt_cols = np.arange(n_rows)
    sort_indices = np.argsort(sim_matrix, axis=1)[:, ::-1]
    ind = []
    top1 = []
    for i in range(n_rows):
        gt_id = music_ids[gt_cols[i]]
        seen = set()
        for idx in sort_indices[i]:
            mid = music_ids[idx]
            if mid not in seen:
                seen.add(mid)
                if mid == gt_id:
                    ind.append(len(seen) - 1)
                    break
        top1.append(music_ids[sort_indices[i][0]])
    ind = np.array(ind)
    return summarize_ranks(ind), ind, top1


def summarize_ranks(ind: np.ndarray):
    """util_test.py:81-96."""
    m = {}
    for k in (1, 3, 5, 10, 20, 25, 50, 100):
        m[f"R{k}"] = float(np.sum(ind < k)) * 100 / len(ind)
    m["MedianR"] = np.median(ind) + 1
    m["MeanR"] = np.mean(ind) + 1
    m["cols"] = [int(i) for i in list(ind)]
    m["MRR"] = np.mean(1.0 / (ind + 1))
    return m


def iou_metrics(iou_list):
    """util_test.py:101-111 (strict >, Q9).  The reference's list holds 0-d fp32 tensors, so the
    thresholds compare in fp32 and python's sum() accumulates sequentially in fp32."""
    t = [torch.as_tensor(float(i), dtype=torch.float32) for i in iou_list]
    n = len(t)
    return {
        "mIoU": float(sum(t) / n),
        "IoU@0.3": sum(1 for i in t if i > 0.3) * 100 / n,
        "IoU@0.5": sum(1 for i in t if i > 0.5) * 100 / n,
        "IoU@0.7": sum(1 for i in t if i > 0.7) * 100 / n,
    }


def composite_metrics(rank_list, iou_list):
    """util_test.py:140-199 including the double division of R*_miou (Q8)."""
    keys = [f"R{r}_{s}" for s in ("iou0.5", "iou0.7", "miou") for r in (1, 10, 50, 100)]
    m = {k: 0.0 for k in keys}
    num = {1: 0, 10: 0, 50: 0, 100: 0}
    for r0, iou in zip(rank_list, iou_list):
        rank = r0 + 1
        iou = float(iou)
        for r in (1, 10, 50, 100):
            if rank <= r:
                m[f"R{r}_iou0.5"] += iou > 0.5
                m[f"R{r}_iou0.7"] += iou > 0.7
                m[f"R{r}_miou"] += iou
                num[r] += 1
    for k in m:
        m[k] /= len(rank_list)
        if "0." in k:
            m[k] *= 100
    for r in (1, 10, 50, 100):
        m[f"R{r}_miou"] = m[f"R{r}_miou"] / num[r] if num[r] > 0 else 0.0
    return m


# ---------------------------------------------------------------------------------------------
# whole job (the bench "step" on the CPU side)
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def evaluate(sd: SD, videos: dict, tracks: dict, music_ids: List[str], track_chunk: int = 64,
             batch: int = 40):
    """One pass of the hot path: encode queries + gallery, full-gallery scoring, ranks,
    moment detection for the paired track (query i ↔ track i), IoU and metrics
    (test-MaDe.py:243-447 with the per-batch criterion/log work left out)."""
    n_q = videos["frame_feats"].shape[0]
    fo, vf = [], []
    for s in range(0, n_q, batch):
        a, b = encode_video(sd, videos["frame_feats"][s:s + batch], videos["frame_mask"][s:s + batch])
        fo.append(a), vf.append(b)
    frame_out, video_feats = torch.cat(fo), torch.cat(vf)
    so, mf = [], []
    n_m = tracks["segment_feats"].shape[0]
    for s in range(0, n_m, batch):
        a, b = encode_music(sd, tracks["segment_feats"][s:s + batch], tracks["segment_mask"][s:s + batch])
        so.append(a), mf.append(b)
    segment_out, music_feats = torch.cat(so), torch.cat(mf)
    single, dual, total = gallery_similarity(sd, video_feats, music_feats, segment_out,
                                             tracks["segment_mask"], track_chunk)
    ret, ind, top1 = recall_metrics(total, music_ids, np.arange(n_q))
    sts, eds, scs = [], [], []
    for s in range(0, n_q, batch):
        e = min(s + batch, n_q)
        src = torch.cat([frame_out[s:e], segment_out[s:e]], dim=1)
        mask = torch.cat([videos["frame_mask"][s:e], tracks["segment_mask"][s:e]], dim=1)
        hs, _ = detr_forward(sd, src, mask, position_embedding_sine(mask), video_feats[s:e].unsqueeze(1))
        om = calc_output(sd, hs, frame_out[s:e])
        st, ed, sc = moment_postproc(om["pred_logits"], om["pred_spans"])
        sts.append(st), eds.append(ed), scs.append(sc)
    pred_st, pred_ed, score = torch.cat(sts), torch.cat(eds), torch.cat(scs)
    iou = detr_iou(pred_st, pred_ed, tracks["gt_moment"][:n_q], tracks["m_duration"][:n_q])
    return dict(single=single, dual=dual, total=total, ind=ind, top1=top1, ret=ret,
                pred_st=pred_st, pred_ed=pred_ed, score=score, iou=iou,
                loc=iou_metrics(list(iou.numpy())), com=composite_metrics(list(ind), list(iou.numpy())),
                video_feats=video_feats, music_feats=music_feats)
