"""Evaluation-time losses returned by Uni_model.forward (model_Uni.py:254-262, 278-289): host glue
around made_detr_losses / made_retrieval_loss."""
from __future__ import annotations

import torch

from . import _lib
from . import config as cfg

_NAMES = ("loss_span", "loss_giou", "loss_label", "class_error", "loss_contrastive_align")


def detr_losses(pred_logits, pred_spans, proj_queries, proj_vid_mem, targets_cw, empty_weight, temperature=0.07):
    """pred_* [n_layers,B,2] → tensor [n_layers,5] in _NAMES order."""
    n_layers, B = pred_logits.shape[0], pred_logits.shape[1]
    out = torch.empty((n_layers, 5), dtype=torch.float32, device=pred_logits.device)
    tg = targets_cw.reshape(B, 2).to(torch.float32).contiguous()
    ew = [float(x) for x in empty_weight.tolist()]
    _lib.check(_lib.load().made_detr_losses(
        _lib.ptr(pred_logits.contiguous()), _lib.ptr(pred_spans.contiguous()),
        None if proj_queries is None else _lib.ptr(proj_queries.contiguous()),
        None if proj_vid_mem is None else _lib.ptr(proj_vid_mem.contiguous()),
        _lib.ptr(tg), B, n_layers, ew[0], ew[1], temperature, _lib.ptr(out), _lib.stream_ptr()))
    return out


def retrieval_loss(dual, single, logit_scale: float):
    out = torch.empty(1, dtype=torch.float32, device=dual.device)
    n = dual.shape[0]
    if dual.shape != (n, n) or single.shape != (n, n) or dual.stride(0) != single.stride(0):
        raise ValueError("retrieval_loss needs two square in-batch similarity matrices with one layout")
    _lib.check(_lib.load().made_retrieval_loss(_lib.ptr(dual), _lib.ptr(single), dual.stride(0), n,
                                               float(logit_scale), _lib.ptr(out), _lib.stream_ptr()))
    return out[0]


def clip_loss(sims, logit_scale: float):
    """CLIPLoss (modules/loss.py:5-24): symmetric cross entropy of sims * exp(logit_scale) with diagonal labels,
    forward value only.  CLIPLoss and InfoNCELoss (audio_id=None) are the same formula, and made_retrieval_loss
    returns their sum over its two arguments: L(s) = (L(s) + L(s)) / 2, exact in floating point."""
    s = sims.to(torch.float32)
    if s.stride(1) != 1:
        s = s.contiguous()
    return retrieval_loss(s, s, logit_scale) * 0.5


def info_nce_loss(sims, logit_scale: float):
    """InfoNCELoss (modules/loss.py:66-123, the audio_id=None path) → (loss, logits_per_video, logits_per_audio)."""
    logits = sims * float(torch.tensor(float(logit_scale)).exp())
    return clip_loss(sims, logit_scale), logits, logits.t()


def eval_losses(model, output_map, single, dual, spans_target):
    """→ loss_map with the reference's keys: retrieval_loss, localization_loss,
    localization_loss_dict (30 entries: 5 names x (final + 5 aux suffixes))."""
    L = cfg.DETR_DEC_LAYERS
    logits = torch.stack([a["pred_logits"][:, 0] for a in output_map["aux_outputs"]] + [output_map["pred_logits"][:, 0]])
    spans = torch.stack([a["pred_spans"][:, 0] for a in output_map["aux_outputs"]] + [output_map["pred_spans"][:, 0]])
    pq = torch.stack([a["proj_queries"][:, 0] for a in output_map["aux_outputs"]] + [output_map["proj_queries"][:, 0]])
    crit = model.criterion
    vals = detr_losses(logits, spans, pq, output_map["proj_vid_mem"], spans_target[:, 0, :], crit.empty_weight,
                       crit.temperature)
    loss_dict = {}
    for j, name in enumerate(_NAMES):
        loss_dict[name] = vals[L - 1, j]
    for i in range(L - 1):
        for j, name in enumerate(_NAMES):
            loss_dict[f"{name}_{i}"] = vals[i, j]
    wd = crit.weight_dict
    loc = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)      # model_Uni.py:289
    ret = retrieval_loss(dual, single, float(model.logit_scale))
    return {"retrieval_loss": ret, "localization_loss": loc, "localization_loss_dict": loss_dict}
