"""Host-side mirror of the reference's model interface for the inference + scoring hot path.

`Uni_model(args, device, logger)` keeps the reference's constructor, `forward` signature, returned
5-tuple, attribute names read by the drivers and state_dict key names (model/model_Uni.py:14-322,
SURVEY.md §8b), but owns no compute: every tensor op is a C-ABI call into libmade_b200.so through
`Engine`.  Parameters are held as ordinary nn.Parameters (fp32 masters, reference names) so that
`.to()`, `.float()`, `.eval()`, `.state_dict()`, `.load_state_dict()` and the param-group getters
behave; the engine repacks them (fp16 operands, folded X-Pool/decoder weights) whenever they change.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops, synth
from . import config as cfg
from .engine import Engine


class _Node(nn.Module):
    """Container that only holds parameters/buffers under the reference's names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("made_b200 parameter containers are not callable; use Uni_model.forward")


def _ensure_path(root: nn.Module, parts: List[str]) -> nn.Module:
    node = root
    for p in parts:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    return node


class _Criterion(_Node):
    """Stands in for SetCriterion (loss_detr.py:47-56): the driver reads `foreground_label`
    (test-MaDe.py:307); `weight_dict` as loss_detr.py:36-45."""

    def __init__(self):
        super().__init__()
        self.foreground_label = 0
        self.background_label = 1
        self.eos_coef = 0.1
        self.temperature = 0.07
        base = {"loss_span": 4, "loss_giou": 1, "loss_label": 0.8, "loss_contrastive_align": 0.2}
        self.weight_dict = dict(base)
        for i in range(cfg.DETR_DEC_LAYERS - 1):
            self.weight_dict.update({f"{k}_{i}": v for k, v in base.items()})


class _XPoolView(_Node):
    """`model.video_guided_to_music_pooling_cross_transformer`: parameter holder + thin
    compatibility callable (test-MaDe.py:392-395 calls it and moves it between devices)."""

    def __init__(self, owner: "Uni_model", which: int = _lib.MUSIC):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        object.__setattr__(self, "_which", which)

    def forward(self, video_embeds, music_embeds, music_mask=None):
        """(For the second module of "XA-music-video" read: guides = music_feats, keys = frame features, [N_v,N_m,256].)
        Transformer_XA.forward (modules/transformer.py:156-180): video_embeds [N_v,256], music_embeds
        [N_m,96,256] (the encoded segments), music_mask [N_m,96] → the MATERIALISED [N_m,N_v,256] fp32 tensor, on the
        model's CUDA device, computed in the reference's fp32 arithmetic (`made_xpool_pooled`).  Inputs may live on
        the CPU (test-MaDe.py:386-394 gathers them there); they are moved.  Above MADE_POOLED_CAP_GB (default 32 GiB
        of output) this raises and points at `score_gallery`, which never forms the tensor."""
        return self._owner._xpool_pooled(video_embeds, music_embeds, music_mask, self._which)

    # test-MaDe.py:392/395 moves this sub-module to the CPU and back around the gallery stage.  The view computes on
    # the model's CUDA device whatever the location of the fp32 master parameters, so moving is accepted and is a
    # deliberate no-op (the parameters stay registered where the model holds them; state_dict is unchanged).
    def cpu(self):
        return self

    def cuda(self, device=None):
        return self

    def to(self, *a, **k):
        return self


class Uni_model(nn.Module):
    def __init__(self, args, device=None, logger=None):
        super().__init__()
        cfg.check_args(args)
        self.args = args
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.logger = logger
        self.dim_input = cfg.D_MODEL
        self.num_moment_queries = 1
        self.aux_loss = 1
        # parameter tree under the reference's names
        self.add_module("video_guided_to_music_pooling_cross_transformer", _XPoolView(self))
        self.add_module("criterion", _Criterion())
        self.xa_video = "video" in str(args.vmr_fusion)         # model_Uni.py:27-28: a second Transformer_XA
        if self.xa_video:
            self.add_module("music_guided_to_video_pooling_cross_transformer", _XPoolView(self, _lib.VIDEO))
        self.ca_fusion = "CA" in str(args.mml_fusion)           # model_Uni.py:33-43: CrossTransformer before DETR
        sd0 = None
        for key, shape, kind in synth.state_dict_spec() + (synth.xa_video_spec() if self.xa_video else []) + \
                (synth.ca_spec() if self.ca_fusion else []):
            parts = key.split(".")
            node = _ensure_path(self, parts[:-1])
            if kind in ("pe", "empty_weight"):
                if sd0 is None:
                    sd0 = synth.make_state_dict(0)
                node.register_buffer(parts[-1], sd0[key].clone())
            else:
                node.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))
        self._engine: Optional[Engine] = None
        self._packed_version = None
        self.reset_parameters()

    # -- parameters ------------------------------------------------------------------------------
    def reset_parameters(self, seed: int = 0):
        sd = synth.make_state_dict(seed, xa_video=self.xa_video, ca=self.ca_fusion)
        with torch.no_grad():
            for k, v in self.state_dict().items():
                v.copy_(sd[k])
        self._packed_version = None

    def load_state_dict(self, state_dict, strict: bool = True):
        sd = {k: v for k, v in state_dict.items() if not k.startswith(("vit_model.", "ast_model."))}
        r = super().load_state_dict(sd, strict=strict)
        self._packed_version = None
        return r

    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def engine(self) -> Engine:
        if self._engine is None:
            dev = self.device if self.device.type == "cuda" else torch.device("cuda")
            self._engine = Engine(dev)
        ver = self._param_version()
        if self._packed_version != ver:
            self._engine.load_state_dict(self.state_dict())
            self._engine.mml_fusion = "CA" if self.ca_fusion else "concat"
            self._packed_version = ver
        return self._engine

    def _group(self, prefixes):
        return [p for n, p in self.named_parameters() if n.startswith(prefixes)]

    def get_temporal_parameter(self):       # model_Uni.py:73-77
        return self._group(("vit_proj.", "ast_proj.", "video_transformer.", "audio_transformer."))

    def get_matching_parameter(self):       # model_Uni.py:80-89
        return self._group(("video_guided_to_music_pooling_cross_transformer.",
                            "music_guided_to_video_pooling_cross_transformer.")) + [self.logit_scale]

    def get_detection_parameter(self):      # model_Uni.py:92-114
        return self._group(("video_music_fusion_cross_transformer.", "detr_transformer.", "span_embed.", "class_embed.",
                            "contrastive_align_projection_"))

    # -- pieces ----------------------------------------------------------------------------------
    def forward_video_encoder_feature(self, frame_feats=None, frame_masks=None, video_ids=None):
        """model_Base.py:544-581 → (frame_feats [B,50,256], video_feats [B,256], frame_masks)."""
        dev = self.engine().device
        seq, seq32, pooled = self.engine().encode(_lib.VIDEO, frame_feats.to(dev), frame_masks.to(dev))
        self._last_frame16 = seq
        return seq32, pooled, frame_masks

    def forward_audio_encoder_feature(self, segment_feats=None, segment_masks=None, music_ids=None):
        """model_Base.py:583-617."""
        dev = self.engine().device
        seq, seq32, pooled = self.engine().encode(_lib.MUSIC, segment_feats.to(dev), segment_masks.to(dev))
        self._last_segment16 = seq
        return seq32, pooled, segment_masks

    def score_gallery(self, video_feats, music_feats, segment_feats, segment_masks, out=None, col_offset=0):
        """Fused replacement of test-MaDe.py:392-403: → (single [N_v,N_m] f32, dual [N_v,N_m] f32).
        The reference's final score is double(single) + double(dual) (`ops.rank_topk` forms it)."""
        eng = self.engine()
        dev = eng.device
        seg = segment_feats.to(dev)
        seg16 = seg if seg.dtype == torch.float16 else seg.to(torch.float16)
        kz, gram, bits = eng.gallery_prepare(seg16, segment_masks.to(dev))
        q, vhat = eng.query_prepare(video_feats.to(dev))
        single = eng.xpool_score(q, vhat, kz, gram, bits)
        dual = ops.cal_distance(video_feats.to(dev), music_feats.to(dev))
        return single, dual

    def _xpool_pooled(self, video_embeds, music_embeds, music_mask, which=_lib.MUSIC):
        L_keys = cfg.L_M if which == _lib.MUSIC else cfg.L_V
        if music_mask is None:
            raise ValueError("Error: fusion_mask=0 (unmasked X-Pool) is not supported by made_b200 (shipped: fusion_mask=1)")
        if video_embeds.dim() != 2 or music_embeds.dim() != 3 or music_embeds.shape[1:] != (L_keys, cfg.D_MODEL) \
                or video_embeds.shape[1] != cfg.D_MODEL or tuple(music_mask.shape) != tuple(music_embeds.shape[:2]):
            raise ValueError(f"expected video_embeds [N_v,{cfg.D_MODEL}], music_embeds [N_m,{L_keys},{cfg.D_MODEL}], "
                             f"music_mask [N_m,{L_keys}]; got {tuple(video_embeds.shape)}, {tuple(music_embeds.shape)}, "
                             f"{tuple(music_mask.shape)}")
        n_m, n_v = music_embeds.shape[0], video_embeds.shape[0]
        cap = float(os.environ.get("MADE_POOLED_CAP_GB", "32")) * (1 << 30)
        if n_m * n_v * cfg.D_MODEL * 4 > cap:
            raise RuntimeError(
                f"the materialised pooled tensor [{n_m}, {n_v}, 256] fp32 is {n_m * n_v * 1024 / 2**30:.1f} GiB "
                f"(cap MADE_POOLED_CAP_GB = {cap / 2**30:.0f}); call model.score_gallery(video_feats, music_feats, "
                "segment_feats, segment_masks), which scores the gallery without forming it (INTEGRATION.md)")
        return self.engine().xpool_pooled(video_embeds, music_embeds, music_mask, which=which)

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, frame_feats, segment_feats, frame_masks, segment_masks, spans_target, v_duration=None,
                video_ids=None, music_ids=None, is_train=False):
        """model_Uni.py:177-322 (inference: is_train must be False; there is no backward here)."""
        if is_train:
            raise ValueError("made_b200 implements the inference/scoring path only (is_train=False)")
        eng = self.engine()
        dev = eng.device
        frame_masks = frame_masks.to(dev)
        segment_masks = segment_masks.to(dev)
        frame_out, video_feats, _ = self.forward_video_encoder_feature(frame_feats, frame_masks)
        segment_out, music_feats, _ = self.forward_audio_encoder_feature(segment_feats, segment_masks)
        if self.ca_fusion:      # model_Uni.py:209-213: the segments attend to the paired video's frames; DETR sees 96 tokens
            fused16, _ = eng.ca_fuse(segment_out, segment_masks, frame_out, frame_masks)
            det = eng.detr_detect(self._last_frame16, torch.zeros_like(frame_masks), fused16, segment_masks, video_feats,
                                  want_proj=True)
        else:
            det = eng.detr_detect(self._last_frame16, frame_masks, self._last_segment16, segment_masks,
                                  video_feats, want_proj=True)
        L = cfg.DETR_DEC_LAYERS
        output_map = {
            "pred_logits": det["pred_logits"][L - 1].unsqueeze(1),
            "pred_spans": det["pred_spans"][L - 1].unsqueeze(1),
            "proj_queries": det["proj_queries"][L - 1].unsqueeze(1),
            "proj_vid_mem": det["proj_vid_mem"],
            "aux_outputs": [{
                "pred_logits": det["pred_logits"][i].unsqueeze(1),
                "pred_spans": det["pred_spans"][i].unsqueeze(1),
                "proj_queries": det["proj_queries"][i].unsqueeze(1),
                "proj_vid_mem": det["proj_vid_mem"],
            } for i in range(L - 1)],
        }
        single, dual = self.score_gallery(video_feats, music_feats, self._last_segment16, segment_masks)
        loss_map = self._eval_losses(output_map, single, dual, spans_target.to(dev))
        feat_map = {"video_feats": video_feats, "music_feats": music_feats, "frame_feats": frame_out,
                    "segment_feats": segment_out}
        mask_map = {"frame_masks": frame_masks, "segment_masks": segment_masks}
        id_map = {"video_ids": video_ids, "music_ids": music_ids}
        return output_map, loss_map, feat_map, mask_map, id_map

    def _eval_losses(self, output_map, single, dual, spans_target):
        from . import losses
        return losses.eval_losses(self, output_map, single, dual, spans_target)
