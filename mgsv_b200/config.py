"""Shipped MaDe configuration (the ONE config this build targets).

Values follow the reference launch script scripts/test_kuai_all_feature.sh:14-57,84-101 and the
derived fields of test-MaDe.py:146-171 (max_snippet_num = int(max_m_duration / stride) = 96).
`check_args` mirrors the reference's guard style (ValueError for unsupported flag combinations,
model_Uni.py:275; test-MaDe.py:159,162): there is no multi-backend dispatch here, so any args
Namespace that asks for a branch outside the shipped config is rejected loudly.
"""
from __future__ import annotations

import argparse
import math

D_MODEL = 256          # dim_input == hidden_dim == detr_hidden_dim
D_VIT = 512            # CLIP ViT-B/32 frame feature width
D_AST = 768            # AST segment feature width
L_V = 50               # max_v_frames
L_M = 96               # max_snippet_num = int(240 / 2.5)
L_DETR = L_V + L_M     # concat fusion: 146 memory tokens
N_HEADS = 8
D_HEAD = 32
D_FF = 1024
DETR_ENC_LAYERS = 2
DETR_DEC_LAYERS = 6
MAX_M_DURATION = 240.0
STRIDE = 2.5
TEMPERATURE_INIT = 0.03
LN_EPS = 1e-5
XPOOL_G_COLS = 112     # per-track X-Pool operand [G (96) | W5 (5) | 0]: csrc/xpool.cu

SHIPPED = dict(
    dim_input=256, hidden_dim=256, detr_hidden_dim=256,
    video_attention_seqlen=250,
    video_transformer_depth=1, audio_transformer_depth=1, SA_temporal_heads=8,
    agg_module="transf", with_cls_token=0, with_last_token=0, with_act_after_proj=0,
    transformer_is_share=0,
    max_v_frames=50, max_m_duration=240, stride=2.5, max_snippet_num=96,
    vmr_fusion="XA-music", vmr_loss="dual_single_loss_fuse", dual_single_loss_weight=1.0,
    fusion_mask=1, mml_fusion="concat", mml_localization="detr", moment_query_type="video",
    num_moment_queries=1, decoder_SA=0, predict_center=0,
    detr_dropout=0.1, detr_nheads=8, detr_dim_feedforward=1024, detr_enc_layers=2,
    detr_dec_layers=6, detr_pre_norm=False, position_embedding="sine", input_dropout=0.5,
    span_loss_type="l1", fb_label="01", l1_loss=1, aux_loss=1, contrastive_align_loss=1,
    moment_loss=0, audio_short_cut=0, contrastive_dim=256, temperature_init_value=0.03,
    ignore_same_music=1,
)

# fields that only the drivers read (test-MaDe.py:245-440); given defaults so that a bare
# Namespace from `default_args()` can be handed to the reference's eval_epoch too.
DRIVER_DEFAULTS = dict(
    name="made_b200", local_rank=0,
    frozen_feature_path="", music_frozen_feature_path="", frame_frozen_feature_path="",
    audio_encoder_type="feature", video_encoder_type="feature",
    do_eval=True, num_display=1, tb_writer=0, epochs=1, ret_loss_weight=1.0,
    loc_loss_weight=1.0, toph_moment=1, distance_type="COS", save_json=0,
    test_csv="", path_log="",
)

# which args decide the compute graph; everything else is ignored by the hot path
_STRICT = (
    "dim_input", "hidden_dim", "video_transformer_depth", "audio_transformer_depth",
    "SA_temporal_heads", "agg_module", "with_cls_token", "with_act_after_proj",
    "transformer_is_share", "max_v_frames", "max_snippet_num", "vmr_fusion", "vmr_loss",
    "fusion_mask", "mml_fusion", "mml_localization", "moment_query_type",
    "num_moment_queries", "predict_center", "detr_nheads", "detr_dim_feedforward",
    "detr_enc_layers", "detr_dec_layers", "detr_pre_norm", "position_embedding",
    "span_loss_type", "fb_label", "aux_loss", "contrastive_align_loss", "moment_loss",
    "audio_short_cut", "contrastive_dim",
)


def default_args(**overrides) -> argparse.Namespace:
    d = dict(SHIPPED)
    d.update(DRIVER_DEFAULTS)
    d.update(overrides)
    return argparse.Namespace(**d)


# flag values accepted beside the shipped one: vmr_fusion "XA-music-video" adds a second Transformer_XA whose output the
# shipped vmr_loss never reads (model_Uni.py:203-204, 254-262), so the compute graph that produces outputs is unchanged
# mml_fusion "CA" (the argparse default of test-MaDe.py:80; the shipped script passes "concat") puts a CrossTransformer
# between the encoders and DETR: built (`made_ca_fuse`), fp16 tcgen05 GEMMs + a CUDA-core cross-attention kernel.
_ALSO = {"vmr_fusion": ("XA-music-video",), "mml_fusion": ("CA",)}


def check_args(args) -> None:
    """Raise ValueError unless `args` selects the shipped compute graph."""
    for k in _STRICT:
        if not hasattr(args, k):
            raise ValueError(f"args.{k} is required by made_b200 (shipped value {SHIPPED[k]!r})")
        got, want = getattr(args, k), SHIPPED[k]
        same = (float(got) == float(want)) if isinstance(want, (int, float)) and not isinstance(want, bool) \
            and isinstance(got, (int, float)) else (got == want)
        if not same and got not in _ALSO.get(k, ()):
            raise ValueError(
                f"Error: args.{k}={got!r} is not supported by made_b200 "
                f"(only the shipped MaDe config is built: {k}={want!r})")
    if int(float(args.max_m_duration) / float(args.stride)) != L_M:
        raise ValueError("max_m_duration/stride must give max_snippet_num=96")


def logit_scale_init() -> float:
    """model_Uni.py:29 — ln(1 / temperature_init_value)."""
    return math.log(1.0 / TEMPERATURE_INIT)
