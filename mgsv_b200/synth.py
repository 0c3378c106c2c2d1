"""Deterministic synthetic weights and MGSV-EC-shaped inputs.

Everything is drawn from numpy's PCG64 so the same seed gives bit-identical tensors in this
container and on the GPU box (the golden fixtures under tests/golden/ only store OUTPUTS; inputs
and weights are regenerated from the seed on both sides).

State-dict key names and shapes are the reference's (listed by instantiating
model/model_Uni.py:14 with the shipped config) so that a reference checkpoint loads into the
B200 model and this synthetic one loads (strict) into the reference model.

Length statistics: the reference's dataset/MGSV-EC/test_data.csv gives n_frames in [6,50]
(mean 23.1, median 21) and n_segments in [13,96] (mean 56.5, bimodal around 14-26 and 62-93),
with the moment width equal to the video duration (corr 0.9996). Those summary figures are
reproduced by a small parametric model below; no dataset rows are stored in this repo.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import config as C

BASE_SEED = 20250817  # S = 20250817 + config_index


# ---------------------------------------------------------------------------------------------
# state dict
# ---------------------------------------------------------------------------------------------
def _mha(prefix: str, kind: str) -> List[Tuple[str, tuple, str]]:
    return [
        (f"{prefix}.in_proj_weight", (768, 256), kind),
        (f"{prefix}.in_proj_bias", (768,), "bias"),
        (f"{prefix}.out_proj.weight", (256, 256), kind),
        (f"{prefix}.out_proj.bias", (256,), "bias"),
    ]


def _ln(prefix: str) -> List[Tuple[str, tuple, str]]:
    return [(f"{prefix}.weight", (256,), "ln_w"), (f"{prefix}.bias", (256,), "ln_b")]


def _lin(prefix: str, out_f: int, in_f: int, kind: str = "linear") -> List[Tuple[str, tuple, str]]:
    return [(f"{prefix}.weight", (out_f, in_f), kind), (f"{prefix}.bias", (out_f,), "bias")]


def state_dict_spec() -> List[Tuple[str, tuple, str]]:
    """(key, shape, init kind) in the reference's state_dict order."""
    s: List[Tuple[str, tuple, str]] = []
    s.append(("logit_scale", (), "logit_scale"))
    for i in range(C.DETR_ENC_LAYERS):
        p = f"detr_transformer.encoder.layers.{i}"
        s += _mha(f"{p}.self_attn", "xavier")
        s += _lin(f"{p}.linear1", 1024, 256, "xavier") + _lin(f"{p}.linear2", 256, 1024, "xavier")
        s += _ln(f"{p}.norm1") + _ln(f"{p}.norm2")
    for i in range(C.DETR_DEC_LAYERS):
        p = f"detr_transformer.decoder.layers.{i}"
        s += _mha(f"{p}.self_attn", "xavier") + _mha(f"{p}.multihead_attn", "xavier")
        s += _lin(f"{p}.linear1", 1024, 256, "xavier") + _lin(f"{p}.linear2", 256, 1024, "xavier")
        s += _ln(f"{p}.norm1") + _ln(f"{p}.norm2") + _ln(f"{p}.norm3")
    s += _ln("detr_transformer.decoder.norm")
    s.append(("video_position_embedding.pe", (1, 250, 256), "pe"))
    s.append(("audio_position_embedding.pe", (1, 300, 256), "pe"))
    for t in ("video_transformer", "audio_transformer"):
        s += _ln(f"{t}.layers.0.0") + _mha(f"{t}.layers.0.1", "xavier") + _ln(f"{t}.layers.0.2")
        s += _lin(f"{t}.layers.0.3.0", 1024, 256) + _lin(f"{t}.layers.0.3.3", 256, 1024)
        s += _lin(f"{t}.final_linear", 256, 256)
    x = "video_guided_to_music_pooling_cross_transformer"
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        s += _lin(f"{x}.cross_attn.{n}", 256, 256, "eye")
    s += _lin(f"{x}.linear_proj", 256, 256, "eye")
    s += _ln(f"{x}.layer_norm1") + _ln(f"{x}.layer_norm2") + _ln(f"{x}.layer_norm3")
    s.append(("decoder_query_embed.weight", (1, 256), "normal"))
    s += _lin("span_embed.layers.0", 256, 256) + _lin("span_embed.layers.1", 256, 256)
    s += _lin("span_embed.layers.2", 2, 256)
    s += _lin("class_embed", 2, 256)
    s += _lin("contrastive_align_projection_query", 256, 256)
    s += _lin("contrastive_align_projection_vid", 256, 256)
    s.append(("criterion.empty_weight", (2,), "empty_weight"))
    s += _lin("vit_proj", 256, 512) + _lin("ast_proj", 256, 768)
    return s


def xa_video_spec() -> List[Tuple[str, tuple, str]]:
    """Extra keys of vmr_fusion "XA-music-video": a second Transformer_XA (model_Uni.py:27-28)."""
    x = "music_guided_to_video_pooling_cross_transformer"
    s: List[Tuple[str, tuple, str]] = []
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        s += _lin(f"{x}.cross_attn.{n}", 256, 256, "eye")
    s += _lin(f"{x}.linear_proj", 256, 256, "eye")
    s += _ln(f"{x}.layer_norm1") + _ln(f"{x}.layer_norm2") + _ln(f"{x}.layer_norm3")
    return s


def ca_spec() -> List[Tuple[str, tuple, str]]:
    """Extra keys of mml_fusion "CA": video_music_fusion_cross_transformer = CrossTransformer(256, depth 1, 8 heads x
    128, mlp 1024, out 256) (model_Uni.py:33-43; model/model_Base.py:169-197, :100-113, :22-31)."""
    x = "video_music_fusion_cross_transformer"
    s: List[Tuple[str, tuple, str]] = [
        (f"{x}.layers.0.0.to_q.weight", (1024, 256), "xavier"),
        (f"{x}.layers.0.0.to_kv.weight", (2048, 256), "xavier"),
    ]
    s += _lin(f"{x}.layers.0.0.to_out.0", 256, 1024, "xavier")
    s += _lin(f"{x}.layers.0.1.net.0", 1024, 256, "xavier") + _lin(f"{x}.layers.0.1.net.3", 256, 1024, "xavier")
    s += _ln(f"{x}.attention_query_layer_norms.0") + _ln(f"{x}.attention_context_layer_norms.0")
    s += _ln(f"{x}.ff_layer_norms.0")
    s += _lin(f"{x}.final_linear", 256, 256)
    return s


def sinusoid_pe(seq_len: int, dim: int = C.D_MODEL) -> torch.Tensor:
    """Buffer of model_Base.py:48-57 (sin on even, cos on odd channels), same op order, fp32."""
    pe = torch.zeros(seq_len, dim)
    position = torch.arange(0, seq_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2).float() * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def make_state_dict(seed: int = 0, xa_video: bool = False, ca: bool = False) -> Dict[str, torch.Tensor]:
    """Random-but-deterministic fp32 weights with reference key names.

    X-Pool linears are eye + 0.05*N(0,1) (the reference eye-initialises them,
    modules/transformer.py:148-154, which would hide transpose bugs), LayerNorm affine params
    are N(1,0.1)/N(0,0.1) and every bias is non-zero, per SURVEY.md §8d.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec() + (xa_video_spec() if xa_video else []) + (ca_spec() if ca else []):
        if kind == "logit_scale":
            a = np.array(C.logit_scale_init(), dtype=np.float32)
        elif kind == "pe":
            sd[key] = sinusoid_pe(shape[1])
            continue
        elif kind == "empty_weight":
            a = np.array([1.0, 0.1], dtype=np.float32)  # loss_detr.py:53-56, fb_label "01"
        elif kind == "xavier":
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "linear":
            bound = 1.0 / math.sqrt(shape[1])
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "eye":
            a = np.eye(shape[0], shape[1]) + 0.05 * rng.standard_normal(size=shape)
        elif kind == "bias":
            a = 0.02 * rng.standard_normal(size=shape)
        elif kind == "ln_w":
            a = 1.0 + 0.1 * rng.standard_normal(size=shape)
        elif kind == "ln_b":
            a = 0.1 * rng.standard_normal(size=shape)
        elif kind == "normal":
            a = rng.standard_normal(size=shape)
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32).copy())
    return sd


def round_state_dict_bf16(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """bf16-round every >=2-D weight matrix (GEMM operands); biases/LN/pe stay fp32.

    This is the weight set the bf16 parity oracle uses: reference math in fp32 on bf16-rounded
    GEMM weights and inputs, so that only internal activation rounding differs.
    """
    out = {}
    for k, v in sd.items():
        if v.dim() >= 2 and not k.endswith(".pe"):
            out[k] = v.to(torch.bfloat16).to(torch.float32)
        else:
            out[k] = v.clone()
    return out


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------
def _lengths(rng: np.random.Generator, n: int):
    v_dur = np.clip(np.exp(rng.normal(math.log(20.2), 0.5, size=n)), 5.2, 49.4)
    short = rng.random(n) < 0.35
    m_dur = np.where(short, rng.uniform(30.1, 70.0, size=n), rng.uniform(120.0, 239.9, size=n))
    width = np.minimum(v_dur, m_dur - 0.5)
    frac = rng.beta(0.5, 1.5, size=n)
    m_start = frac * (m_dur - width)
    m_end = m_start + width
    n_frames = np.minimum(np.floor(v_dur).astype(np.int64) + 1, C.L_V)
    n_segments = np.minimum(np.floor(m_dur / C.STRIDE).astype(np.int64) + 1, C.L_M)
    return v_dur, m_dur, m_start, m_end, n_frames, n_segments


def make_videos(n: int, seed: int, dtype=torch.float32):
    """frame_feats [n,50,512] ~ N(0,1) with prefix masks; padded rows zeroed
    (dataloader_MGSV_EC_feature.py:57-61)."""
    rng = np.random.Generator(np.random.PCG64([seed, 1]))
    v_dur, _, _, _, n_frames, _ = _lengths(rng, n)
    feats = rng.standard_normal(size=(n, C.L_V, C.D_VIT), dtype=np.float32)
    mask = (np.arange(C.L_V)[None, :] < n_frames[:, None]).astype(np.float32)
    feats *= mask[:, :, None]
    return dict(
        frame_feats=torch.from_numpy(feats).to(dtype),
        frame_mask=torch.from_numpy(mask),
        v_duration=torch.from_numpy(v_dur.astype(np.float32)),
        n_frames=torch.from_numpy(n_frames),
    )


def make_tracks(n: int, seed: int, dtype=torch.float32):
    """segment_feats [n,96,768] ~ N(0,1) with prefix masks, plus a ground-truth moment per track
    (dataloader_MGSV_EC_feature.py:18-27,63-67)."""
    rng = np.random.Generator(np.random.PCG64([seed, 2]))
    _, m_dur, m_start, m_end, _, n_segments = _lengths(rng, n)
    feats = rng.standard_normal(size=(n, C.L_M, C.D_AST), dtype=np.float32)
    mask = (np.arange(C.L_M)[None, :] < n_segments[:, None]).astype(np.float32)
    feats *= mask[:, :, None]
    gt = np.stack([m_start, m_end], axis=-1).astype(np.float32)[:, None, :]  # [n,1,2] seconds
    gt_t = torch.from_numpy(gt)
    e = torch.clamp(gt_t[..., 1], max=C.MAX_M_DURATION)
    spans_target = torch.stack([(gt_t[..., 0] + e) / 2.0 / C.MAX_M_DURATION,
                                (e - gt_t[..., 0]) / C.MAX_M_DURATION], dim=-1)  # [n,1,2] (c,w)
    return dict(
        segment_feats=torch.from_numpy(feats).to(dtype),
        segment_mask=torch.from_numpy(mask),
        m_duration=torch.from_numpy(m_dur.astype(np.float32)),
        gt_moment=gt_t,
        spans_target=spans_target,
        n_segments=torch.from_numpy(n_segments),
    )


def make_eval_set(n_queries: int, n_tracks: int, seed: int, dtype=torch.float32):
    """Query i is paired with track i (i < n_queries); tracks >= n_queries are distractors
    (cfg 1/2)."""
    assert n_tracks >= n_queries
    v = make_videos(n_queries, seed, dtype)
    m = make_tracks(n_tracks, seed, dtype)
    ids = dict(
        video_ids=[f"v{i:07d}" for i in range(n_queries)],
        music_ids=[f"m{i:07d}" for i in range(n_tracks)],
    )
    return v, m, ids


def make_span_pairs(n: int, m: int, seed: int):
    """cfg 3: (c,w) spans, c~U(0,1), w~U(0.01,0.31), logits~N(0,1), with edge
    rows: zero-width prediction, identical spans, disjoint spans, zero-width target."""
    rng = np.random.Generator(np.random.PCG64([seed, 3]))
    a = np.stack([rng.uniform(0, 1, n), rng.uniform(0.01, 0.31, n)], -1).astype(np.float32)
    b = np.stack([rng.uniform(0, 1, m), rng.uniform(0.01, 0.31, m)], -1).astype(np.float32)
    logits = rng.standard_normal(size=(n, 2)).astype(np.float32)
    if n >= 4 and m >= 4:
        a[0] = (0.5, 0.0)          # zero-width prediction
        b[0] = (0.5, 0.0)          # zero-width target (matcher drops it, matcher.py:59)
        a[1] = b[1]                # identical spans
        a[2] = (0.1, 0.05)
        b[2] = (0.9, 0.05)         # disjoint
        a[3] = (0.5, 1.0)          # full-range span
    return torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(logits)
