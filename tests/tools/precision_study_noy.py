"""CPU study for a Y-free X-Pool epilogue: the cosine of LN3(alpha*Y + b') with v^ needs
only SUMS over the 256 features of Y = e.Z''.  Five are linear with constant weights (the W5 columns of
the current kernel); the other three can be produced without ever forming Y:

    sum Y^2          = e^T (Z'' Z''^T) e                 Gram matrix GZ  [96 x 96] per track
    sum g3^2 Y^2     = e^T (Z'' diag(g3^2) Z''^T) e      Gram matrix GZ3 [96 x 96] per track
    sum u Y          = sum_t e_t (Z''_t . u_q)           S2 = u Z''^T, a [128 x 96] MMA like S = q K^T

so the 128x256x96 Y MMA and the 256-column Y sweep of the epilogue become a 128x304x96 MMA against
[G | GZ | GZ3 | W5] plus a 128x96x256 MMA, and three 96-term dot products per row.  This script checks
the algebra in fp64 and measures what rounding the new operands to fp16 costs, next to the current
formulation ("Y in fp32 from fp16 operands").  Test infrastructure only.

    python tests/tools/precision_study_noy.py
"""
from __future__ import annotations

import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mgsv_b200 import synth  # noqa: E402
from oracle import made_oracle as O  # noqa: E402
from precision_study import D, X, rnd  # noqa: E402


def hilo(t):
    hi = t.to(torch.float16).to(t.dtype)
    return hi + (t - hi).to(torch.float16).to(t.dtype)


def sims(sd, vf, seg, mask, mode, r16=True, gram="fp16"):
    dd = torch.float64
    g = lambda k: sd[f"{X}.{k}"].to(dd)
    ln = lambda x, n: torch.nn.functional.layer_norm(x, (D,), g(f"layer_norm{n}.weight"), g(f"layer_norm{n}.bias"), 1e-5)
    R = (lambda t: rnd(t, "fp16")) if r16 else (lambda t: t)
    RG = {"fp16": lambda t: rnd(t, "fp16"), "hilo": hilo, "exact": lambda t: t}[gram]
    Wq, bq = g("cross_attn.q_proj.weight") / 16, g("cross_attn.q_proj.bias") / 16
    Wk, bk = g("cross_attn.k_proj.weight"), g("cross_attn.k_proj.bias")
    Wv, bv = g("cross_attn.v_proj.weight"), g("cross_attn.v_proj.bias")
    Wo, bo = g("cross_attn.out_proj.weight"), g("cross_attn.out_proj.bias")
    Wl, bl = g("linear_proj.weight"), g("linear_proj.bias")
    g2, b2 = g("layer_norm2.weight"), g("layer_norm2.bias")
    g3, b3 = g("layer_norm3.weight"), g("layer_norm3.bias")
    Wvo = Wo @ Wv
    bvo = Wo @ bv + bo
    Wvo = Wvo - Wvo.mean(0, keepdim=True)
    bvo = bvo - bvo.mean()
    Wp = (torch.eye(D, dtype=dd) + Wl) * g2[None, :]
    bp = b2 + bl + Wl @ b2
    sp = R(ln(R(seg.to(dd)), 1))
    K = R(sp @ R(Wk).T + bk)
    V = R(sp @ R(Wvo).T + bvo)
    Z = R(sp @ R(Wp @ Wvo).T + Wp @ bvo)
    G = R(V @ V.transpose(-1, -2))
    vp = R(ln(vf.to(dd), 1))
    q = R(vp @ R(Wq).T + bq)
    vhat = vf.to(dd) / vf.to(dd).norm(dim=-1, keepdim=True)
    u = R(vhat * g3)                                             # fp16 in TMEM today
    S = torch.einsum("vd,mtd->mvt", q, K).masked_fill(mask[:, None, :] == 0, float("-inf"))
    e = R(torch.exp(S - S.max(-1, keepdim=True).values))
    l = e.sum(-1)
    qf = (e * torch.einsum("mvt,mts->mvs", e, G)).sum(-1)
    alpha = torch.rsqrt((qf / l / l / D).clamp_min(0) + 1e-5) / l
    # the five linear sums (W5 columns, fp16 hi+lo: exact to fp32)
    c5 = torch.stack([torch.ones(D, dtype=dd), bp, g3 * g3, g3 * g3 * bp, g3 * b3], 1)     # [256,5]
    lin = torch.einsum("mvt,mtk->mvk", e, Z @ c5)                                          # [m,v,5]
    if mode == "y":                       # today: Y accumulated in fp32 from fp16 operands, swept by the epilogue
        Y = torch.einsum("mvt,mtd->mvd", e, Z)
        sy2 = (Y * Y).sum(-1)
        sg2y2 = (g3 * g3 * Y * Y).sum(-1)
        suy = torch.einsum("vd,mvd->mv", u, Y)
    else:                                 # Y-free: two more Gram quadratic forms and S2 = u Z''^T
        GZ = RG(Z @ Z.transpose(-1, -2))
        GZ3 = RG((Z * (g3 * g3)) @ Z.transpose(-1, -2))
        sy2 = (e * torch.einsum("mvt,mts->mvs", e, GZ)).sum(-1)
        sg2y2 = (e * torch.einsum("mvt,mts->mvs", e, GZ3)).sum(-1)
        S2 = torch.einsum("vd,mtd->mvt", u, Z)
        suy = (e * S2).sum(-1)
    s1, sb, sg2, sg2b, sgb = lin.unbind(-1)
    B1, B2 = bp.sum(), (bp * bp).sum()
    mu = (alpha * s1 + B1) / D
    var = (alpha * alpha * sy2 + 2 * alpha * sb + B2) / D - mu * mu
    inv = torch.rsqrt(var.clamp_min(0) + 1e-5)
    G2, G2b, G2b2 = (g3 * g3).sum(), (g3 * g3 * bp).sum(), (g3 * g3 * bp * bp).sum()
    Gb, Gbb, Bb = (g3 * b3).sum(), (g3 * b3 * bp).sum(), (b3 * b3).sum()
    su, sub, svb = u.sum(-1), (u * bp).sum(-1), (vhat * b3).sum(-1)                        # per query
    dot = inv * (alpha * suy + sub[None, :] - mu * su[None, :]) + svb[None, :]
    q2 = alpha * alpha * sg2y2 + 2 * alpha * sg2b + G2b2 - 2 * mu * (alpha * sg2 + G2b) + mu * mu * G2
    n2 = inv * inv * q2 + 2 * inv * (alpha * sgb + Gbb - mu * Gb) + Bb
    return (dot * torch.rsqrt(n2)).transpose(0, 1)


def main():
    sd = synth.make_state_dict(0)
    v, m, _ = synth.make_eval_set(64, 128, synth.BASE_SEED + 100)
    _, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
    so, _ = O.encode_music(sd, m["segment_feats"], m["segment_mask"])
    mask = m["segment_mask"]
    sd64 = {k: t.double() for k, t in sd.items()}
    ref = O.sim_matrix_music_pooling(vf.double(), O.xpool(sd64, vf.double(), so.double(), mask))
    scale = ref.abs().max()

    def rep(name, got):
        d = (got - ref).abs()
        print(f"{name:58s} max|d| {d.max():.3e} = {d.max() / scale:.2e} of scale, rms {d.pow(2).mean().sqrt() / scale:.2e}")

    rep("closed form, Y swept, no rounding (algebra check)", sims(sd, vf, so, mask, "y", r16=False))
    rep("closed form, Y-free, no rounding (algebra check)", sims(sd, vf, so, mask, "noy", r16=False, gram="exact"))
    rep("today: fp16 operands, Y swept", sims(sd, vf, so, mask, "y"))
    rep("Y-free, Grams GZ/GZ3 exact", sims(sd, vf, so, mask, "noy", gram="exact"))
    rep("Y-free, Grams GZ/GZ3 fp16", sims(sd, vf, so, mask, "noy", gram="fp16"))
    rep("Y-free, Grams GZ/GZ3 fp16 hi+lo", sims(sd, vf, so, mask, "noy", gram="hilo"))


if __name__ == "__main__":
    main()
