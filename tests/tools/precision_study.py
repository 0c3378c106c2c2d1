"""CPU emulation of the fused X-Pool kernel's algebra (csrc/xpool.cu, api.cu:load_xpool) with a
configurable rounding at every point where the CUDA path stores a 16-bit operand.  Used to decide
which operands need fp16 (10-bit mantissa) instead of bf16 (7-bit).  Test infrastructure only.

    python tests/tools/precision_study.py
"""
from __future__ import annotations

import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from mgsv_b200 import synth  # noqa: E402
from oracle import made_oracle as O  # noqa: E402

X = "video_guided_to_music_pooling_cross_transformer"
D = 256


def rnd(t, kind):
    if kind == "bf16":
        return t.to(torch.bfloat16).to(t.dtype)
    if kind == "fp16":
        return t.to(torch.float16).to(t.dtype)
    return t


def folded(sd, vf, seg, mask, r):
    """r: dict point -> 'bf16' | 'fp16' | None"""
    dd = torch.float64
    g = lambda k: sd[f"{X}.{k}"].to(dd)
    ln = lambda x, n: torch.nn.functional.layer_norm(x, (D,), g(f"layer_norm{n}.weight"), g(f"layer_norm{n}.bias"), 1e-5)
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
    bprime = b2 + bl + Wl @ b2
    Wz = Wp @ Wvo
    bz = Wp @ bvo
    # operands
    seg = rnd(seg.to(dd), r.get("seg"))
    sp = rnd(ln(seg, 1), r.get("sp"))
    K = rnd(sp @ rnd(Wk, r.get("w")).T + bk, r.get("k"))
    V = rnd(sp @ rnd(Wvo, r.get("w")).T + bvo, r.get("v"))
    if r.get("z_from_v"):
        Z = rnd(V @ rnd(Wp, r.get("w")).T, r.get("z"))
    else:
        Z = rnd(sp @ rnd(Wz, r.get("w")).T + bz, r.get("z"))
    G = rnd(V @ V.transpose(-1, -2), r.get("g"))                   # [N,96,96]
    vp = rnd(ln(vf.to(dd), 1), r.get("sp"))
    q = rnd(vp @ rnd(Wq, r.get("w")).T + bq, r.get("q"))
    vhat = rnd(vf.to(dd) / vf.to(dd).norm(dim=-1, keepdim=True), r.get("vhat"))
    S = torch.einsum("vd,mtd->mvt", q, K)                          # [N_m,N_v,96]
    S = S.masked_fill(mask[:, None, :] == 0, float("-inf"))
    e = rnd(torch.exp(S - S.max(-1, keepdim=True).values), r.get("p"))
    l = e.sum(-1)
    T = torch.einsum("mvt,mts->mvs", e, G)
    qf = (e * T).sum(-1)
    Y = torch.einsum("mvt,mtd->mvd", e, Z)
    var2 = qf / l / l / D
    alpha = torch.rsqrt(var2.clamp_min(0) + 1e-5) / l
    o = alpha[..., None] * Y + bprime
    p = torch.nn.functional.layer_norm(o, (D,), g3, b3, 1e-5)
    p = p / p.norm(dim=-1, keepdim=True)
    return torch.einsum("vd,mvd->vm", vhat, p)


def main():
    torch.manual_seed(0)
    sd = synth.make_state_dict(0)
    nq, nm = 64, 64
    v, m, ids = synth.make_eval_set(nq, nm, synth.BASE_SEED + 100)
    fo, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
    so, mf = O.encode_music(sd, m["segment_feats"], m["segment_mask"])
    mask = m["segment_mask"]
    sd64 = {k: t.double() for k, t in sd.items()}
    pooled = O.xpool(sd64, vf.double(), so.double(), mask)
    ref = O.sim_matrix_music_pooling(vf.double(), pooled)
    print(f"ref sim range [{ref.min():.4f}, {ref.max():.4f}]  mean|ref| {ref.abs().mean():.4f}")

    def rep(name, r):
        got = folded(sd, vf, so, mask, r)
        d = (got - ref).abs()
        rel = (d / (1e-3 * ref.abs() + 1e-5)).max()
        print(f"{name:58s} max|d| {d.max():.3e} mean|d| {d.mean():.3e}  max d/(1e-3|ref|+1e-5) {rel:.2f}")

    rep("exact fold (no rounding)", {})
    allb = dict(seg="bf16", sp="bf16", w="bf16", k="bf16", v="bf16", z="bf16", g="bf16", q="bf16", vhat="fp16", p="bf16")
    rep("current CUDA path (all bf16)", allb)
    for key in ("seg", "sp", "w", "k", "v", "z", "g", "q", "p"):
        rep(f"  only {key} bf16", {key: "bf16"})
    for key in ("seg", "sp", "w", "k", "v", "z", "g", "q", "p"):
        r = dict(allb)
        r[key] = None
        rep(f"  all bf16 except {key} exact", r)
    allh = {k: "fp16" for k in allb}
    rep("all fp16", allh)
    r = dict(allh); r["seg"] = "bf16"
    rep("all fp16, seg bf16 (encoder output bf16)", r)
    r = dict(allh); r["w"] = "bf16"
    rep("all fp16, weights bf16", r)
    r = dict(allb); r["p"] = "fp16"; r["g"] = "fp16"; r["z"] = "fp16"; r["v"] = "fp16"
    rep("bf16 but p,g,z,v fp16", r)
    r = dict(allb); r["p"] = "fp16"
    rep("bf16 but p fp16", r)


if __name__ == "__main__":
    main()
