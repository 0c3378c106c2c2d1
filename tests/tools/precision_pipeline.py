"""CPU emulation of the WHOLE similarity pipeline of the CUDA path (temporal encoders -> X-Pool
operands -> fused scoring, and the dual-tower cosine) with a rounding switch at every point where
the CUDA path stores or consumes a 16-bit operand.  It attributes the end-to-end similarity error
of configs[1] to its sources and evaluates candidate fixes (fp16 hi+lo operands at chosen points)
before any GPU time is spent.  Test infrastructure only.

    python tests/tools/precision_pipeline.py [n_q] [n_m]
"""
from __future__ import annotations

import math
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from mgsv_b200 import synth  # noqa: E402
from oracle import made_oracle as O  # noqa: E402

X = "video_guided_to_music_pooling_cross_transformer"
D = 256
dd = torch.float64


def rnd(t, kind):
    """kind: None = exact, 'h' = fp16, 'hl' = fp16 hi + fp16 lo (~22 bits), 'b' = bf16."""
    if kind is None:
        return t
    if kind == "h":
        return t.to(torch.float16).to(dd)
    if kind == "b":
        return t.to(torch.bfloat16).to(dd)
    if kind == "hl":
        hi = t.to(torch.float16).to(dd)
        lo = (t - hi).to(torch.float16).to(dd)
        return hi + lo
    if kind == "f":
        return t.to(torch.float32).to(dd)
    raise ValueError(kind)


def encoder(sd, feats, masks, proj, tr, pe_key, r, tag):
    """Emulates made_encode (api.cu encode_packed).  r maps rounding-point names to kinds; names are
    prefixed with tag ('v' / 'm') and fall back to the un-prefixed name."""
    def k(name):
        return r.get(f"{tag}.{name}", r.get(name))
    g = lambda key: sd[key].to(dd)
    L = feats.shape[1]
    valid = masks.bool()
    x0 = rnd(feats.to(dd).masked_fill(~valid.unsqueeze(-1), 0), k("in"))
    lin = lambda x, p, wk: x @ rnd(g(p + ".weight"), k(wk)).T + g(p + ".bias")
    ln = lambda x, p: F.layer_norm(x, (D,), g(p + ".weight"), g(p + ".bias"), 1e-5)
    x1f = ln(lin(x0, proj, "w_proj") + g(pe_key)[:, :L], f"{tr}.layers.0.0")
    x1 = rnd(x1f, k("x1"))
    w_in = rnd(g(f"{tr}.layers.0.1.in_proj_weight"), k("w_in"))
    qkv = rnd(x1 @ w_in.T + g(f"{tr}.layers.0.1.in_proj_bias"), k("qkv"))
    B = feats.shape[0]
    q, kk, v = (qkv[..., i * D:(i + 1) * D].view(B, L, 8, 32).transpose(1, 2) for i in range(3))
    s = (q @ kk.transpose(-1, -2)) / math.sqrt(32)
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    e = rnd(torch.exp(s - s.max(-1, keepdim=True).values), k("p"))
    att = (e @ v) / e.sum(-1, keepdim=True)
    att = rnd(att.transpose(1, 2).reshape(B, L, D), k("att"))
    x2f = ln(lin(att, f"{tr}.layers.0.1.out_proj", "w_out") + x1f, f"{tr}.layers.0.2")
    x2 = rnd(x2f, k("x2"))
    h = rnd(F.gelu(lin(x2, f"{tr}.layers.0.3.0", "w_ff1")), k("h"))
    x3 = rnd(lin(h, f"{tr}.layers.0.3.3", "w_ff2") + x2f, k("x3"))
    seqf = lin(x3, f"{tr}.final_linear", "w_fin").masked_fill(~valid.unsqueeze(-1), 0)
    pooled = seqf.sum(1) / masks.to(dd).sum(1, keepdim=True)
    pooled = pooled / pooled.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return rnd(seqf, k("seq")), rnd(pooled, "f")


def xpool_folded(sd, vf, seg, mask, r):
    """Emulates made_gallery_prepare / made_query_prepare / xpool_score_kernel (folded algebra)."""
    k = lambda name: r.get(f"x.{name}")
    g = lambda key: sd[f"{X}.{key}"].to(dd)
    ln = lambda x, n: F.layer_norm(x, (D,), g(f"layer_norm{n}.weight"), g(f"layer_norm{n}.bias"), 1e-5)
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
    sp = rnd(ln(seg, 1), k("sp"))
    K = rnd(sp @ rnd(Wk, k("w_k")).T + bk, k("k"))
    V = rnd(sp @ rnd(Wvo, k("w_v")).T + bvo, k("v"))
    Z = rnd(sp @ rnd(Wz, k("w_z")).T + bz, k("z"))
    G = rnd(V @ V.transpose(-1, -2), k("g"))
    vp = rnd(ln(vf, 1), k("vp"))
    q = rnd(vp @ rnd(Wq, k("w_q")).T + bq, k("q"))
    vhat = vf / vf.norm(dim=-1, keepdim=True)
    u = rnd(rnd(vhat, k("vhat")) * g3, k("u"))
    S = torch.einsum("vd,mtd->mvt", q, K)
    S = S.masked_fill(mask[:, None, :] == 0, float("-inf"))
    e = rnd(torch.exp(S - S.max(-1, keepdim=True).values), k("p"))
    l = e.sum(-1)
    T = torch.einsum("mvt,mts->mvs", e, G)
    qf = (e * T).sum(-1)
    Y = torch.einsum("mvt,mtd->mvd", e, Z)
    var2 = qf / l / l / D
    alpha = torch.rsqrt(var2.clamp_min(0) + 1e-5) / l
    o = alpha[..., None] * Y + bprime
    mean = o.mean(-1, keepdim=True)
    var3 = o.var(-1, unbiased=False, keepdim=True)
    rs = torch.rsqrt(var3 + 1e-5)
    # dot = rs * sum(u * (o - mean)) + sum(vhat * beta3)
    dot = rs[..., 0] * ((o - mean) * u[None]).sum(-1) + (rnd(vhat, k("vhat")) * b3).sum(-1)[None]
    p = (o - mean) * rs * g3 + b3
    return (dot / p.norm(dim=-1)).T


def main():
    torch.manual_seed(0)
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nm = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    sd = synth.make_state_dict(0)
    sd64 = {k_: t.double() for k_, t in sd.items()}
    v, m, ids = synth.make_eval_set(nq, nm, synth.BASE_SEED + 2)
    with torch.no_grad():
        fo, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
        so, mf = O.encode_music(sd, m["segment_feats"], m["segment_mask"])
        smask = m["segment_mask"]
        single_ref, dual_ref, _ = O.gallery_similarity(sd, vf, mf, so * smask.unsqueeze(-1), smask)
    single_ref, dual_ref = single_ref.double(), dual_ref.double()
    print(f"scale single {single_ref.abs().max():.4f} dual {dual_ref.abs().max():.4f}")

    def run(r):
        with torch.no_grad():
            seq_v, pv = encoder(sd, v["frame_feats"], v["frame_mask"], "vit_proj", "video_transformer",
                                "video_position_embedding.pe", r, "v")
            seq_m, pm = encoder(sd, m["segment_feats"], m["segment_mask"], "ast_proj", "audio_transformer",
                                "audio_position_embedding.pe", r, "m")
            single = xpool_folded(sd, pv, seq_m, smask, r)
            dual = pv @ pm.T
        return single, dual

    def rep(name, r):
        s, d = run(r)
        out = []
        for nm_, got, ref in (("single", s, single_ref), ("dual", d, dual_ref)):
            e = (got - ref).abs()
            sc = ref.abs().max()
            out.append(f"{nm_}: max {e.max() / sc:.2e} rms {e.pow(2).mean().sqrt() / sc:.2e}")
        print(f"{name:64s} " + " | ".join(out), flush=True)

    quick = os.environ.get("QUICK")
    enc_pts = ["in", "w_proj", "x1", "w_in", "qkv", "p", "att", "w_out", "x2", "w_ff1", "h", "w_ff2", "x3", "w_fin", "seq"]
    xp_pts = ["x.sp", "x.w_k", "x.w_v", "x.w_z", "x.k", "x.v", "x.z", "x.g", "x.vp", "x.w_q", "x.q", "x.vhat", "x.u", "x.p"]
    cur = {p: "h" for p in enc_pts + xp_pts}
    rep("exact (no rounding)", {})
    rep("current CUDA path (fp16 everywhere)", cur)
    if not quick:
        print("--- only ONE point rounded to fp16 ---")
        for p in enc_pts:
            rep(f"  only encoder {p}", {p: "h"})
        for p in xp_pts:
            rep(f"  only {p}", {p: "h"})
        print("--- current, with ONE point upgraded to hi+lo ---")
        for p in enc_pts + xp_pts:
            r = dict(cur)
            r[p] = "hl"
            rep(f"  current but {p} hi+lo", r)
    print("--- candidate fixes ---")
    allw = [p for p in enc_pts + xp_pts if "w_" in p]
    r = dict(cur)
    for p in allw:
        r[p] = "hl"
    rep("all weights hi+lo", r)
    r2 = dict(r)
    for p in ("in", "x1", "x2", "x3", "att", "seq", "x.sp", "x.vp"):
        r2[p] = "hl"
    rep("all weights + token activations (A operands of the K=256 GEMMs) hi+lo", r2)
    r3 = dict(r2)
    for p in ("h", "qkv"):
        r3[p] = "hl"
    rep("  + h, qkv hi+lo", r3)
    ra = dict(r2); ra["x.vhat"] = None
    rep("weights + token activations hi+lo, vhat fp32", ra)
    rb = dict(ra); rb["x.u"] = "hl"
    rep("  + u hi+lo", rb)
    rc = dict(r); rc["x.vhat"] = None
    rep("weights hi+lo only, vhat fp32", rc)
    rd = dict(ra); rd["seq"] = "h"; rd["x.sp"] = "h"
    rep("weights + enc-internal activations hi+lo (seq, x.sp fp16), vhat fp32", rd)
    re_ = dict(ra); re_["m.x1"] = re_["m.x2"] = re_["m.x3"] = re_["m.att"] = re_["m.in"] = "h"
    rep("weights hi+lo; video tokens hi+lo, music enc tokens fp16; seq,x.sp,x.vp hi+lo; vhat fp32", re_)
    c1 = dict(cur)
    for p in ("w_fin", "x3", "w_out", "w_in", "w_proj", "in", "x.w_z", "x.w_k", "x.w_v", "x.sp", "x.w_q", "x.vp"):
        c1[p] = "hl"
    c1["x.vhat"] = None
    rep("C1: proj/in/out/final + xpool projections split, FF fp16, vhat fp32", c1)
    c2 = dict(c1); c2["x1"] = "hl"
    rep("C2: C1 + x1 hi+lo", c2)
    c3 = dict(c2); c3["seq"] = "hl"
    rep("C3: C2 + seq hi+lo", c3)
    c4 = dict(c1); c4["in"] = "h"; c4["w_proj"] = "h"
    rep("C4: C1 without the input projection split", c4)
    c5 = dict(c3); c5["x.u"] = "hl"
    rep("C5: C3 + u hi+lo", c5)
    r4 = dict(r3)
    for p in ("x.k", "x.z", "x.q"):
        r4[p] = "hl"
    rep("  + x.k x.z x.q hi+lo", r4)


if __name__ == "__main__":
    main()
