"""GPU diagnostics: run one stage of the CUDA path against the CPU oracle and print error stats.

    python tests/tools/gpu_diag.py <stage> [...]     stages: span rank gemm attn encode xpool detr

Each stage is meant to be run under its own `timeout` (see scripts/gpu_diag.sh) so that a hang
or a trap in one kernel does not take the others down.  Test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from mgsv_b200 import _lib, ops, synth  # noqa: E402
from mgsv_b200.engine import Engine  # noqa: E402
from oracle import made_oracle as O  # noqa: E402

DEV = torch.device("cuda:0")


def stats(name, got, ref, extra=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    d = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    print(f"  {name:28s} max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e} "
          f"max|ref|={denom:.3e} rel={d.max().item() / denom:.3e} nan={int(torch.isnan(got).sum())} {extra}",
          flush=True)


def stage_span():
    a, b, logits = synth.make_span_pairs(1000, 1003, 7)
    a, b = a.to(DEV), b.to(DEV)
    g = ops.generalized_temporal_iou(ops.span_cw_to_se(a), ops.span_cw_to_se(b))
    ref = O.generalized_temporal_iou(O.span_cw_to_se(a.cpu()), O.span_cw_to_se(b.cpu()))
    print("  giou bit-exact:", bool(np.array_equal(g.cpu().numpy(), ref.numpy(), equal_nan=True)))
    a2, b2, logits = synth.make_span_pairs(1000, 1000, 8)
    prob = logits.softmax(-1)[:, 0].contiguous()
    tgt = b2[b2[:, 1] != 0]
    c = ops.matcher_cost(prob.to(DEV), a2.to(DEV), tgt.to(DEV))
    refc = O.matcher_cost(prob, a2, tgt)
    print("  cost bit-exact:", bool(np.array_equal(c.cpu().numpy(), refc.numpy(), equal_nan=True)), tuple(c.shape))
    iou, uni = ops.temporal_iou(ops.span_cw_to_se(a2.to(DEV)), ops.span_cw_to_se(b2.to(DEV)))
    ri, ru = O.temporal_iou(O.span_cw_to_se(a2), O.span_cw_to_se(b2))
    print("  iou/union bit-exact:", bool(np.array_equal(iou.cpu().numpy(), ri.numpy(), equal_nan=True)),
          bool(np.array_equal(uni.cpu().numpy(), ru.numpy(), equal_nan=True)))


def stage_rank():
    rng = np.random.default_rng(0)
    n = 300
    single = torch.from_numpy(rng.standard_normal((n, n)).astype(np.float32))
    dual = torch.from_numpy(rng.standard_normal((n, n)).astype(np.float32))
    ids = [f"m{i}" for i in range(n)]
    for j in range(40):
        ids[n - 1 - j] = ids[j]
        single[:, n - 1 - j] = single[:, j] + (0.0 if j % 2 else 0.01)
        dual[:, n - 1 - j] = dual[:, j]
    total = single.numpy() * 1.0 + dual.numpy().astype(np.float64)
    m, ind, top1 = O.recall_metrics(total, ids)
    prev, gt_col, _ = ops.dedup_tables(ids)
    r = ops.rank_topk(single.to(DEV), dual.to(DEV), torch.from_numpy(gt_col).to(DEV), torch.from_numpy(prev).to(DEV), k=100)
    print("  rank equal:", bool(np.array_equal(r["rank"].cpu().numpy(), ind)), "mismatch", int((r["rank"].cpu().numpy() != ind).sum()))
    tv, ti = torch.topk(torch.from_numpy(total), 100, dim=1)
    print("  topk score equal:", bool(torch.equal(r["topk_score"].cpu(), tv)),
          "idx mismatch:", int((r["topk_idx"].cpu().long() != ti).sum()))
    x = torch.randn(70, 256)
    y = torch.randn(130, 256)
    stats("cosine", ops.cal_distance(x.to(DEV), y.to(DEV)), O.cal_distance_cos(x, y))
    cs = torch.randn(50, 300, dtype=torch.float64)
    ci = torch.arange(300, dtype=torch.int32).repeat(50, 1)
    oi, os_ = ops.topk_merge(cs.to(DEV), ci.to(DEV), 100)
    tv, ti = torch.topk(cs, 100, dim=1)
    print("  merge equal:", bool(torch.equal(os_.cpu(), tv)), bool(torch.equal(oi.cpu().long(), ti)))


def stage_gemm():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (300, 256, 512), (1000, 768, 256), (5000, 1024, 256),
                      (4800, 256, 1024), (40000, 256, 768)]:
        a = torch.randn(M, K).to(torch.float16)
        w = (torch.randn(N, K) / K ** 0.5).to(torch.float16)
        bias = torch.randn(N)
        ref = a.float() @ w.float().t() + bias
        t0 = time.time()
        out = ops.gemm_f16(a.to(DEV), w.to(DEV), bias=bias.to(DEV), out_dtype=torch.float32)
        torch.cuda.synchronize()
        stats(f"gemm {M}x{N}x{K} f32out", out, ref, f"{time.time() - t0:.3f}s")
    M, N, K = 1000, 256, 256
    a = torch.randn(M, K).to(torch.float16)
    w = (torch.randn(N, K) / K ** 0.5).to(torch.float16)
    bias, res = torch.randn(N), torch.randn(M, N)
    g, b = torch.randn(N), torch.randn(N)
    base = a.float() @ w.float().t() + bias
    out = ops.gemm_f16(a.to(DEV), w.to(DEV), bias=bias.to(DEV), residual=res.to(DEV), ln=(g.to(DEV), b.to(DEV)),
                        out_dtype=torch.float32)
    stats("gemm +res +LN", out, torch.nn.functional.layer_norm(base + res, (N,), g, b))
    out = ops.gemm_f16(a.to(DEV), w.to(DEV), bias=bias.to(DEV), act=1, out_dtype=torch.float16)
    stats("gemm gelu bf16out", out.float(), torch.nn.functional.gelu(base))
    out = ops.gemm_f16(a.to(DEV), w.to(DEV), bias=bias.to(DEV), act=2, out_dtype=torch.float32)
    stats("gemm relu", out, torch.relu(base))


def stage_attn():
    torch.manual_seed(1)
    for L in (50, 96, 146):
        B = 7
        q, k, v = [torch.randn(B, L, 256).to(torch.float16) for _ in range(3)]
        n_valid = torch.randint(1, L + 1, (B,))
        n_valid[0] = L
        mask = (torch.arange(L)[None] < n_valid[:, None]).float()
        out = ops.mha_core(q.to(DEV), k.to(DEV), v.to(DEV), mask.to(DEV))
        qh = q.float().view(B, L, 8, 32).transpose(1, 2)
        kh = k.float().view(B, L, 8, 32).transpose(1, 2)
        vh = v.float().view(B, L, 8, 32).transpose(1, 2)
        s = (qh @ kh.transpose(-1, -2)) / 32 ** 0.5
        s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
        ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, L, 256)
        stats(f"mha L={L}", out.float(), ref)


def _engine(sd):
    eng = Engine(DEV)
    eng.load_state_dict(sd)
    return eng


def stage_encode():
    sd = synth.make_state_dict(0)
    sdb = synth.round_state_dict_bf16(sd)
    eng = _engine(sd)
    v, m, ids = synth.make_eval_set(48, 48, synth.BASE_SEED + 100)
    for name, mod, feats, mask, fn in (("video", _lib.VIDEO, v["frame_feats"], v["frame_mask"], O.encode_video),
                                       ("music", _lib.MUSIC, m["segment_feats"], m["segment_mask"], O.encode_music)):
        seq, seq32, pooled = eng.encode(mod, feats.to(DEV), mask.to(DEV))
        torch.cuda.synchronize()
        rs, rp = fn(sd, feats, mask)
        rsb, rpb = fn(sdb, feats.to(torch.float16).float(), mask)
        stats(f"{name} seq vs fp32 oracle", seq32, rs)
        stats(f"{name} seq vs bf16w oracle", seq32, rsb)
        stats(f"{name} pooled vs fp32", pooled, rp)
        stats(f"{name} pooled vs bf16w", pooled, rpb)
        stats(f"{name} seq bf16 out", seq.float(), rs)


def stage_xpool():
    sd = synth.make_state_dict(0)
    eng = _engine(sd)
    nq, nm = 200, 150
    v, m, ids = synth.make_eval_set(nq, nq, synth.BASE_SEED + 100)
    fo, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
    so, mf = O.encode_music(sd, m["segment_feats"][:nm], m["segment_mask"][:nm])
    single, dual, total = O.gallery_similarity(sd, vf, mf, so, m["segment_mask"][:nm])
    seg_bf16 = so.to(torch.float16).to(DEV)
    kz, gram, bits = eng.gallery_prepare(seg_bf16, m["segment_mask"][:nm].to(DEV))
    q, vhat = eng.query_prepare(vf.to(DEV))
    torch.cuda.synchronize()
    print("  prepare ok", flush=True)
    sim = eng.xpool_score(q, vhat, kz, gram, bits)
    torch.cuda.synchronize()
    stats("xpool single sim", sim, single)
    stats("dual sim", ops.cal_distance(vf.to(DEV), mf.to(DEV)), dual)
    # ranking agreement
    tot_gpu = sim.double().cpu().numpy() + dual.double().numpy()
    agree = (np.argmax(tot_gpu, 1) == np.argmax(total, 1)).mean()
    print(f"  top-1 agreement {agree:.4f}; sim range [{single.min():.4f}, {single.max():.4f}]")


def stage_detr():
    sd = synth.make_state_dict(0)
    eng = _engine(sd)
    B = 40
    v, m, ids = synth.make_eval_set(B, B, synth.BASE_SEED + 100)
    fo, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
    so, mf = O.encode_music(sd, m["segment_feats"], m["segment_mask"])
    src = torch.cat([fo, so], 1)
    mask = torch.cat([v["frame_mask"], m["segment_mask"]], 1)
    hs, memory = O.detr_forward(sd, src, mask, O.position_embedding_sine(mask), vf.unsqueeze(1))
    om = O.calc_output(sd, hs, fo)
    r = eng.detr_detect(fo.to(torch.float16).to(DEV), v["frame_mask"].to(DEV), so.to(torch.float16).to(DEV),
                        m["segment_mask"].to(DEV), vf.to(DEV), want_proj=True, want_memory=True)
    torch.cuda.synchronize()
    stats("memory", r["memory"], memory)
    stats("hs", r["hs"], hs[:, :, 0])
    stats("pred_logits(last)", r["pred_logits"][-1], om["pred_logits"][:, 0])
    stats("pred_spans(last)", r["pred_spans"][-1], om["pred_spans"][:, 0])
    stats("proj_queries(last)", r["proj_queries"][-1], om["proj_queries"][:, 0])
    stats("proj_vid_mem", r["proj_vid_mem"], om["proj_vid_mem"])
    st, ed, sc, iou = ops.moment_postproc(r["pred_logits"][-1], r["pred_spans"][-1], m["gt_moment"].to(DEV),
                                          m["m_duration"].to(DEV))
    rst, red, rsc = O.moment_postproc(om["pred_logits"], om["pred_spans"])
    stats("pred_st (s)", st, rst)
    stats("iou", iou, O.detr_iou(rst, red, m["gt_moment"], m["m_duration"]))


STAGES = dict(span=stage_span, rank=stage_rank, gemm=stage_gemm, attn=stage_attn, encode=stage_encode,
              xpool=stage_xpool, detr=stage_detr)

if __name__ == "__main__":
    names = sys.argv[1:] or list(STAGES)
    print(torch.cuda.get_device_name(0), "lib", _lib.lib_path(), flush=True)
    for nme in names:
        print(f"[{nme}]", flush=True)
        t0 = time.time()
        STAGES[nme]()
        torch.cuda.synchronize()
        print(f"[{nme}] done in {time.time() - t0:.1f}s", flush=True)
