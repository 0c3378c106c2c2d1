"""Ablation timings of the fused X-Pool kernel (MADE_XPOOL_DEBUG bits) at the bench's chunk shape. Diagnostics."""
# needs a diagnostics build: MADE_DIAG=1 python -m mgsv_b200.build --force  (rebuild without MADE_DIAG afterwards)
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import _lib, synth
from mgsv_b200.engine import Engine
dev = torch.device("cuda:0")
eng = Engine(dev); eng.load_state_dict(synth.make_state_dict(0))
nq, nm = 2000, 1000
v, m, _ = synth.make_eval_set(nq, nq, synth.BASE_SEED + 2)
m = {k: t[:nm] for k, t in m.items() if isinstance(t, torch.Tensor)}
seq_m, _, _ = eng.encode(_lib.MUSIC, m["segment_feats"].to(dev), m["segment_mask"].to(dev), want_f32=False)
_, _, pooled_v = eng.encode(_lib.VIDEO, v["frame_feats"].to(dev), v["frame_mask"].to(dev), want_f32=False)
q, vhat = eng.query_prepare(pooled_v)
full = torch.ones_like(m["segment_mask"])
for name, mask in (("MGSV-shaped lengths (mean %.1f)" % m["segment_mask"].sum(1).mean().item(), m["segment_mask"]), ("all 96 valid", full)):
    kz, gram, bits = eng.gallery_prepare(seq_m, mask.to(dev))
    for dbg in [int(a) for a in (sys.argv[1:] or ["0", "1", "2", "4", "8", "9", "16", "32", "63"])]:
        os.environ["MADE_XPOOL_DEBUG"] = str(dbg)
        f = lambda: eng.xpool_score(q, vhat, kz, gram, bits)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"{name}: debug={dbg:2d} (1 no sweep, 2 no exp, 4 no qf, 8 no Y mma, 16 no S mma, 32 no T mma): {us:7.1f} us", flush=True)
os.environ["MADE_XPOOL_DEBUG"] = "0"
