"""Per-stage CUDA-event timings of the job's C-ABI calls in isolation (one stream, no overlap). Diagnostics."""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import _lib, ops, synth
from mgsv_b200.engine import Engine
from mgsv_b200.pipeline import GalleryEvaluator

dev = torch.device("cuda:0")
eng = Engine(dev); eng.load_state_dict(synth.make_state_dict(0))
nq, nm = 2000, 4000
v, m, _ = synth.make_eval_set(nq, nm, synth.BASE_SEED + 2)
dv = {k: v[k].to(dev) for k in ("frame_feats", "frame_mask")}
dm = {k: m[k].to(dev) for k in ("segment_feats", "segment_mask", "gt_moment", "m_duration")}

def timeit(name, fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"{name:34s} device {e0.elapsed_time(e1)/n:8.3f} ms   host enqueue {host:7.3f} ms", flush=True)

print("precision", eng.precision, "fused_ffn", os.environ.get("MADE_FUSED_FFN", "1"))
timeit("encode music 1000", lambda: eng.encode(_lib.MUSIC, dm["segment_feats"][:1000], dm["segment_mask"][:1000], want_f32=False))
timeit("encode video 1000", lambda: eng.encode(_lib.VIDEO, dv["frame_feats"][:1000], dv["frame_mask"][:1000], want_f32=False))
seq_m, _, pooled_m = eng.encode(_lib.MUSIC, dm["segment_feats"][:2000], dm["segment_mask"][:2000], want_f32=False)
seq_v, _, pooled_v = eng.encode(_lib.VIDEO, dv["frame_feats"], dv["frame_mask"], want_f32=False)
timeit("gallery_prepare 1000", lambda: eng.gallery_prepare(seq_m[:1000], dm["segment_mask"][:1000]))
kz, gram, bits = eng.gallery_prepare(seq_m[:1000], dm["segment_mask"][:1000])
q, vhat = eng.query_prepare(pooled_v)
timeit("xpool 2000x1000", lambda: eng.xpool_score(q, vhat, kz, gram, bits))
timeit("detr_detect 2000", lambda: eng.detr_detect(seq_v, dv["frame_mask"], seq_m, dm["segment_mask"][:2000], pooled_v))
for ds in ("1", "0"):
    os.environ["MADE_DETECT_STREAM"] = ds
    ev = GalleryEvaluator(eng, k=100, music_chunk=1000, video_chunk=1000)
    gt = torch.arange(nq, dtype=torch.int32, device=dev)
    timeit(f"whole job (detect stream {ds})", lambda: ev.run(dv, dm, gt), n=10)
    # true host cost of one step: enqueue onto an idle GPU (no back-pressure from a full launch queue)
    hs = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); ev.run(dv, dm, gt); hs.append(1e3 * (time.perf_counter() - t0))
    torch.cuda.synchronize()
    print(f"   host enqueue of one step onto an idle GPU: {min(hs):.3f} ms (min of 5), launches {ev.launches}", flush=True)
