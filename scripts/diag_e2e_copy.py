"""How fast do the batched valid-row copies of one job run by themselves (no kernels), against a plain bulk copy?  And the
e2e step at several chunk sizes.  Diagnostics."""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import synth
from mgsv_b200.engine import Engine
from mgsv_b200.pipeline import GalleryEvaluator
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
nq, nm = 2000, 4000
v, m, ids = synth.make_eval_set(nq, nm, synth.BASE_SEED + 2)
hv = {k: v[k].contiguous().pin_memory() for k in ("frame_feats", "frame_mask")}
hm = {k: m[k].contiguous().pin_memory() for k in ("segment_feats", "segment_mask", "gt_moment", "m_duration")}
gt = torch.arange(nq, dtype=torch.int32)
eng = Engine(dev); eng.load_state_dict(synth.make_state_dict(0))


def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


side = torch.cuda.Stream(dev)
side2 = torch.cuda.Stream(dev)
for chunk in (500, 4000):
    dv = torch.empty((min(chunk, nq),) + tuple(hv["frame_feats"].shape[1:]), device=dev)
    dm = torch.empty((min(chunk, nm),) + tuple(hm["segment_feats"].shape[1:]), device=dev)
    nbytes = [0]

    def copies(two=False):
        nbytes[0] = 0
        cur = torch.cuda.current_stream()
        side.wait_stream(cur); side2.wait_stream(cur)
        i = 0
        for s in range(0, nq, chunk):
            e = min(nq, s + chunk)
            with torch.cuda.stream(side2 if (two and i % 2) else side):
                nbytes[0] += eng.h2d_valid_rows(hv["frame_feats"][s:e], hv["frame_mask"][s:e], dv[:e - s])
            i += 1
        for s in range(0, nm, chunk):
            e = min(nm, s + chunk)
            with torch.cuda.stream(side2 if (two and i % 2) else side):
                nbytes[0] += eng.h2d_valid_rows(hm["segment_feats"][s:e], hm["segment_mask"][s:e], dm[:e - s])
            i += 1
        cur.wait_stream(side); cur.wait_stream(side2)
    for two in (False, True):
        ms = ev_time(lambda: copies(two))
        print(f"valid-row copies only, chunk {chunk}, {'two streams' if two else 'one stream'}: {ms:.2f} ms for {nbytes[0] / 1e6:.0f} MB = {nbytes[0] / ms / 1e6:.1f} GB/s", flush=True)
big_h = torch.empty(764 << 20, dtype=torch.uint8).pin_memory(); big_d = torch.empty(764 << 20, dtype=torch.uint8, device=dev)
ms = ev_time(lambda: big_d.copy_(big_h, non_blocking=True))
print(f"one bulk copy of 801 MB: {ms:.2f} ms = {(764 << 20) / ms / 1e6:.1f} GB/s", flush=True)
for chunk in (500,):
    ev = GalleryEvaluator(eng, k=100, music_chunk=chunk, video_chunk=chunk)
    ms = ev_time(lambda: ev.to_host(ev.run(hv, hm, gt, on_host=True)), reps=10)
    print(f"e2e step, chunk {chunk}: {ms:.2f} ms", flush=True)
