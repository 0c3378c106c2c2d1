"""Timeline of one e2e step: when each chunk's host->device copy starts / ends on the copy stream and when the step's last
kernel ends, relative to the step start.  Diagnostics."""
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
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ev = GalleryEvaluator(eng, k=100, music_chunk=chunk, video_chunk=chunk)
for _ in range(5):
    ev.to_host(ev.run(hv, hm, gt, on_host=True))
marks = []
orig = eng.h2d_valid_rows


def wrapped(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    r = orig(*a, **k)
    e1.record()
    marks.append((e0, e1, r, t_host))
    return r


eng.h2d_valid_rows = wrapped
torch.cuda.synchronize()
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
s0.record()
out = ev.run(hv, hm, gt, on_host=True)
t_enq = time.perf_counter()
s1.record()
host = ev.to_host(out)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"chunk {chunk}: step wall {1e3 * (t1 - t0):.2f} ms, host enqueue {1e3 * (t_enq - t0):.2f} ms, device end {s0.elapsed_time(s1):.2f} ms")
for i, (e0, e1, nbytes, th) in enumerate(marks):
    print(f"  copy {i:2d}: issued by host at {1e3 * (th - t0):6.2f}  starts {s0.elapsed_time(e0):6.2f}  ends {s0.elapsed_time(e1):6.2f} ms  "
          f"{nbytes / 1e6:6.1f} MB  {nbytes / max(e0.elapsed_time(e1), 1e-6) / 1e6:5.1f} GB/s")
