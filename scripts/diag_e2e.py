"""Per-step wall times of the e2e path (no extra syncs).  Test infrastructure."""
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
# MADE_DIAG_CHUNKS=384,512,1000: sweep the ingest chunk size (tracks / videos per host->device copy + encode chunk)
chunks = [int(c) for c in os.environ.get("MADE_DIAG_CHUNKS", "512").split(",")]
for mode, chunk in [(mo, c) for mo in (sys.argv[1:] or ["dma", "zerocopy"]) for c in chunks]:
    ev = GalleryEvaluator(eng, k=100, music_chunk=chunk, video_chunk=chunk)
    ev.h2d_mode = mode
    mode = f"{mode} chunk {chunk}"
    ts = []
    for it in range(int(os.environ.get("MADE_DIAG_STEPS", "30"))):
        t0 = time.perf_counter()
        out = ev.to_host(ev.run(hv, hm, gt, on_host=True))
        ts.append(1e3 * (time.perf_counter() - t0))
    print(mode, " ".join(f"{t:.1f}" for t in ts), flush=True)
    tt = sorted(ts[6:])
    print(mode, f"after 6 warm-up steps: mean {sum(tt) / len(tt):.2f} median {tt[len(tt) // 2]:.2f} max {tt[-1]:.2f} "
                f"steps above 1.05 x median: {sum(1 for t in tt if t > 1.05 * tt[len(tt) // 2])} of {len(tt)}", flush=True)
    # CPU-side enqueue cost of one step (no sync until the end)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); o = ev.run(hv, hm, gt, on_host=True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(mode, f"enqueue {1e3*(t1-t0):.1f} ms, drain {1e3*(t2-t1):.1f} ms", flush=True)
