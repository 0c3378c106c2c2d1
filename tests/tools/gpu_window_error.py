"""Similarity error of the whole CUDA pipeline at configs[1] size against the CPU oracle on a window of the
2000 x 4000 matrices, for each precision mode.  Test infrastructure only (imports oracle/).

    python tests/tools/gpu_window_error.py [n_window_queries] [n_window_tracks]
"""
from __future__ import annotations

import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from mgsv_b200 import synth  # noqa: E402
from mgsv_b200.engine import Engine  # noqa: E402
from mgsv_b200.pipeline import GalleryEvaluator  # noqa: E402
from oracle import made_oracle as O  # noqa: E402


def main():
    wq = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    wm = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda:0")
    nq, nm = 2000, 4000
    sd = synth.make_state_dict(0)
    v, m, _ = synth.make_eval_set(nq, nm, synth.BASE_SEED + 2)
    qi = torch.arange(wq)
    ti = torch.cat([torch.arange(wm // 2), torch.arange(3000, 3000 + wm - wm // 2)])
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        _, vf = O.encode_video(sd, v["frame_feats"][qi], v["frame_mask"][qi])
        so, mf = O.encode_music(sd, m["segment_feats"][ti], m["segment_mask"][ti])
        smask = m["segment_mask"][ti]
        single, dual, _ = O.gallery_similarity(sd, vf, mf, so * smask.unsqueeze(-1), smask)
    gt = torch.arange(nq, dtype=torch.int32, device=dev)
    dv = {k: v[k].to(dev) for k in ("frame_feats", "frame_mask")}
    dm = {k: m[k].to(dev) for k in ("segment_feats", "segment_mask", "gt_moment", "m_duration")}
    for prec in ("fp16", "split"):
        eng = Engine(dev, precision=prec)
        eng.load_state_dict(sd)
        ev = GalleryEvaluator(eng, k=100, music_chunk=1000, video_chunk=1000)
        out = ev.run(dv, dm, gt, want_sims=True)
        torch.cuda.synchronize()
        for name, ref in (("single", single), ("dual", dual)):
            got = out[name][:wq][:, ti.to(dev)].double().cpu()
            d = (got - ref.double()).abs()
            sc = ref.abs().max().item()
            pe = (d <= 1e-3 * ref.double().abs() + 1e-5).double().mean().item()
            print(f"[{prec:5s}] {name:6s} window {wq}x{wm}: max {d.max().item() / sc:.3e} rms "
                  f"{d.pow(2).mean().sqrt().item() / sc:.3e} of scale {sc:.4f}; p99.9 "
                  f"{d.flatten().kthvalue(int(0.999 * d.numel())).values.item() / sc:.3e}; per-element rule {100 * pe:.2f} %",
                  flush=True)
        vf_err = (out["video_feats"][:wq].cpu() - vf).abs().max().item()
        mf_err = (out["music_feats"][ti.to(dev)].cpu() - mf).abs().max().item()
        print(f"[{prec:5s}] pooled embedding max|d|: video {vf_err:.3e} music {mf_err:.3e}", flush=True)
        del ev, eng


if __name__ == "__main__":
    main()
