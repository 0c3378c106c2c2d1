"""Micro-benchmarks of the memory-bound kernels (BASELINE.json configs[2] and the HBM-roofline part
of the north star): achieved algorithmic GB/s against the measured HBM copy peak.

    python scripts/micro_bench.py > gpurun_out/micro_bench.json

Algorithmic bytes per kernel: SURVEY.md §8(d) / DESIGN.md §4.5.  CUDA-event timing, 3 warm-ups,
outputs larger than the 126 MB L2 where the size allows (the 1M-pair case is launch-bound and is
reported as such).  Measurement infrastructure, not product code.
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops, synth  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timed_graph(fn, launches=32, reps=10):
    """Per-launch time of `launches` back-to-back calls replayed from ONE CUDA graph: the device-side cost of a small
    kernel without the ~20 us of Python / ctypes / allocator time a single eager call spends on the host."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(launches):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * launches)


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    rows = []

    def rec(name, ms, nbytes, note=""):
        gbs = nbytes / (ms / 1e3) / 1e9
        rows.append(dict(kernel=name, ms=ms, algorithmic_bytes=nbytes, gbs=gbs, frac_of_hbm_peak=gbs / peak, note=note))

    for n in (1000, 16384):
        a, b, logits = synth.make_span_pairs(n, n + 8, 3)
        a, b = a.to(dev), b.to(dev)
        b = b[b[:, 1] != 0][:n].contiguous()      # matcher.py:59 drops zero-width targets; keep n of them
        se_a, se_b = ops.span_cw_to_se(a), ops.span_cw_to_se(b)
        prob = logits.softmax(-1)[:, 0].contiguous().to(dev)
        m = b.shape[0]
        note = "launch-bound (4 MB out)" if n == 1000 else "1.07 GB out"
        rec(f"giou {n}x{m}", timed(lambda: ops.generalized_temporal_iou(se_a, se_b, check=False)), n * m * 4 + (n + m) * 8, note)
        rec(f"matcher_cost {n}x{m}", timed(lambda: ops.matcher_cost(prob, a, b)), n * m * 4 + (n + m) * 8 + n * 4, note)
        rec(f"temporal_iou {n}x{m}", timed(lambda: ops.temporal_iou(se_a, se_b)), 2 * n * m * 4 + (n + m) * 8, note)
        if n == 1000:      # configs[2] itself: 1M pairs per call, 32 calls replayed from one CUDA graph
            gnote = "configs[2]: 1M pairs per launch, 32 launches in one CUDA graph (device time per launch)"
            for nme, fn, nb in (("giou", lambda: ops.generalized_temporal_iou(se_a, se_b, check=False), n * m * 4 + (n + m) * 8),
                                ("matcher_cost", lambda: ops.matcher_cost(prob, a, b), n * m * 4 + (n + m) * 8 + n * 4),
                                ("temporal_iou", lambda: ops.temporal_iou(se_a, se_b), 2 * n * m * 4 + (n + m) * 8)):
                try:
                    rec(f"{nme} {n}x{m} (graph)", timed_graph(fn), nb, gnote)
                except Exception as exc:      # a host-side check inside the op that cannot be captured
                    sys.stderr.write(f"graph timing of {nme} skipped: {exc}\n")
                    torch.cuda.synchronize()
    for nq, nm in ((2000, 4000), (8192, 16384)):
        single = torch.randn(nq, nm, device=dev)
        dual = torch.randn(nq, nm, device=dev)
        gt = torch.randint(0, nm, (nq,), device=dev, dtype=torch.int32)
        rec(f"rank_topk k=100 {nq}x{nm}", timed(lambda: ops.rank_topk(single, dual, gt, None, k=100)),
            nq * nm * 8 + nq * 100 * 12, "fp64-sum keys, dedup-aware rank + exact top-100")
        if nq == 2000:     # the job's own call: device time per launch without the host's share of an eager call
            try:
                rec(f"rank_topk k=100 {nq}x{nm} (graph)", timed_graph(lambda: ops.rank_topk(single, dual, gt, None, k=100), launches=16),
                    nq * nm * 8 + nq * 100 * 12, "16 launches in one CUDA graph (device time per launch)")
            except Exception as exc:
                sys.stderr.write(f"graph timing of rank_topk skipped: {exc}\n")
                torch.cuda.synchronize()
    n = 1 << 20
    lg = torch.randn(n, 2, device=dev)
    sp = torch.rand(n, 2, device=dev)
    gtm = torch.sort(torch.rand(n, 2, device=dev) * 240, dim=-1)[0]
    md = torch.rand(n, device=dev) * 200 + 40
    rec("moment_postproc 1M", timed(lambda: ops.moment_postproc(lg, sp, gtm, md)), n * (8 + 8 + 8 + 4 + 16),
        "eager call: four output allocations + ctypes on the host per launch")
    try:
        n4 = 1 << 22
        lg4, sp4 = torch.randn(n4, 2, device=dev), torch.rand(n4, 2, device=dev)
        gtm4 = torch.sort(torch.rand(n4, 2, device=dev) * 240, dim=-1)[0]
        md4 = torch.rand(n4, device=dev) * 200 + 40
        rec("moment_postproc 4M (graph)", timed_graph(lambda: ops.moment_postproc(lg4, sp4, gtm4, md4), launches=16),
            n4 * (8 + 8 + 8 + 4 + 16), "16 launches in one CUDA graph (device time per launch; 185 MB per launch, larger than the L2)")
    except Exception as exc:
        sys.stderr.write(f"graph timing of moment_postproc skipped: {exc}\n")
        torch.cuda.synchronize()
    out = dict(hbm_peak_gbs=peak, peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback", kernels=rows)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
