#!/usr/bin/env python
"""Benchmark of the MaDe inference + scoring hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU

A "step" = one pass of the whole job over one batch of synthetic input: encode the query videos and
the 4000-track gallery, full-gallery X-Pool + dual similarity, fp64 ranking/top-100, DETR moment
detection for the paired track, IoU (BASELINE.json configs[1]; fp16 GEMM operands - (hi, lo) pairs where
the similarity error is made - fp32 accumulate).  `value` times it with inputs resident in HBM; `e2e` times it
from pinned host buffers (fp32 feature tensors, the reference-facing dtype): the valid feature rows are moved
host->device inside the timed region (copy engines, overlapped with the kernels of the previous chunk), and the
results are read back to the host.

N = 1: 2000 queries x 4000 tracks.  N > 1: the 4000-track gallery is sharded over the ranks; `--scaling weak`
(default) gives every rank 2000 queries of its own (2000 N queries against the sharded 4k gallery: per-GPU work is
fixed - 2000 x 4000 pairs scored, 2000 queries encoded and detected - while the gallery encode shrinks),
`--scaling strong` runs the same 2000 x 4000 job on N GPUs (a 7 ms job: latency-bound beyond 2 GPUs; reported as
`strong_same_job` beside the weak line).  Before timing, every N > 1 run checks that the sharded path returns the
single-GPU path's ranks / top-k / spans bit for bit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "queries/sec (match + moment detect, 4k-track gallery)"
UNIT = "queries/s"
N_QUERIES, N_TRACKS, TOPK = 2000, 4000, 100
WORKLOAD = "MaDe full-gallery inference: 2000 synthetic query videos x 4000 music tracks (configs[1])"
# algorithmic FLOPs of the reference's dense formulation (2 M N K, padded shapes)
F_XPOOL_PAIR = 360_960.0
F_XPOOL_PAIR_EXEC = 2.0 * (96 * 256 + 96 * 112 + 96 * 256)       # S, [T|L], Y MMAs of the folded algebra
F_VIDEO_GEMM = (100.86 - 2.56) * 1e6      # per video, attention (QK^T, PV) excluded
F_MUSIC_GEMM = (210.76 - 9.44) * 1e6      # per track
F_XPOOL_KV = 25.17e6                      # per track (K/V projections); 0.13e6 per query (q projection)
F_DETR_GEMM = (2 * (251.47 - 21.82) + 6 * (40.26 - 0.15) + 8.9) * 1e6   # per query
F_TOTAL_JOB = 5.54e12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="made_b200", choices=["made_b200", "reference"])
    ap.add_argument("--queries", type=int, default=N_QUERIES, help="queries (per GPU with --scaling weak)")
    ap.add_argument("--tracks", type=int, default=N_TRACKS)
    ap.add_argument("--chunk", type=int, default=2000, help="tracks / videos per ingest+encode chunk (device-resident steps; "
                    "2000: 6 - 7 waves of 128-row tiles per GEMM launch, half the launches of 1000)")
    ap.add_argument("--e2e-chunk", type=int, default=768, help="chunk size of the host-input (e2e) steps: the e2e step is "
                    "PCIe bound, and smaller chunks shorten the fill / drain of the copy-compute pipeline (measured with "
                    "the gallery-first copy order, ms per step: 256 15.6, 384 15.6, 512 14.9, 768 14.7, 1000 15.6)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1 only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--detect-topk", type=int, default=5, help="N = 1: also time retrieve-then-detect (DETR on the k best "
                    "retrieved tracks of every query, SURVEY.md 8f rank 2); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded-vs-single-GPU parity check (N > 1)")
    return ap.parse_args()


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# -------------------------------------------------------------------------------------------------
# cpu_baseline (inside the made_b200 arm) runs a BOUNDED sample of the workload.  The oracle's per-unit costs
# fall with the batch size until the host threads are busy (measured on 8 cores: full-job time 64 s composed from
# a 48 x 256 sample, 35 s from 192 x 1024, 33 s from 384 x 2048), so the sample is large enough to sit on that
# plateau: a smaller one would flatter the GPU/CPU ratio.  `--impl reference` runs the WHOLE job instead.
CPU_SAMPLE_Q, CPU_SAMPLE_M = 192, 1024


def cpu_reference_time(n_q_sample: int, n_m_sample: int, n_queries: int, n_tracks: int, repeats: int = 1):
    """Time the reference algorithm (oracle port, fp32, all host threads) on bounded samples of the
    workload and compose the full-job time from the measured per-unit costs (each stage is linear
    in its unit count): video encode per query, music encode per track, X-Pool + similarity per
    (query, track) pair, DETR + post-processing per query, ranking per query."""
    from mgsv_b200 import synth
    from oracle import made_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_state_dict(0)
    v, m, ids = synth.make_eval_set(n_q_sample, n_m_sample, synth.BASE_SEED + 2)
    best = None
    for _ in range(repeats + 1):   # first pass = warm-up
        t = {}
        with torch.no_grad():
            t0 = time.perf_counter()
            fo, vf = O.encode_video(sd, v["frame_feats"], v["frame_mask"])
            t["video_enc"] = (time.perf_counter() - t0) / n_q_sample
            t0 = time.perf_counter()
            so, mf = O.encode_music(sd, m["segment_feats"], m["segment_mask"])
            t["music_enc"] = (time.perf_counter() - t0) / n_m_sample
            t0 = time.perf_counter()
            single, dual, total = O.gallery_similarity(sd, vf, mf, so, m["segment_mask"], track_chunk=64)
            t["pair"] = (time.perf_counter() - t0) / (n_q_sample * n_m_sample)
            t0 = time.perf_counter()
            O.recall_metrics(total, ids["music_ids"], np.arange(n_q_sample))
            t["rank_pair"] = (time.perf_counter() - t0) / (n_q_sample * n_m_sample)
            t0 = time.perf_counter()
            nd = min(n_q_sample, n_m_sample)
            src = torch.cat([fo[:nd], so[:nd]], 1)
            mask = torch.cat([v["frame_mask"][:nd], m["segment_mask"][:nd]], 1)
            hs, _ = O.detr_forward(sd, src, mask, O.position_embedding_sine(mask), vf[:nd].unsqueeze(1))
            om = O.calc_output(sd, hs, fo[:nd])
            st, ed, sc = O.moment_postproc(om["pred_logits"], om["pred_spans"])
            O.detr_iou(st, ed, m["gt_moment"][:nd], m["m_duration"][:nd])
            t["detr"] = (time.perf_counter() - t0) / nd
        tot = (t["video_enc"] + t["detr"]) * n_queries + t["music_enc"] * n_tracks + \
              (t["pair"] + t["rank_pair"]) * n_queries * n_tracks
        if best is None or tot < best[0]:
            best = (tot, t)
    return best[0], best[1]


class CpuFullJob:
    """The whole 2000 x 4000 job through the oracle port (reference algorithm, fp32, all host threads), every
    stage on the full unit counts: batched encoders (batch 256, the reference's training batch; its eval batch of
    40 is slower), gallery X-Pool in 64-track chunks so that the [N_m, N_v, 256] intermediate the reference
    materialises (8.2 GB) stays at 131 MB, fp64 sum, Recall_metrics' argsort + python walk, DETR + post-processing
    + IoU for the paired tracks."""

    def __init__(self, n_queries: int, n_tracks: int):
        from mgsv_b200 import synth
        torch.set_num_threads(os.cpu_count() or 1)
        self.sd = synth.make_state_dict(0)
        self.v, self.m, self.ids = synth.make_eval_set(n_queries, n_tracks, synth.BASE_SEED + 2)
        self.nq, self.nm = n_queries, n_tracks

    def step(self):
        from oracle import made_oracle as O
        sd, v, m, nq, nm = self.sd, self.v, self.m, self.nq, self.nm
        t = {}
        bs = 256
        with torch.no_grad():
            t0 = time.perf_counter()
            fo, vf = zip(*[O.encode_video(sd, v["frame_feats"][s:s + bs], v["frame_mask"][s:s + bs]) for s in range(0, nq, bs)])
            fo, vf = torch.cat(fo), torch.cat(vf)
            t["video_enc"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            so, mf = zip(*[O.encode_music(sd, m["segment_feats"][s:s + bs], m["segment_mask"][s:s + bs]) for s in range(0, nm, bs)])
            so, mf = torch.cat(so), torch.cat(mf)
            t["music_enc"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            single, dual, total = O.gallery_similarity(sd, vf, mf, so, m["segment_mask"], track_chunk=64)
            t["gallery_similarity"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            O.recall_metrics(total, self.ids["music_ids"], np.arange(nq))
            t["recall_metrics"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            for s in range(0, nq, bs):
                e = min(nq, s + bs)
                src = torch.cat([fo[s:e], so[s:e]], 1)
                mask = torch.cat([v["frame_mask"][s:e], m["segment_mask"][s:e]], 1)
                hs, _ = O.detr_forward(sd, src, mask, O.position_embedding_sine(mask), vf[s:e].unsqueeze(1))
                om = O.calc_output(sd, hs, fo[s:e])
                st, ed, sc = O.moment_postproc(om["pred_logits"], om["pred_spans"])
                O.detr_iou(st, ed, m["gt_moment"][s:e], m["m_duration"][s:e])
            t["detr"] = time.perf_counter() - t0
        return sum(t.values()), t


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line (the driver parses it): keep a private handle to the real
    stdout and point fd 1 at stderr, so that library chatter written to stdout during the run (NCCL's
    version banner at NCCL_DEBUG >= VERSION, extension build logs) cannot land beside it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args, rank: int):
    """The reference's CPU implementation of the path, timed for real on the whole configs[1] job (one full job
    per step; the run stops early - with the steps done so far - once it has used ~4 minutes)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    job = CpuFullJob(args.queries, args.tracks)
    steps_ms, parts = [], None
    t_begin = time.perf_counter()
    budget_s = 240.0
    for i in range(args.warmup + args.steps):
        tot, p = job.step()
        if i >= args.warmup or (time.perf_counter() - t_begin) > 0.5 * budget_s:
            steps_ms.append(tot * 1e3)     # a slow host: warm-up passes count once half the budget is gone
            parts = p
        if time.perf_counter() - t_begin > budget_s and steps_ms:
            break
    ms = float(np.mean(steps_ms))
    value = args.queries / (ms / 1e3)
    sample = (f"the WHOLE {args.queries} x {args.tracks} job per step through the oracle port (reference algorithm, fp32, "
              f"torch CPU on {cores} threads): encoders in batches of 256, gallery X-Pool in 64-track chunks, fp64 sum, "
              "argsort + python dedup walk, DETR + IoU for the paired tracks")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(steps_ms), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.gpus > 1 and args.scaling == "weak" else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_queries": args.queries, "n_tracks": args.tracks,
                   "note": "one host runs one 2000 x 4000 job per step whatever --gpus says (the reference has no "
                           "working multi-process evaluation, SURVEY.md 2.3)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "parts_s": parts},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# -------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            # nvidia-smi takes 0.1 - 0.5 s to attach to the driver and has stalled a concurrent step for
            # tens of ms while doing so: wait for its first sample so that this happens before the
            # warm-up, not inside the timed region
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0 and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Timed region starts here: only samples taken from now on are reported."""
        self.first = len(self.lines)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[self.first:] or self.lines    # a region shorter than one period: keep the warm-up samples
        for ln in lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])), mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, val in zip(names, p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def executed_gemm_flops(n_frames: float, n_segments: float, n_queries: int, n_tracks: int, n_detr_tokens: float):
    """FLOPs the GEMM-family kernels execute usefully per step on token-packed inputs (each product counted once:
    the extra passes of the split-precision GEMMs are overhead, not work): valid video / music tokens, the padded
    X-Pool operand rows (N_m x 96), valid DETR encoder tokens, one decoder query per sequence."""
    video_tok = 2.0 * (512 * 256 + 256 * 768 + 256 * 256 + 2 * 256 * 1024 + 256 * 256)
    music_tok = 2.0 * (768 * 256 + 256 * 768 + 256 * 256 + 2 * 256 * 1024 + 256 * 256)
    xp_row = 2.0 * 256 * 768
    gram = 2.0 * 96 * 96 * 256
    detr_tok = 2 * 2.0 * (256 * 512 + 256 * 256 + 256 * 256 + 2 * 256 * 1024)
    dec_q = 6 * 2.0 * (256 * 256 + 256 * 2048 + 2048 * 256 + 2 * 256 * 1024) + 6 * 2 * 2.0 * 256 * 256
    return (n_frames * video_tok + n_segments * music_tok + n_tracks * (96 * xp_row + gram) + n_queries * 2.0 * 256 * 256 +
            n_detr_tokens * detr_tok + n_queries * dec_q)


def tensor_pipe_note():
    """ncu `sm__pipe_tensor_cycles_active` of the newest committed captures (profiles/*_tensor_pipe.json)."""
    try:
        cands = sorted(f for f in os.listdir(os.path.join(REPO, "profiles")) if f.endswith("_tensor_pipe.json"))
        if cands:
            return json.load(open(os.path.join(REPO, "profiles", cands[-1])))
    except Exception:
        pass
    return None


def step_dram_note():
    """ncu DRAM bytes of one step per kernel (newest committed profiles/*_step_dram.json: `dram__bytes_read.sum +
    dram__bytes_write.sum` of every launch between two rank_topk kernels)."""
    try:
        cands = sorted(f for f in os.listdir(os.path.join(REPO, "profiles")) if f.endswith("_step_dram.json"))
        if cands:
            d = json.load(open(os.path.join(REPO, "profiles", cands[-1])))
            d.pop("per_kernel", None)
            d["file"] = "profiles/" + cands[-1]
            return d
    except Exception:
        pass
    return None


# -------------------------------------------------------------------------------------------------
def main():
    args = parse()
    claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from mgsv_b200 import _lib, synth
    from mgsv_b200.engine import Engine
    from mgsv_b200.parallel import ShardedEvaluator, shard_bounds
    from mgsv_b200.pipeline import GalleryEvaluator

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: made_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # NCCL errors visible (its banner goes to stderr: claim_stdout)
        dist.init_process_group("nccl", device_id=dev)
    weak = world > 1 and args.scaling == "weak"

    nq, nm = args.queries, args.tracks
    m0, m1 = shard_bounds(nm, rank, world)
    # every rank draws the full synthetic configs[1] set from the same seed and keeps its slices
    v, m, ids = synth.make_eval_set(nq, nm, synth.BASE_SEED + 2)
    pin = lambda d, keys, a, b: {k: d[k][a:b].contiguous().pin_memory() for k in keys}
    vkeys, mkeys = ("frame_feats", "frame_mask"), ("segment_feats", "segment_mask", "gt_moment", "m_duration")
    host_m = pin(m, mkeys, m0, m1)
    dev_m = {k: t.to(dev) for k, t in host_m.items()}

    eng = Engine(dev)
    eng.load_state_dict(synth.make_state_dict(0))
    ev = GalleryEvaluator(eng, k=TOPK, music_chunk=args.chunk, video_chunk=args.chunk)
    sharded = ShardedEvaluator(ev, rank, world) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: the sharded path must return the single-GPU path's results bit for bit (checked before timing) ----
    sharded_parity = None
    if world > 1 and not args.no_check:
        q0, q1 = shard_bounds(nq, rank, world)
        full_v = {k: v[k].to(dev) for k in vkeys}
        full_m = {k: m[k].to(dev) for k in mkeys}
        gt_full = torch.arange(nq, dtype=torch.int32)
        one = ev.run(full_v, full_m, gt_full.to(dev))
        one = {k: one[k].clone() for k in ("rank", "topk_idx", "topk_score", "pred_st", "pred_ed", "iou", "score")}
        sh = sharded.run({k: t[q0:q1].contiguous() for k, t in full_v.items()}, dev_m, gt_full, nq, nm)
        torch.cuda.synchronize()
        ok = all(torch.equal(sh[k], one[k]) for k in ("rank", "topk_idx", "topk_score")) and \
            all(torch.equal(sh[k], one[k][q0:q1]) for k in ("pred_st", "pred_ed", "iou", "score"))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise RuntimeError(f"rank {rank}: sharded evaluation differs from the single-GPU path")
        sharded_parity = "rank / top-k / spans / IoU of the sharded path bit-identical to the single-GPU path (checked before timing)"
        del full_v, full_m, one, sh

    # ---- this rank's queries ----
    if weak:     # 2000 fresh queries per rank; query g of the job is paired with track g % n_tracks
        nq_total = nq * world
        q0, q1 = rank * nq, (rank + 1) * nq
        v_mine = synth.make_videos(nq, synth.BASE_SEED + 2 + 1000 * rank) if rank > 0 else v
        gt_col = torch.arange(nq_total, dtype=torch.int32) % nm
    else:
        nq_total = nq
        q0, q1 = shard_bounds(nq, rank, world)
        v_mine = {k: v[k][q0:q1] for k in vkeys}
        gt_col = torch.arange(nq, dtype=torch.int32)
    host_v = pin(v_mine, vkeys, 0, q1 - q0)
    dev_v = {k: t.to(dev) for k, t in host_v.items()}
    n_frames_local = float(host_v["frame_mask"].sum().item())
    n_segments_local = float(host_m["segment_mask"].sum().item())
    seg_len_all = m["segment_mask"].sum(1)
    del v, m

    def make_step(hv, dv, gt, n_total):
        def step(on_host: bool):
            if sharded is not None:
                out = sharded.run(hv if on_host else dv, host_m if on_host else dev_m, gt, n_total, nm, on_host=on_host,
                                  gather_results=False)
                # a step ends when its results exist on every stream: without this the host runs ahead of the
                # NCCL / ingest streams and steps start to interleave pathologically (measured: 3x slower)
                torch.cuda.synchronize()
            else:
                out = ev.run(hv if on_host else dv, host_m if on_host else dev_m, gt, on_host=on_host)
            return ev.to_host(out) if on_host else out
        return step

    step = make_step(host_v, dev_v, gt_col, nq_total)

    def timed(step_fn, on_host: bool, steps: int, warmup: int, on_start=None, profile: bool = False, min_warm_s: float = 0.0):
        """Device time of `steps` back-to-back steps (CUDA events, barrier + synchronize on both sides,
        max over ranks).  Host-input steps end with a device->host read, so they are also timed one
        by one with the wall clock: the shared hosts of this pool stall a step now and then
        (PCIe / OS jitter, 25 - 800 ms, seen with every H2D mode), which the per-step list exposes."""
        for _ in range(warmup):
            step_fn(on_host)
        # A fresh box gives its first process a slow start (first bench of a box: 8.5 - 11 ms per step, the next
        # processes 7.1 - 7.2 ms with every kernel at its usual duration): keep warming up, untimed, until the device has
        # been busy for `min_warm_s` seconds in this process.  The K timed steps below are unaffected.
        # (a fixed step count, the same on every rank: the sharded step contains collectives)
        for _ in range(int(min_warm_s / 7.5e-3)):
            step_fn(on_host)
        barrier()
        if on_start:
            on_start()
        if profile:
            _lib.prof_collect()
            _lib.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_step = []
        t_wall = time.perf_counter()
        e0.record()
        for _ in range(steps):
            ts = time.perf_counter()
            step_fn(on_host)
            per_step.append(1e3 * (time.perf_counter() - ts))
        e1.record()
        barrier()
        wall = time.perf_counter() - t_wall
        prof = None
        if profile:
            _lib.prof_enable(False)
            prof = _lib.prof_collect()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms, float(np.median(per_step)), float(np.max(per_step))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), wall / steps * 1e3, float(t[1].item()), float(t[2].item()), prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, _, _, _, _ = timed(step, False, args.steps, max(args.warmup, 3), on_start=sampler.mark, min_warm_s=1.5)
    launches = ev.launches
    clocks = sampler.stop()      # sampled during the device-resident timed region (20 ms period)
    # kernel-family times: a second, shorter timed region with CUDA events around every tensor-core kernel launch
    # (on the launching stream); kept apart from the headline region so that its ~300 event records per step
    # cannot perturb `value`
    # ... and run on ONE stream (no ingest / detection overlap): a kernel's event-bracketed duration otherwise includes the
    # time it waits for SMs held by the other streams' kernels (6.6 ms of "GEMM time" in a 7.2 ms step)
    prof_steps = min(args.steps, 5)
    ev_serial = GalleryEvaluator(eng, k=TOPK, music_chunk=args.chunk, video_chunk=args.chunk, single_stream=True)
    if sharded is not None:
        serial_sharded = ShardedEvaluator(ev_serial, rank, world)

        def step_serial(on_host: bool):
            out = serial_sharded.run(dev_v, dev_m, gt_col, nq_total, nm, on_host=False, gather_results=False)
            torch.cuda.synchronize()
            return out
    else:
        def step_serial(on_host: bool):
            return ev_serial.run(dev_v, dev_m, gt_col, on_host=False)
    ms_prof, _, _, _, prof = timed(step_serial, False, prof_steps, 2, profile=True)

    strong = None
    if weak:     # the same 2000 x 4000 job on N GPUs, for the record
        sq0, sq1 = shard_bounds(nq, rank, world)
        v_s, _, _ = synth.make_eval_set(nq, nq, synth.BASE_SEED + 2) if False else (None, None, None)
        sv = synth.make_videos(nq, synth.BASE_SEED + 2)
        dsv = {k: sv[k][sq0:sq1].contiguous().to(dev) for k in vkeys}
        s_step = make_step(None, dsv, torch.arange(nq, dtype=torch.int32), nq)
        s_ms, _, _, _, _ = timed(s_step, False, min(args.steps, 10), 3)
        strong = {"workload": f"{nq} queries x {nm} tracks on {world} GPUs (gallery and queries sharded)",
                  "ms_per_step": s_ms, "value": nq / (s_ms / 1e3), "unit": UNIT}
        del sv, dsv

    # ---- retrieve-then-detect (serving mode): the same job + one moment per (query, retrieved track) for the top-k ----
    rtd = None
    if world == 1 and args.detect_topk > 0:
        kd = args.detect_topk
        ms_rtd, _, _, _, _ = timed(lambda on_host: ev.run(dev_v, dev_m, gt_col, detect_topk=kd), False, min(args.steps, 5), 2)
        rtd = {"k_det": kd, "ms_per_step": ms_rtd, "value": nq_total / (ms_rtd / 1e3), "unit": UNIT,
               "extra_detections_per_step": nq_total * kd,
               "extra_ms_per_1000_detections": (ms_rtd - ms_dev) / (nq_total * kd / 1e3),
               "note": "configs[1] job + DETR moment detection on the k_det best retrieved tracks of every query "
                       "(GalleryEvaluator.run(detect_topk=k)); checked against the oracle in "
                       "tests/test_gpu_parity.py::test_retrieve_then_detect_vs_oracle"}

    e2e = None
    if not args.no_e2e:
        # reference point for the e2e number: plain pinned-host -> device copy rate of this box
        probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe_h, non_blocking=True)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            probe_d.copy_(probe_h, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        link_gbs = 3 * (256 << 20) / (p0.elapsed_time(p1) / 1e3) / 1e9
        del probe_h, probe_d
        ev_e2e = GalleryEvaluator(eng, k=TOPK, music_chunk=args.e2e_chunk, video_chunk=args.e2e_chunk)
        sharded_e2e = ShardedEvaluator(ev_e2e, rank, world) if world > 1 else None

        def step_e2e(on_host: bool):
            if sharded_e2e is not None:
                out = sharded_e2e.run(host_v, host_m, gt_col, nq_total, nm, on_host=True, gather_results=False)
                torch.cuda.synchronize()
            else:
                out = ev_e2e.run(host_v, host_m, gt_col, on_host=True)
            return ev_e2e.to_host(out)
        # the first steps of a process run 0.3 - 0.6 ms slow (staging buffers, pinned pages, allocator pools): 10 untimed
        # warm-up steps instead of 3 (scripts/diag_e2e.py: 15.5 15.2 15.2 14.9 14.9 ... per step)
        e2e_warm = max(args.warmup, 10)
        ms_e2e, wall_e2e, med_e2e, max_e2e, _ = timed(step_e2e, True, args.steps, e2e_warm)
        ev.h2d_mode = ev_e2e.h2d_mode
        # every host-input step ends with a synchronous device->host read, so its wall time is its
        # end-to-end time: median and max are reported beside the mean
        jitter = max_e2e > 2.0 * med_e2e
        tensors = list(host_v.values()) + list(host_m.values())
        h2d_padded = sum(t.numel() * t.element_size() for t in tensors) + gt_col.numel() * 4
        # bytes that actually cross PCIe: only the rows whose mask is 1 are transferred
        small = host_v["frame_mask"].numel() * 4 + gt_col.numel() * 4 + \
            sum(host_m[k].numel() * host_m[k].element_size() for k in ("segment_mask", "gt_moment", "m_duration"))
        esz = 2 if ev.h2d_mode == "dma16" else 4
        h2d = int(n_frames_local) * 512 * esz + int(n_segments_local) * 768 * esz + small
        d2h = (q1 - q0) * (4 + TOPK * 4 + 4 * 4)
        tot = torch.tensor([float(h2d), float(h2d_padded), float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        e2e = {"value": nq_total / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
               "statistic": "K back-to-back steps between CUDA events (each step ends with its device->host read)",
               "ms_per_step_median": med_e2e, "ms_per_step_max": max_e2e, "host_jitter_seen": bool(jitter),
               "wall_ms_per_step": wall_e2e, "warmup_steps": e2e_warm,
               "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[2].item()),
               "h2d_bytes_if_padded_rows_were_copied": int(tot[1].item()),
               "h2d_link_gbs_measured": link_gbs, "h2d_mode": ev.h2d_mode, "chunk": args.e2e_chunk,
               "h2d_bound_ms": (h2d / (link_gbs * 1e9) * 1e3) if link_gbs else None,
               "host_dtype": "f32 features (reference-facing dtype) in pinned host memory; only the valid rows cross "
                             "PCIe (copy engines, one batched copy per chunk)" +
                             ("; rounded to fp16 by host threads first" if ev.h2d_mode == "dma16" else ""),
               "host_threads": ev.host_threads if ev.h2d_mode == "dma16" else 0}

    # ---- roofline of the dominant kernel FAMILY: the tcgen05 GEMMs (gemm_tc_kernel + ffn_fused_kernel) ----
    peaks = {}
    pk_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline = None
    dram = step_dram_note()
    if prof:
        per_step = {k: (ms / prof_steps, n / prof_steps) for k, (ms, n) in prof.items()}
        gemm_ms = per_step["gemm"][0] + per_step["ffn"][0]
        gemm_n = per_step["gemm"][1] + per_step["ffn"][1]
        nq_loc, nm_loc = q1 - q0, m1 - m0
        pairs = float(nq_total) * nm_loc                       # (query, track) pairs this rank scores per step
        alg = nq_loc * (F_VIDEO_GEMM + F_DETR_GEMM) + nm_loc * (F_MUSIC_GEMM + F_XPOOL_KV) + nq_total * 0.13e6
        det_tok = n_frames_local + float(seg_len_all[(gt_col[q0:q1].long())].sum().item())
        exe = executed_gemm_flops(n_frames_local, n_segments_local, nq_loc, nm_loc, det_tok)
        ach = alg / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        xp_ms, xp_n = per_step["xpool"]
        xp_ach = F_XPOOL_PAIR * pairs / (xp_ms / 1e3) / 1e12 if xp_ms > 0 else 0.0
        xp_exe = F_XPOOL_PAIR_EXEC * pairs / (xp_ms / 1e3) / 1e12 if xp_ms > 0 else 0.0
        fam_total = sum(ms for ms, _ in per_step.values())
        roofline = {
            "kernel": "GEMM family: gemm_tc_kernel + ffn_fused_kernel (every Linear of the encoders, X-Pool operand "
                      "projections, DETR; the largest share of the step)",
            "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            "how": "algorithmic FLOPs of the family per step (SURVEY.md 8d dense formulation, attention and X-Pool pairs "
                   "excluded: 2.52 TF at 2000 x 4000) / summed CUDA-event duration of the family's launches per step "
                   "(events on the launching stream around each launch; separate short timed region in which the whole step "
                   "runs on ONE stream so that no other stream's kernels share the SMs); rank 0's share at N > 1",
            "algorithmic_flops_per_step": alg, "executed_flops_per_step": exe,
            "executed_tflops": exe / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0,
            "executed_frac": exe / (gemm_ms / 1e3) / 1e12 / peak_tf if gemm_ms > 0 else 0.0,
            "kernel_ms_per_step": gemm_ms, "launches_per_step": gemm_n,
            "ffn_fused_ms_per_step": per_step["ffn"][0], "ffn_fused_launches_per_step": per_step["ffn"][1],
            "share_of_step": gemm_ms / ms_prof,
            "share_of_tensor_kernel_time": gemm_ms / fam_total if fam_total > 0 else None,
            # DRAM bytes of the family per launch (average over the step's launches) from the committed ncu launch list;
            # `hbm_view` relates the family's DRAM bytes per step to its live kernel time
            "traffic": (dram["gemm_family"]["dram_read_bytes"] + dram["gemm_family"]["dram_write_bytes"]) /
                       max(dram["gemm_family"]["launches"], 1) if dram and world == 1 else None,
            "hbm_view": {"dram_bytes_per_step_ncu": dram["gemm_family"]["dram_read_bytes"] + dram["gemm_family"]["dram_write_bytes"],
                         "achieved_gbs": (dram["gemm_family"]["dram_read_bytes"] + dram["gemm_family"]["dram_write_bytes"]) /
                                         (gemm_ms / 1e3) / 1e9,
                         "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                         "frac": (dram["gemm_family"]["dram_read_bytes"] + dram["gemm_family"]["dram_write_bytes"]) /
                                 (gemm_ms / 1e3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                         "whole_step_dram_bytes_ncu": dram["step_dram_bytes"],
                         "whole_step_frac": dram["step_dram_bytes"] / (ms_dev / 1e3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                         "source": dram["file"],
                         "note": "K = N = 256 linears on (hi, lo) activation pairs move 2 KB per row for 0.39 MFLOP (3 passes): "
                                 "192 flop per byte, the ridge of this machine (1413 TF/s / 6.46 TB/s = 219) - the family is "
                                 "neither tensor- nor HBM-saturated; DESIGN.md 4.1 lists what bounds a tile"}
                        if dram and world == 1 and gemm_ms > 0 else None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
            "xpool": {"kernel": "xpool_score_kernel", "achieved": xp_ach, "frac": xp_ach / peak_tf,
                      "executed_tflops": xp_exe, "executed_frac": xp_exe / peak_tf, "kernel_ms_per_step": xp_ms,
                      "launches_per_step": xp_n, "share_of_step": xp_ms / ms_prof,
                      "algorithmic_flops_per_pair": F_XPOOL_PAIR, "executed_flops_per_pair": F_XPOOL_PAIR_EXEC},
            "other_families_ms_per_step": {"attention": per_step["attn"][0], "rank_topk": per_step["rank"][0]},
            "serial_step_ms": ms_prof,
            "streams_overlap_note": "family times come from a single-stream pass (serial_step_ms per step); the timed "
                                    "headline step overlaps ingest, scoring and detection on three streams",
            "tensor_pipe_active_ncu": tensor_pipe_note(),
            "whole_step_tflops": F_TOTAL_JOB * (nq_total / N_QUERIES) / (ms_dev / 1e3) / 1e12 if nm == N_TRACKS else None,
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tot, parts = cpu_reference_time(CPU_SAMPLE_Q, CPU_SAMPLE_M, nq, nm, repeats=1)
        cpu_baseline = {"value": nq / tot, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"oracle port (fp32, all host threads) on {CPU_SAMPLE_Q} queries x {CPU_SAMPLE_M} tracks, best of 2; full-job "
                                  "time composed from measured per-query/per-track/per-pair costs (`--impl reference` runs the whole job)",
                        "parts_us": {k: val * 1e6 for k, val in parts.items()}}
    if rank == 0:
        line = {
            "metric": METRIC, "value": nq_total / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "extra_untimed_warmup_steps": int(1.5 / 7.5e-3),
            "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": f"f16 operands ({eng.precision} precision), f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else
                       (f"{nq_total} synthetic query videos ({nq} per GPU) x {nm} music tracks sharded over {world} GPUs" if weak
                        else WORKLOAD + f", gallery and queries sharded over {world} GPUs"),
                       "n_queries": nq_total, "n_tracks": nm, "top_k": TOPK, "precision": eng.precision,
                       "parallelism": f"gallery-shard x{world}" if world > 1 else "single GPU",
                       "l2": "inputs (1.39 GB of features per 2000 x 4000 job) exceed the 126 MB L2; no explicit flush",
                       "steps_overlap": "no synchronisation between steps at N = 1; at N > 1 every step ends with a "
                                        "device synchronisation (the NCCL stream must not fall behind)"},
            "e2e": e2e, "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if rtd is not None:
            line["retrieve_then_detect"] = rtd
        if strong is not None:
            line["strong_same_job"] = strong
        if sharded_parity is not None:
            line["sharded_parity"] = sharded_parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
