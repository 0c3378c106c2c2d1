"""Epilogue ablations (MADE_GEMM_DEBUG: 1 no bulk stores, 3 no staging writes either, 7 no TMEM loads either) of the split-2
GEMM, single-CTA and CTA-pair forms: is a tile bound by its mainloop or by its epilogue?  Needs MADE_DIAG=1 build."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def timed(f):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100


K = 256
for N, tiles_m in ((256, 16), (768, 6), (256, 4)):
    M = 148 * 128 * tiles_m
    xp = ops.split_pair(torch.randn(M, K, generator=g)).to(dev)
    x16 = xp[:, :K].contiguous()
    wp = ops.split_pair(torch.randn(N, K, generator=g) / 16).to(dev)
    w16 = wp[:, :K].contiguous()
    bias = torch.zeros(N, device=dev)
    per = tiles_m * (N // 256)
    for name, f in (("split=2 fp16 out", lambda: ops.gemm_f16_split_h(xp, wp, 2, bias=bias)),
                    ("split=2 pair out", lambda: ops.gemm_f16_split(xp, wp, 2, bias=bias, out_pair=True)),
                    ("fp16 plain fp16 out", lambda: ops.gemm_f16(x16, w16, bias=bias))):
        for pair in ("0", "1"):
            os.environ["MADE_GEMM_PAIR"] = pair
            row = []
            for dbg in ("0", "1", "3", "7"):
                os.environ["MADE_GEMM_DEBUG"] = dbg
                row.append(timed(f) / per)
            print(f"M={M} N={N} {name:20s} pair={pair}: us per tile per CTA: full {row[0]:5.2f}  no stores {row[1]:5.2f}  "
                  f"no staging {row[2]:5.2f}  no TMEM loads {row[3]:5.2f}", flush=True)
os.environ.pop("MADE_GEMM_DEBUG")
