"""Fixed vs per-tile cost of the tcgen05 GEMM: K = N = 256, M = 148 * 128 * k rows. Diagnostics."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
K = N = 256
w32 = torch.randn(N, K, generator=g) / 16
w = w32.to(torch.float16).to(dev); wp = ops.split_pair(w32).to(dev)
bias = torch.zeros(N, device=dev)
for k in (1, 2, 4, 8, 16):
    M = 148 * 128 * k
    x32 = torch.randn(M, K, generator=g)
    x = x32.to(torch.float16).to(dev); xp = ops.split_pair(x32).to(dev)
    for name, f in (("fp16 bias", lambda: ops.gemm_f16(x, w, bias=bias)),
                    ("split=2 bias pair out", lambda: ops.gemm_f16_split(xp, wp, 2, bias=bias, out_pair=True))):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"tiles per CTA {k:2d}  {name:24s} {us:7.1f} us  ({us / k:6.2f} us per tile per CTA)", flush=True)
