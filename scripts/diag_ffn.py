"""Ablation timings of the fused FFN kernel (MADE_FFN_DEBUG bits) at the music-chunk size. Diagnostics."""
# needs a diagnostics build: MADE_DIAG=1 python -m mgsv_b200.build --force  (rebuild without MADE_DIAG afterwards)
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 56576
g = torch.Generator().manual_seed(0)
x = torch.randn(M, 512, generator=g).to(torch.float16).to(dev)
w1 = (torch.randn(1024, 256, generator=g) / 16).to(torch.float16).to(dev)
w2 = (torch.randn(256, 1024, generator=g) / 32).to(torch.float16).to(dev)
b1, b2 = torch.zeros(1024, device=dev), torch.zeros(256, device=dev)
gam, bet = torch.ones(256, device=dev), torch.zeros(256, device=dev)
for act, ln in ((1, False), (2, True)):
    for dbg in (0, 1, 2, 4, 3, 7):
        os.environ["MADE_FFN_DEBUG"] = str(dbg)
        f = lambda: ops.ffn_fused(x, w1, b1, w2, b2, act, residual=x, ln=(gam, bet) if ln else None, pair=True)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        tiles = (M + 127) // 128
        print(f"act={act} ln={ln} debug={dbg} (1=no weight loads 2=no out epilogue 4=no act): {us:7.1f} us  "
              f"{us / -(-tiles // 148):6.1f} us per tile-wave, tensor-bound {8.62 * -(-tiles // 148):.1f} us", flush=True)
os.environ["MADE_FFN_DEBUG"] = "0"
