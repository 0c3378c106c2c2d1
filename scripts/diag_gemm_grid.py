"""Does the time of a split-2 GEMM launch depend on how many CTAs share the chip?  (MADE_GEMM_GRID, diagnostics build)"""
import os, sys, torch
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
for N, M in ((256, 148 * 128 * 16), (768, 148 * 128 * 6), (256, 148 * 128 * 3)):
    xp = ops.split_pair(torch.randn(M, K, generator=g)).to(dev)
    wp = ops.split_pair(torch.randn(N, K, generator=g) / 16).to(dev)
    bias = torch.zeros(N, device=dev)
    tiles = (M // 128) * (N // 256)
    for dbg in ("0", "7"):
        os.environ["MADE_GEMM_DEBUG"] = dbg
        row = []
        for grid in (148, 111, 74, 37):
            os.environ["MADE_GEMM_GRID"] = str(grid)
            us = timed(lambda: ops.gemm_f16_split_h(xp, wp, 2, bias=bias))
            row.append(f"grid {grid}: {us:7.1f} us ({us * grid / tiles:5.2f} us per tile per CTA)")
        print(f"M={M} N={N} split=2 fp16 out, {'full' if dbg == '0' else 'mainloop only'}: " + "  ".join(row), flush=True)
