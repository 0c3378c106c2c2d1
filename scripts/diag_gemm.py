"""Where a K = 256, N = 256 GEMM of the encoder chains spends its time: epilogue variants at the music-chunk size. Diagnostics."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 56576
g = torch.Generator().manual_seed(0)
def t(f, name):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    tiles = (M + 127) // 128
    print(f"{name:70s} {us:7.1f} us   {us / -(-tiles // 148):6.1f} us per wave of 148 tiles", flush=True)
for K, N in ((256, 256), (256, 768), (768, 256)):
    x32 = torch.randn(M, K, generator=g)
    x = x32.to(torch.float16).to(dev)
    xp = ops.split_pair(x32).to(dev)
    w32 = torch.randn(N, K, generator=g) / 16
    w = w32.to(torch.float16).to(dev)
    wp = ops.split_pair(w32).to(dev)
    bias = torch.zeros(N, device=dev)
    res32 = torch.randn(M, N, generator=g)
    res = res32.to(dev)
    resp = ops.split_pair(res32).to(dev)
    gam, bet = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    print(f"--- M={M} K={K} N={N}")
    t(lambda: ops.gemm_f16(x, w, bias=bias), "fp16: bias only, fp16 out")
    if N == 256:
        t(lambda: ops.gemm_f16(x, w, bias=bias, ln=(gam, bet)), "fp16: bias + LN, fp16 out")
        t(lambda: ops.gemm_f16(x, w, bias=bias, residual=res), "fp16: bias + fp32 residual, fp16 out")
        t(lambda: ops.gemm_f16(x, w, bias=bias, residual=res, ln=(gam, bet)), "fp16: bias + fp32 residual + LN, fp16 out")
    for ws in ("1", "0"):
        os.environ["MADE_GEMM_WS128"] = ws
        t(lambda: ops.gemm_f16_split_h(x, wp, 1, bias=bias), f"split=1, fp16 out, weight-stationary={ws}")
        t(lambda: ops.gemm_f16_split_h(xp, wp, 2, bias=bias), f"split=2, fp16 out, weight-stationary={ws}")
    os.environ["MADE_GEMM_WS128"] = "0"
    t(lambda: ops.gemm_f16_split(x, wp, 1, bias=bias, out_pair=True), "split=1 (W pair): bias, pair out")
    t(lambda: ops.gemm_f16_split(xp, wp, 2, bias=bias, out_pair=True), "split=2 (A and W pairs): bias, pair out")
    if N == 256:
        t(lambda: ops.gemm_f16_split(xp, wp, 2, bias=bias, residual_pair=resp, out_pair=True), "split=2: bias + pair residual, pair out")
        t(lambda: ops.gemm_f16_split(xp, wp, 2, bias=bias, residual_pair=resp, ln=(gam, bet), out_pair=True), "split=2: bias + pair residual + LN, pair out")
