"""CTA-pair (cta_group::2) form of the tcgen05 GEMM against the single-CTA form: same bits, time per launch. Diagnostics."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def timed(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / reps


cases = [(56000, 256, 256, 2, True), (56000, 768, 256, 2, False), (56000, 256, 768, 2, True), (56000, 256, 256, 1, True),
         (56000, 256, 256, 0, False), (160000, 256, 256, 2, True), (37889, 768, 256, 2, False), (23000, 256, 512, 2, True),
         (56000, 256, 256, 1, "ln_res"), (160000, 256, 256, 2, "ln_res"), (148 * 128 * 16, 256, 256, 2, True),
         (148 * 128 * 16, 256, 256, 0, False)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for M, N, K, split, pair_out in cases:
    a32 = torch.randn(M, K, generator=g)
    w32 = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g).to(dev)
    if split == 0:
        a, w = a32.to(torch.float16).to(dev), w32.to(torch.float16).to(dev)
        f = lambda: ops.gemm_f16(a, w, bias=bias, out_dtype=torch.float32)
    else:
        a = ops.split_pair(a32).to(dev) if split == 2 else a32.to(torch.float16).to(dev)
        w = ops.split_pair(w32).to(dev)
        if pair_out == "ln_res":      # out_proj-style epilogue: pair residual + LayerNorm, pair out
            rp = ops.split_pair(torch.randn(M, N, generator=g)).to(dev)
            gam, bet = torch.randn(N, generator=g).to(dev), torch.randn(N, generator=g).to(dev)
            f = lambda: ops.gemm_f16_split(a, w, split, bias=bias, residual_pair=rp, ln=(gam, bet), out_pair=True)
        else:
            f = (lambda: ops.gemm_f16_split(a, w, split, bias=bias, out_pair=True)) if pair_out else \
                (lambda: ops.gemm_f16_split_h(a, w, split, bias=bias, act=0))
    res = {}
    for mode in ("0", "1"):
        os.environ["MADE_GEMM_PAIR"] = mode
        out = f()
        torch.cuda.synchronize()
        res[mode] = (out.clone(), timed(f))
    same = torch.equal(res["0"][0], res["1"][0])
    diff = (res["0"][0].float() - res["1"][0].float()).abs().max().item()
    tiles = ((M + 127) // 128) * (N // 256) / 148
    print(f"M={M} N={N} K={K} split={split} {pair_out}: single {res['0'][1]:7.1f} us  pair {res['1'][1]:7.1f} us  "
          f"({res['0'][1] / tiles:5.2f} -> {res['1'][1] / tiles:5.2f} us per tile per CTA)  bits equal: {same}  max diff {diff:.3g}", flush=True)
