"""Clock-stamp timeline of CTA 0: N = 768 split-2 GEMM with fp16 output (the in_proj / X-Pool operand shape), single-CTA
and CTA-pair forms.  Needs a diagnostics build (MADE_DIAG=1)."""
import ctypes as C, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(0)
K, N = 256, 768
w32 = torch.randn(N, K, generator=g) / 16
wp = ops.split_pair(w32).to(dev)
bias = torch.zeros(N, device=dev)
names = ["entry", "setup done", "first TMA issue", "MMA: first stage full", "MMA: last stage of tile 0 full", "E: acc 0 full", "E: tile 0 done",
         "E: acc 1 full", "E: tile 1 done", "E: acc 2 full", "E: tile 2 done", "E: acc 3 full", "E: tile 3 done", "E: before final store wait",
         "E: after final store wait", "dealloc done"]
M = 148 * 128 * 3
xp = ops.split_pair(torch.randn(M, K, generator=g)).to(dev)
for mode in ("0", "1"):
    os.environ["MADE_GEMM_PAIR"] = mode
    f = lambda: ops.gemm_f16_split_h(xp, wp, 2, bias=bias, act=0)
    for _ in range(3): f()
    buf = torch.zeros(48, dtype=torch.int64, device=dev)
    lib.made_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
    f(); torch.cuda.synchronize()
    lib.made_debug_gemm_trace(None)
    t = buf.cpu().tolist()
    print(f"--- pair={mode}: 9 tiles per CTA, N=768 split=2 fp16 out: cycles after kernel entry")
    print("   " + "; ".join(f"{n} {t[i] - t[0]}" for i, n in enumerate(names) if t[i]))
    if t[16]:
        ev = ["chunk start", "tmem loaded", "math done", "box free", "store issued"]
        print("   tile 1 epilogue, cycles after its acc-full: " + " | ".join(
            f"chunk {j}: " + ", ".join(f"{ev[e]} {t[16 + 5 * j + e] - t[7]}" for e in range(5) if t[16 + 5 * j + e]) for j in range(4)))
