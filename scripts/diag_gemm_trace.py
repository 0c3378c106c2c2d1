"""Clock-stamp timeline of CTA 0 of one tcgen05 GEMM launch (K = N = 256): where the fixed ~20 us go. Diagnostics."""
# needs a diagnostics build: MADE_DIAG=1 python -m mgsv_b200.build --force  (rebuild without MADE_DIAG afterwards)
import ctypes as C, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(0)
K = N = 256
w32 = torch.randn(N, K, generator=g) / 16
w = w32.to(torch.float16).to(dev); wp = ops.split_pair(w32).to(dev)
bias = torch.zeros(N, device=dev)
names = ["entry", "setup done", "first TMA issue", "MMA: first stage full", "MMA: last stage of tile 0 full", "E: acc 0 full", "E: tile 0 done",
         "E: acc 1 full", "E: tile 1 done", "E: acc 2 full", "E: tile 2 done", "E: acc 3 full", "E: tile 3 done", "E: before final store wait",
         "E: after final store wait", "dealloc done"]
for k in (1, 4):
    M = 148 * 128 * k
    x32 = torch.randn(M, K, generator=g)
    x = x32.to(torch.float16).to(dev); xp = ops.split_pair(x32).to(dev)
    for name, f in (("fp16 bias", lambda: ops.gemm_f16(x, w, bias=bias)),
                    ("split=2 bias pair out", lambda: ops.gemm_f16_split(xp, wp, 2, bias=bias, out_pair=True))):
        for _ in range(3): f()
        buf = torch.zeros(48, dtype=torch.int64, device=dev)
        lib.made_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
        f(); torch.cuda.synchronize()
        lib.made_debug_gemm_trace(None)
        t = buf.cpu().tolist()
        print(f"--- tiles per CTA {k}, {name}: cycles after kernel entry")
        print("   " + "; ".join(f"{n} {t[i] - t[0]}" for i, n in enumerate(names) if t[i]))
        if t[16]:
            ev = ["chunk start", "tmem loaded", "math done", "box free", "store issued"]
            print("   tile 1 epilogue, cycles after its acc-full: " + " | ".join(
                f"chunk {j}: " + ", ".join(f"{ev[e]} {t[16 + 5 * j + e] - t[7]}" for e in range(5) if t[16 + 5 * j + e]) for j in range(4)))
