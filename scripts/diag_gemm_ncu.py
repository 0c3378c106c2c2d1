import os, sys, torch
sys.path.insert(0, "/root/repo")
from mgsv_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M, N, K = 148 * 128 * 8, 256, 256
a32 = torch.randn(M, K, generator=g); w32 = torch.randn(N, K, generator=g) / 16
bias = torch.randn(N, generator=g).to(dev)
a = ops.split_pair(a32).to(dev); w = ops.split_pair(w32).to(dev)
a16 = a32.to(torch.float16).to(dev); w16 = w32.to(torch.float16).to(dev)
for mode in ("0", "1"):
    os.environ["MADE_GEMM_PAIR"] = mode
    for _ in range(2):
        ops.gemm_f16_split(a, w, 2, bias=bias, out_pair=True)
        ops.gemm_f16(a16, w16, bias=bias)
torch.cuda.synchronize()
