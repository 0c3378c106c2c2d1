"""ncu target: rank + top-100 of 8192 x 16384 and of 2000 x 4000 (one launch each after a warm-up). Diagnostics."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
for nq, nm in ((8192, 16384), (2000, 4000)):
    single = torch.randn(nq, nm, device=dev); dual = torch.randn(nq, nm, device=dev)
    gt = torch.randint(0, nm, (nq,), device=dev, dtype=torch.int32)
    for _ in range(2):
        ops.rank_topk(single, dual, gt, None, k=100)
torch.cuda.synchronize()
