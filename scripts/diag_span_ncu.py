"""ncu target: gIoU and matcher cost on 16384 x 16384 pairs (one launch each after a warm-up). Diagnostics."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import ops
dev = torch.device("cuda:0")
n = 16384
g = torch.Generator().manual_seed(5)
a = torch.stack([torch.rand(n, generator=g), torch.rand(n, generator=g) * 0.3 + 0.01], 1).to(dev)
b = torch.stack([torch.rand(n, generator=g), torch.rand(n, generator=g) * 0.3 + 0.01], 1).to(dev)
prob = torch.rand(n, generator=g).to(dev)
sa, sb = ops.span_cw_to_se(a), ops.span_cw_to_se(b)
for _ in range(2):
    ops.generalized_temporal_iou(sa, sb, check=False)
    ops.matcher_cost(prob, a, b)
torch.cuda.synchronize()
ops.generalized_temporal_iou(sa, sb, check=False)   # launches 5 and 6 of span_pair_kernel: the profiled ones (-s 4 -c 2)
ops.matcher_cost(prob, a, b)
torch.cuda.synchronize()
