import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import _lib, ops, synth
from mgsv_b200 import config as cfg
from mgsv_b200.engine import Engine
from mgsv_b200.index import GalleryIndex
from mgsv_b200.pipeline import GalleryEvaluator
dev = torch.device("cuda:0")
eng = Engine(dev); eng.load_state_dict(synth.make_state_dict(0))
ev = GalleryEvaluator(eng, k=100)
nq, nm = 2000, 4096
v = synth.make_videos(nq, synth.BASE_SEED + 2)
m = synth.make_tracks(nm, 5)
_, vf, _ = ev.encode_queries(v["frame_feats"].to(dev), v["frame_mask"].to(dev))
idx = GalleryIndex(ev, capacity=nm)
for s in range(0, nm, 1000):
    idx.add(m["segment_feats"][s:s+1000].to(dev), m["segment_mask"][s:s+1000].to(dev))
gt = (torch.arange(nq) * 7919) % nm
qprep = eng.query_prepare(vf)
single = torch.empty((nq, nm), device=dev); dual = torch.empty((nq, nm), device=dev)
L = cfg.L_M
eng.xpool_score(qprep[0], qprep[1], idx.gal["kz"], idx.gal["gram"], idx.gal["bits"], out=single)
ops.cal_distance(vf, idx.gal["pooled"], out=dual)
gs = idx.gt_scores(vf, gt, qprep)
tot = single.double() + dual.double()
ref = tot.gather(1, gt.to(dev)[:, None]).squeeze(1)
print("gt_scores mismatches:", int((gs != ref).sum()), "max diff", float((gs - ref).abs().max()))
# separately
for pc in (64, 128, 256):
    qi = torch.arange(pc, device=dev)
    sub = idx._subgallery(gt.to(dev)[qi])
    s2 = eng.xpool_score(qprep[0][qi].contiguous(), qprep[1][qi].contiguous(), sub["kz"], sub["gram"], sub["bits"])
    d2 = ops.cal_distance(vf[qi].contiguous(), sub["pooled"])
    rs = single[qi][:, gt.to(dev)[qi]]
    rd = dual[qi][:, gt.to(dev)[qi]]
    print(pc, "single mismatch", int((s2 != rs).sum()), float((s2 - rs).abs().max()), "dual mismatch", int((d2 != rd).sum()), float((d2 - rd).abs().max()))
