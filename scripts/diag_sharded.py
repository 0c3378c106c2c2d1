"""Wall-clock phase timing of the sharded evaluator (run under torchrun). Test infrastructure."""
import os, sys, time
import torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import synth
from mgsv_b200.engine import Engine
from mgsv_b200.parallel import ShardedEvaluator, shard_bounds
from mgsv_b200.pipeline import GalleryEvaluator

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
nq, nm = 2000, 4000
q0, q1 = shard_bounds(nq, rank, world); m0, m1 = shard_bounds(nm, rank, world)
v, m, ids = synth.make_eval_set(nq, nm, synth.BASE_SEED + 2)
hv = {k: v[k][q0:q1].contiguous().pin_memory() for k in ("frame_feats", "frame_mask")}
hm = {k: m[k][m0:m1].contiguous().pin_memory() for k in ("segment_feats", "segment_mask", "gt_moment", "m_duration")}
eng = Engine(dev); eng.load_state_dict(synth.make_state_dict(0))
ev = GalleryEvaluator(eng, k=100, music_chunk=512, video_chunk=512)
sh = ShardedEvaluator(ev, rank, world)
gt = torch.arange(nq, dtype=torch.int32, device=dev)

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

for it in range(4):
    t0 = sync()
    fs, vf, fm = ev.encode_queries(hv["frame_feats"], hv["frame_mask"])
    t1 = sync()
    gal = ev.encode_gallery(hm["segment_feats"], hm["segment_mask"])
    t2 = sync()
    out = sh.run(hv, hm, gt, nq, nm, on_host=True)
    t3 = sync()
    host = ev.to_host(out)
    t4 = sync()
    dv = {k: t.to(dev) for k, t in hv.items()}; dm = {k: t.to(dev) for k, t in hm.items()}
    t5 = sync()
    out = sh.run(dv, dm, gt, nq, nm)
    t6 = sync()
    if rank == 0:
        print(f"it{it}: enc_q {1e3*(t1-t0):.1f} enc_gal {1e3*(t2-t1):.1f} full_host {1e3*(t3-t2):.1f} to_host {1e3*(t4-t3):.1f} "
              f"full_dev {1e3*(t6-t5):.1f} ms", flush=True)
dist.destroy_process_group()
