"""Clock-stamp trace of one CTA of the fused X-Pool kernel: where a track's ~3 us go.  Diagnostics."""
# needs a diagnostics build: MADE_DIAG=1 python -m mgsv_b200.build --force  (rebuild without MADE_DIAG afterwards)
import ctypes as C, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mgsv_b200 import _lib, synth
from mgsv_b200.engine import Engine
dev = torch.device("cuda:0")
eng = Engine(dev); eng.load_state_dict(synth.make_state_dict(0))
nq, nm = 2000, 1000
v, m, _ = synth.make_eval_set(nq, nq, synth.BASE_SEED + 2)
m = {k: t[:nm] for k, t in m.items() if isinstance(t, torch.Tensor)}
seq_m, _, _ = eng.encode(_lib.MUSIC, m["segment_feats"].to(dev), m["segment_mask"].to(dev), want_f32=False)
_, _, pooled_v = eng.encode(_lib.VIDEO, v["frame_feats"].to(dev), v["frame_mask"].to(dev), want_f32=False)
q, vhat = eng.query_prepare(pooled_v)
kz, gram, bits = eng.gallery_prepare(seq_m, m["segment_mask"].to(dev))
lib = _lib.load()
names = {0: "P k_empty ok -> issue K", 1: "P zg_empty ok -> issue ZG", 2: "M k_full", 3: "M t_free", 4: "M p_full", 5: "M zg_full",
         6: "M y_free", 8: "E s_full", 9: "E max done", 10: "E bar1", 11: "E p arrive", 12: "E t_full", 13: "E t_free arrive",
         14: "E y_full", 15: "E y_free arrive", 16: "E end"}
for dbg in [int(a) for a in (sys.argv[1:] or ["0", "63"])]:
    os.environ["MADE_XPOOL_DEBUG"] = str(dbg)
    buf = torch.zeros(64 * 24, dtype=torch.int64, device=dev)
    for _ in range(2): eng.xpool_score(q, vhat, kz, gram, bits)
    lib.made_debug_xpool_trace(C.c_void_p(buf.data_ptr()))
    eng.xpool_score(q, vhat, kz, gram, bits)
    torch.cuda.synchronize()
    lib.made_debug_xpool_trace(None)
    t = buf.cpu().view(64, 24)
    nb = [(int(x) + 15) // 16 for x in m["segment_mask"].sum(1)[0:1000:9][:64]]   # CTA 0: slice 0 of 9 (tracks 0, 9, 18, ...)
    print(f"debug={dbg}: per-track period (E end to E end), tracks 20..40: "
          f"{[(int(t[u, 16] - t[u - 1, 16])) for u in range(20, 40)]}")
    for u in (24, 25, 26):
        base = int(t[u - 1, 16])
        ev = sorted((int(t[u, e]) - base, names[e]) for e in names)
        print(f"  track {u} (blocks {nb[u]}), cycles after the end of track {u - 1}: " + "; ".join(f"{n} {c}" for c, n in ev))
os.environ["MADE_XPOOL_DEBUG"] = "0"
