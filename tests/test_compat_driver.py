"""The reference's own `eval_epoch` (test-MaDe.py:243-447, UNMODIFIED source) driven over the made_b200 mirrors.

CPU container only: needs /root/reference (absent on the GPU box → skipped there).  `mgsv_b200.compat.install()`
puts the reference module names over the mirrors, the driver is loaded from its file, and `eval_epoch` runs on a
`mgsv_b200.model.Uni_model` whose device work is replaced by oracle-backed stand-ins (this is the test of the HOST
interface: argument order, dict keys, shapes, dtypes, the `.cpu()` / `.to(device)` dance around the X-Pool view,
`Recall_metrics(sim_matrix, dedup=True, all_music_ids_list=...)` on a host float64 matrix).  Since the stand-ins
are the pinned oracle, the driver must print the reference's own metrics: compared with tests/golden/cfg1_256.npz.
"""
import importlib
import importlib.util
import logging
import os
import sys
import types

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MADE_REFERENCE", "/root/reference")

from mgsv_b200 import config, losses, metrics, ops, synth      # noqa: E402
from oracle import made_oracle as O                            # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "test-MaDe.py")),
                                reason="the reference tree is only present in the build container")


class _OracleEngine:
    """Stands in for mgsv_b200.engine.Engine: same methods, oracle arithmetic, CPU tensors."""

    def __init__(self):
        self.device = torch.device("cpu")
        self.precision = "fp32"
        self.sd = None

    def load_state_dict(self, sd):
        self.sd = {k: v.detach().clone().float() for k, v in sd.items()}

    def encode(self, modality, feats, masks, **kw):
        f = O.encode_video if modality == 0 else O.encode_music
        seq, pooled = f(self.sd, feats.float(), masks.float())
        return seq, seq, pooled

    def detr_detect(self, frame16, frame_masks, seg16, seg_masks, video_feats, want_proj=False, track_idx=None):
        src = torch.cat([frame16, seg16], 1)
        mask = torch.cat([frame_masks, seg_masks], 1).float()
        hs, _ = O.detr_forward(self.sd, src, mask, O.position_embedding_sine(mask), video_feats.unsqueeze(1))
        out = O.calc_output(self.sd, hs, frame16)
        lay = out["aux_outputs"] + [out]
        return dict(pred_logits=torch.stack([a["pred_logits"][:, 0] for a in lay]),
                    pred_spans=torch.stack([a["pred_spans"][:, 0] for a in lay]),
                    proj_queries=torch.stack([a["proj_queries"][:, 0] for a in lay]),
                    proj_vid_mem=out["proj_vid_mem"])

    def gallery_prepare(self, seg16, masks, out=None):
        return seg16.float(), masks.float(), None

    def query_prepare(self, video_feats):
        return video_feats, None

    def xpool_score(self, q, vhat, kz, gram, bits, out=None, col_offset=0):
        return O.sim_matrix_music_pooling(q, O.xpool(self.sd, q, kz, gram))

    def xpool_pooled(self, video_feats, segment_feats, segment_masks, out=None, track_chunk=None, which=1):
        return O.xpool(self.sd, video_feats.float(), segment_feats.float(), segment_masks.float())


def _cpu_rank_topk(single, dual, gt_col=None, prev_same=None, k=0, **kw):
    tot = single.double() + (0 if dual is None else dual.double())
    n = tot.shape[0]
    order = torch.sort(-tot, dim=1, stable=True).indices
    rank = torch.zeros(n, dtype=torch.int32)
    for i in range(n):
        cols = [int(gt_col[i])]
        while prev_same is not None and int(prev_same[cols[-1]]) >= 0:
            cols.append(int(prev_same[cols[-1]]))
        s = tot[i, cols].max()
        beat = (tot[i] > s).nonzero().flatten().tolist()
        if prev_same is None:
            rank[i] = len(beat)
        else:       # distinct ids: a column counts iff no earlier column of its id also beats s
            roots = set()
            for c in beat:
                r = c
                while int(prev_same[r]) >= 0:
                    r = int(prev_same[r])
                roots.add(r)
            rank[i] = len(roots)
    return dict(rank=rank, topk_idx=order[:, :k].to(torch.int32), topk_score=torch.gather(tot, 1, order[:, :k]),
                gt_score=None)


def _cpu_detr_losses(pred_logits, pred_spans, proj_queries, proj_vid_mem, targets_cw, empty_weight, temperature=0.07):
    rows = []
    tg = targets_cw.reshape(-1, 1, 2)
    for l in range(pred_logits.shape[0]):
        d = O.criterion_losses(dict(pred_logits=pred_logits[l][:, None], pred_spans=pred_spans[l][:, None],
                                    proj_queries=proj_queries[l][:, None], proj_vid_mem=proj_vid_mem), tg,
                               empty_weight, temperature)
        rows.append(torch.stack([d[k].float() for k in losses._NAMES]))
    return torch.stack(rows)


def _load_driver():
    """test-MaDe.py initialises NCCL at import (:25): give it gloo, load it from its file, from its directory."""
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, *a, **k: None if dist.is_initialized() else real_init("gloo", *a, **k)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        spec = importlib.util.spec_from_file_location("test_made_driver_compat", os.path.join(REF, "test-MaDe.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
        dist.init_process_group = real_init
    mod.logger = logging.getLogger("compat.driver")
    return mod


_REF_NAMES = ("model", "modules", "utils", "music_detr", "dataloaders")


@pytest.fixture
def compat_env(monkeypatch):
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_NAMES}
    for k in saved_mods:
        del sys.modules[k]
    for name in ["clip", "wget", "timm", "timm.models", "timm.models.layers"]:     # dataloaders import them unused
        monkeypatch.setitem(sys.modules, name, sys.modules.get(name) or types.ModuleType(name))
    # device work → oracle stand-ins (BEFORE the compat modules bind the names)
    monkeypatch.setattr(ops, "_to_cuda", lambda t: t)
    monkeypatch.setattr(ops, "_default_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(ops, "rank_topk", _cpu_rank_topk)
    monkeypatch.setattr(ops, "cal_distance", lambda x, y, distance_type="COS", out=None, col_offset=0: (
        O.cal_distance_cos(torch.as_tensor(x), torch.as_tensor(y)).numpy().astype(np.float64)
        if isinstance(x, np.ndarray) else O.cal_distance_cos(x, y)))
    monkeypatch.setattr(ops, "sim_matrix_music_pooling", lambda v, p, out=None, col_offset=0: O.sim_matrix_music_pooling(v, p))
    monkeypatch.setattr(ops, "span_cw_to_se", lambda cw: O.span_cw_to_se(cw))
    monkeypatch.setattr(ops, "span_iou", lambda st, ed, gt, md, mx=240.0: O.detr_iou(st, ed, gt.reshape(-1, 1, 2), md))
    monkeypatch.setattr(losses, "detr_losses", _cpu_detr_losses)
    monkeypatch.setattr(losses, "retrieval_loss", lambda dual, single, ls: (
        O.info_nce_loss(dual, torch.tensor(ls)) + O.clip_loss(single, torch.tensor(ls))))
    from mgsv_b200 import compat
    compat.install()
    sys.path.insert(1, REF)
    yield compat
    sys.path[:] = saved_path
    for k in [k for k in sys.modules if k.split(".")[0] in _REF_NAMES]:
        del sys.modules[k]
    sys.modules.update(saved_mods)


def test_reference_eval_epoch_runs_unmodified_over_the_mirrors(compat_env, golden_dir):
    drv = _load_driver()
    from mgsv_b200.model import Uni_model
    # the driver's names now resolve to the mirrors; what is not replaced still comes from the reference
    assert drv.Uni_model is Uni_model
    assert drv.Recall_metrics is metrics.Recall_metrics_matrix and drv.IoU_metrics is metrics.IoU_metrics
    assert drv.Composite_metrics is metrics.Composite_metrics and drv.calc_similarity is metrics.calc_similarity
    assert sys.modules["modules.metrics"].__file__.startswith(compat_env.COMPAT_DIR)
    assert sys.modules["utils.util_train"].__file__.startswith(REF)
    assert drv.CLIPLoss.__module__ == "modules.loss" and callable(drv.InfoNCELoss) and callable(drv.cal_distance)

    # configs[0] exactly as oracle/gen_golden.py:gen_cfg1 fed it to the reference model
    from oracle.gen_golden import dup_tracks
    N, bs = 256, 32
    v, m, ids = synth.make_eval_set(N, N, synth.BASE_SEED + 1)
    dup_tracks(m, ids, n_dup=16)
    batches = []
    for s in range(0, N, bs):
        e = s + bs
        batches.append((dict(frame_feats=v["frame_feats"][s:e].clone(), frame_mask=v["frame_mask"][s:e],
                             segment_feats=m["segment_feats"][s:e].clone(), segment_mask=m["segment_mask"][s:e]),
                        dict(video_id=ids["video_ids"][s:e], music_id=ids["music_ids"][s:e], gt_moment=m["gt_moment"][s:e],
                             m_duration=m["m_duration"][s:e], v_duration=v["v_duration"][s:e]),
                        m["spans_target"][s:e]))
    args = config.default_args(name="compat")
    model = Uni_model(args, torch.device("cpu"), None)
    model.load_state_dict(synth.make_state_dict(0))
    model._engine = _OracleEngine()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    loss_avg, ret_m, loc_m, com_m = drv.eval_epoch(1, args, model, batches, torch.device("cpu"))

    g = np.load(os.path.join(golden_dir, "cfg1_256.npz"))
    for name, got in (("ret", ret_m), ("loc", loc_m), ("com", com_m)):
        for k, want in zip(g[f"{name}_keys"], g[f"{name}_vals"]):
            assert abs(float(got[str(k)]) - float(want)) <= 1e-4 * max(1.0, abs(float(want))), (name, k, got[str(k)], want)
    assert [int(c) for c in ret_m["cols"]] == [int(c) for c in g["ind"]]
    assert abs(float(loss_avg) - float(g["driver_loss_avg"])) < 1e-3


def test_xpool_view_contract(compat_env):
    from mgsv_b200.model import Uni_model
    model = Uni_model(config.default_args(), torch.device("cpu"), None)
    model._engine = _OracleEngine()
    view = model.video_guided_to_music_pooling_cross_transformer
    assert view.cpu() is view and view.to(torch.device("cpu")) is view          # test-MaDe.py:392/395
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    vid, seg, msk = torch.randn(5, 256), torch.randn(3, 96, 256), torch.ones(3, 96)
    msk[1, 40:] = 0
    out = view(vid, seg, msk)
    assert out.shape == (3, 5, 256)
    assert torch.equal(out, O.xpool(model._engine.sd, vid, seg, msk))
    assert all(torch.equal(v, sd_before[k]) for k, v in model.state_dict().items())
    with pytest.raises(ValueError):
        view(vid, seg, None)                                                    # fusion_mask=0 is not shipped
    with pytest.raises(ValueError):
        view(vid, seg[:, :50], msk[:, :50])
    os.environ["MADE_POOLED_CAP_GB"] = "0.000001"
    try:
        with pytest.raises(RuntimeError, match="score_gallery"):
            view(vid, seg, msk)
    finally:
        del os.environ["MADE_POOLED_CAP_GB"]


def test_recall_metrics_matrix_signature_and_dedup(compat_env):
    """utils/util_test.py:32-97 called the reference's way: host float64 matrix, dedup flag, id list."""
    rng = np.random.default_rng(5)
    single = rng.standard_normal((40, 40)).astype(np.float32)
    dual = rng.standard_normal((40, 40)).astype(np.float32)
    sim = single.astype(np.float64) * 1.0 + dual.astype(np.float64) * 1.0      # test-MaDe.py:403
    ids = [f"m{i % 31}" for i in range(40)]                                    # repeated tracks
    sim[:, 31:] = sim[:, :9]                                                   # a repeated id has identical columns
    got_m, got_ind, got_res = metrics.Recall_metrics_matrix(sim, dedup=True, all_music_ids_list=ids)
    want_m, want_ind, want_res = O.recall_metrics(sim, ids)
    assert list(got_ind) == list(want_ind)
    assert all(abs(float(got_m[k]) - float(want_m[k])) < 1e-9 for k in want_m if k != "cols")
    assert [r["rank"] for r in got_res] == [int(i) + 1 for i in want_ind]
    assert all(r["music_id"] == ids[i] for i, r in enumerate(got_res))
