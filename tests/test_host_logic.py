"""CPU tests: host logic, ABI surface, metric bookkeeping (no CUDA compute calls)."""
import ctypes
import json
import os
import re
import sys

import numpy as np
import pytest
import torch

from mgsv_b200 import _lib, config, metrics, ops, synth
from mgsv_b200.parallel import owner_of, shard_bounds
from oracle import build_c
from oracle import made_oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_check_args_accepts_shipped_and_rejects_other_branches():
    config.check_args(config.default_args())
    for k, ok in (("vmr_fusion", "XA-music-video"), ("mml_fusion", "CA")):     # the two built variants beside the shipped config
        config.check_args(config.default_args(**{k: ok}))
    for k, bad in (("vmr_fusion", "XA-video"), ("mml_fusion", "add"), ("detr_dec_layers", 2),
                   ("agg_module", "mlp"), ("fusion_mask", 0), ("num_moment_queries", 5)):
        with pytest.raises(ValueError):
            config.check_args(config.default_args(**{k: bad}))


def test_state_dict_spec_matches_reference_inventory():
    spec = synth.state_dict_spec()
    keys = [k for k, _, _ in spec]
    assert len(keys) == len(set(keys)) == 199          # 199 entries
    n_param = sum(int(np.prod(s)) if s else 1 for k, s, kind in spec if kind not in ("pe", "empty_weight"))
    assert n_param == 10_534_917
    sd = synth.make_state_dict(0)
    sd2 = synth.make_state_dict(0)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert torch.allclose(sd["logit_scale"], torch.tensor(np.log(1 / 0.03), dtype=torch.float32))


def test_synthetic_inputs_shape_and_masks():
    v, m, ids = synth.make_eval_set(32, 64, 5)
    assert v["frame_feats"].shape == (32, 50, 512) and m["segment_feats"].shape == (64, 96, 768)
    assert (v["frame_mask"].sum(1) >= 6).all() and (m["segment_mask"].sum(1) >= 13).all()
    # padded rows zeroed, masks are prefixes
    assert (v["frame_feats"][v["frame_mask"] == 0] == 0).all()
    assert ((m["segment_mask"][:, 1:] <= m["segment_mask"][:, :-1])).all()
    assert (m["gt_moment"][:, 0, 1] <= m["m_duration"]).all()
    v2, _, _ = synth.make_eval_set(32, 64, 5)
    assert torch.equal(v["frame_feats"], v2["frame_feats"])


def test_c_oracle_matches_torch_oracle_bitwise():
    a, b, logits = synth.make_span_pairs(300, 200, 21)
    g_c = build_c.giou(build_c.cw_to_se(a.numpy()), build_c.cw_to_se(b.numpy()))
    g_t = O.generalized_temporal_iou(O.span_cw_to_se(a), O.span_cw_to_se(b)).numpy()
    assert np.array_equal(g_c, g_t, equal_nan=True)
    prob = logits.softmax(-1)[:, 0]
    tgt = b[b[:, 1] != 0]
    assert np.array_equal(build_c.matcher_cost(prob.numpy(), a.numpy(), tgt.numpy()),
                          O.matcher_cost(prob, a, tgt).numpy(), equal_nan=True)
    st, ed = torch.rand(50) * 200 - 10, torch.rand(50) * 260
    gt = torch.sort(torch.rand(50, 1, 2) * 240, dim=-1)[0]
    md = torch.rand(50) * 200 + 40
    assert np.array_equal(build_c.detr_iou(st.numpy(), ed.numpy(), gt.numpy(), md.numpy()),
                          O.detr_iou(st, ed, gt, md).numpy())


def test_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "made_b200.h")).read()
    declared = set(re.findall(r"\b(made_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/made_b200.h but not exported"
    assert set(_lib.SIGNATURES) | {"made_last_error_string", "made_ragged_index_words"} == declared
    assert lib.made_abi_version() == 2


def test_product_path_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mgsv_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()
    with pytest.raises(RuntimeError):
        ops.span_cw_to_se(torch.zeros(4, 2))


def test_uni_model_interface_on_cpu():
    from mgsv_b200.model import Uni_model
    model = Uni_model(config.default_args(), torch.device("cpu"), None)
    sd = model.state_dict()
    ref = synth.make_state_dict(0)
    assert list(sd.keys()) != [] and set(sd.keys()) == set(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in ref)
    assert model.criterion.foreground_label == 0
    assert len(model.criterion.weight_dict) == 24
    assert hasattr(model, "video_guided_to_music_pooling_cross_transformer")
    assert sum(p.numel() for p in model.parameters()) == 10_534_917
    groups = model.get_temporal_parameter() + model.get_matching_parameter() + model.get_detection_parameter()
    # every parameter except decoder_query_embed (absent from the reference's groups too) is in a group
    assert sum(p.numel() for p in groups) == 10_534_917 - 256
    model.load_state_dict(synth.make_state_dict(3))
    model.eval().float()
    with pytest.raises(ValueError):
        Uni_model(config.default_args(vmr_loss="dual"), torch.device("cpu"), None)
    # vmr_fusion "XA-music-video": the second Transformer_XA's 16 tensors join the state_dict (model_Uni.py:27-28)
    m2 = Uni_model(config.default_args(vmr_fusion="XA-music-video"), torch.device("cpu"), None)
    extra = set(m2.state_dict()) - set(sd)
    assert len(extra) == 16 and all(k.startswith("music_guided_to_video_pooling_cross_transformer.") for k in extra)
    assert sum(p.numel() for p in m2.parameters()) == 10_534_917 + 330_496
    with pytest.raises(ValueError):
        Uni_model(config.default_args(vmr_fusion="XA-video"), torch.device("cpu"), None)


def test_dedup_tables_and_rank_summary():
    ids = ["a", "b", "a", "c", "b"]
    prev, gt_col, has = ops.dedup_tables(ids)
    assert list(prev) == [-1, -1, 0, -1, 1] and list(gt_col) == [2, 4, 2, 3, 4] and has
    ind = np.array([0, 4, 11, 0, 99, 150])
    assert metrics.summarize_ranks(ind) == O.summarize_ranks(ind)
    iou = np.array([0.0, 0.31, 0.5, 0.71, 0.9, 0.3], dtype=np.float32)
    a, b = metrics.IoU_metrics(list(iou)), O.iou_metrics(list(iou))
    assert all(abs(a[k] - b[k]) < 1e-9 for k in a)
    c1, c2 = metrics.Composite_metrics(ind, iou), O.composite_metrics(list(ind), list(iou))
    assert list(c1.keys()) == list(c2.keys())
    assert all(abs(c1[k] - c2[k]) < 1e-6 for k in c1)


def test_shard_bounds_partition():
    for n in (4000, 4001, 7, 2000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            cols = torch.arange(n)
            own = owner_of(cols, n, w)
            for r in range(w):
                assert (own[b[r][0]:b[r][1]] == r).all()


def test_ingest_chunk_bounds_partition_every_size():
    """Host inputs are ingested in tapered chunks (small first chunks so the copy engine starts early, small last chunks
    so little work is left after the last copy): whatever the taper, the chunks tile [0, n) in order without gaps."""
    from mgsv_b200.pipeline import GalleryEvaluator
    for n in (0, 1, 31, 64, 65, 499, 500, 999, 1000, 2000, 4000, 4001):
        for chunk in (32, 500, 1024, 2000):
            for head in (False, True):
                for tail in (False, True):
                    b = GalleryEvaluator._chunk_bounds(n, chunk, head, tail)
                    assert (b == []) == (n == 0)
                    if b:
                        assert b[0][0] == 0 and b[-1][1] == n
                        assert all(e > s for s, e in b) and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
                        assert max(e - s for s, e in b) <= chunk
    assert GalleryEvaluator._chunk_bounds(4000, 500, False, False) == [(s, s + 500) for s in range(0, 4000, 500)]
    assert GalleryEvaluator._chunk_bounds(4000, 500, False, True)[-3:] == [(3500, 3750), (3750, 3875), (3875, 4000)]
    assert GalleryEvaluator._chunk_bounds(2000, 500, True, False)[:3] == [(0, 62), (62, 187), (187, 437)]


def test_feature_store_reads_reference_file_layout(tmp_path):
    """dataloaders/dataloader_MGSV_EC_feature.py:46-75: {id}.pt files + CSV -> one batch schema."""
    import pandas as pd
    from mgsv_b200.ingest import FeatureStore, get_cw_proportion
    g = torch.Generator().manual_seed(3)
    fr, mu = tmp_path / "frames", tmp_path / "music"
    for d in (fr / "vit_feature", fr / "vit_mask", mu / "ast_feature", mu / "ast_mask"):
        d.mkdir(parents=True)
    vids, mids = ["v1", "v2", "v3", "v4"], ["m1", "m2", "m1", "m3"]       # m1 is used by two videos
    vfeat, mfeat = {}, {}
    for i, v in enumerate(vids):
        n = 5 + 3 * i
        f = torch.zeros(50, 512); f[:n] = torch.randn(n, 512, generator=g)
        m = torch.zeros(50); m[:n] = 1
        torch.save(f, fr / "vit_feature" / f"{v}.pt"); torch.save(m, fr / "vit_mask" / f"{v}.pt")
        vfeat[v] = (f, m)
    for i, mname in enumerate(sorted(set(mids))):
        n = 20 + 10 * i
        f = torch.zeros(96, 768); f[:n] = torch.randn(n, 768, generator=g)
        m = torch.zeros(96); m[:n] = 1
        torch.save(f, mu / "ast_feature" / f"{mname}.pt"); torch.save(m, mu / "ast_mask" / f"{mname}.pt")
        mfeat[mname] = (f, m)
    df = pd.DataFrame(dict(video_id=vids, music_id=mids, video_start=[0, 1, 2, 3], video_end=[10, 12, 14, 16],
                           music_start=[5.0, 30.0, 50.0, 100.0], music_end=[15.0, 41.0, 62.0, 250.0],
                           music_total_duration=[120.0, 200.0, 120.0, 260.0]))
    csv = tmp_path / "test.csv"
    df.to_csv(csv, index=False)
    st = FeatureStore.from_csv(str(csv), str(fr), str(mu), pin=False)
    assert st.videos["frame_feats"].shape == (4, 50, 512) and st.tracks["segment_feats"].shape == (4, 96, 768)
    for i, v in enumerate(vids):
        assert torch.equal(st.videos["frame_feats"][i], vfeat[v][0]) and torch.equal(st.videos["frame_mask"][i], vfeat[v][1])
    for i, mname in enumerate(mids):
        assert torch.equal(st.tracks["segment_feats"][i], mfeat[mname][0])
    # reference gallery = one column per row; the repeated track is marked for the dedup-aware rank
    assert list(st.gt_col) == [2, 1, 2, 3] and list(st.prev_same) == [-1, -1, 0, -1]
    assert torch.allclose(st.meta["spans_target"][3, 0], torch.tensor([(100 + 240) / 2 / 240, (240 - 100) / 240]))
    assert torch.equal(st.meta["spans_target"][:, 0], get_cw_proportion(st.videos["gt_moment"][:, 0]))
    assert torch.allclose(st.videos["v_duration"], torch.tensor([10., 11., 12., 13.]))
    st2 = FeatureStore.from_csv(str(csv), str(fr), str(mu), pin=False, dedup_tracks=True)
    assert st2.tracks["segment_feats"].shape[0] == 3 and list(st2.gt_col) == [0, 1, 0, 2] and st2.prev_same is None
    with pytest.raises(FileNotFoundError):
        FeatureStore.from_csv(str(csv), str(fr), str(tmp_path / "nowhere"), pin=False)


def test_save_results_json_schema(tmp_path):
    """utils/util_test.py:202-226."""
    import json
    from mgsv_b200.ingest import save_results_json
    ret = [dict(music_id="m1", rank=3, topk_music_ids=["m9"])]
    loc = [dict(video_id="v1", music_id="m1", m_duration=120.0, gt_moment=[[5.04321, 15.98765]], pred_st=-2.0, pred_ed=250.123456)]
    path = tmp_path / "r.json"
    save_results_json(ret, loc, [torch.tensor(0.123456)], str(path))
    got = json.load(open(path))
    assert got == [dict(video_id="v1", music_id="m1", topk_mids=["m9"], gt_mid_rank=3, iou=0.1235, m_duration=120.0,
                        gt_st=5.043, gt_ed=15.988, pred_st=0, pred_ed=240)]


def test_calc_similarity_mirror_host_logic(monkeypatch):
    """utils/util_test.py:10-29: block lists in, one [val_len, val_len] numpy matrix out, float32 for tensor
    blocks / float64 for numpy blocks.  The cosine kernel is replaced by torch here (CPU suite): only the
    host logic of the mirror is under test; the GPU suite checks the same call against the oracle."""
    from mgsv_b200 import metrics as M, ops
    calls = []

    def fake_cal_distance(x, y, distance_type="COS", out=None, col_offset=0):
        calls.append((type(x), tuple(x.shape), tuple(y.shape)))
        as_np = isinstance(x, np.ndarray)
        xt, yt = torch.as_tensor(x, dtype=torch.float32), torch.as_tensor(y, dtype=torch.float32)
        r = torch.nn.functional.normalize(xt, dim=1) @ torch.nn.functional.normalize(yt, dim=1).t()
        return r.numpy().astype(np.float64) if as_np else r

    monkeypatch.setattr(ops, "cal_distance", fake_cal_distance)
    monkeypatch.setattr(ops, "_to_cuda", lambda t: t)
    g = torch.Generator().manual_seed(5)
    vb = [torch.randn(n, 256, generator=g) for n in (40, 40, 17)]
    ab = [torch.randn(n, 256, generator=g) for n in (40, 57)]
    ref = O.cal_distance_cos(torch.cat(vb), torch.cat(ab)).numpy()
    got = M.calc_similarity(vb, ab)
    assert got.dtype == np.float32 and got.shape == (97, 97) and len(calls) == 1
    np.testing.assert_allclose(got, ref, atol=1e-6)
    got64 = M.calc_similarity([b.numpy() for b in vb], [b.numpy() for b in ab])
    assert got64.dtype == np.float64 and calls[-1][0] is np.ndarray
    np.testing.assert_allclose(got64, ref, atol=1e-6)
    with pytest.raises(ValueError):
        M.calc_similarity([], ab)


def test_bench_clock_sampler_reports_only_the_timed_region():
    """bench.py's nvidia-smi sampler: samples before `mark()` (attach + warm-up) are dropped, throttle
    reasons are collected, and a region shorter than one period falls back to the warm-up samples."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class _Proc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0

    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.lines = ["1200, 1965, Not Active, Not Active, Not Active, Not Active"]      # warm-up: clocks still ramping
    s.mark()
    s.lines += ["1965, 1965, Not Active, Not Active, Not Active, Active",
                "1950, 1965, Not Active, Not Active, Not Active, Not Active", "garbage"]
    out = s.stop()
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 2
    assert out["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.lines = ["1965, 1965, Not Active, Not Active, Not Active, Not Active"]
    s.mark()
    assert s.stop()["samples"] == 1
    s = bench.ClockSampler(0)                 # nvidia-smi missing
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]


def test_hungarian_matcher_mirror_host_logic(monkeypatch):
    """music_detr/matcher.py:36-92 mirror: target filtering (width != 0), per-sample blocks of the cost
    matrix, scipy assignment on the host.  The two kernels it calls are replaced by the oracle here (CPU
    suite); the GPU suite runs the real ones (test_hungarian_matcher_mirror)."""
    import types
    from mgsv_b200 import matcher as M

    def fake_postproc(logits, spans, *a, **k):
        return None, None, logits.softmax(-1)[:, 0], None

    def fake_cost(prob_fg, out_spans, tgt_spans, cost_span, cost_giou, cost_class):
        giou = O.generalized_temporal_iou(O.span_cw_to_se(out_spans), O.span_cw_to_se(tgt_spans))
        return cost_span * torch.cdist(out_spans, tgt_spans, p=1) - cost_giou * giou - cost_class * prob_fg[:, None]

    monkeypatch.setattr(ops, "moment_postproc", fake_postproc)
    monkeypatch.setattr(ops, "matcher_cost", fake_cost)
    m = M.build_matcher(types.SimpleNamespace(span_loss_type="l1", max_snippet_num=100, fb_label="01"))
    g = torch.Generator().manual_seed(1)
    bs, nq = 6, 3
    out = dict(pred_logits=torch.randn(bs, nq, 2, generator=g), pred_spans=torch.rand(bs, nq, 2, generator=g) * 0.3 + 0.1)
    tg = torch.rand(bs, 2, 2, generator=g) * 0.3 + 0.1
    tg[2, :, 1] = 0            # sample 2: no target survives the width filter
    tg[4, 1, 1] = 0            # sample 4: one target
    res = m(out, tg)
    assert [len(r) for r, _ in res] == [2, 2, 0, 2, 1, 2]
    C, sizes = m.cost_matrix(out, tg)
    assert sizes == [2, 2, 0, 2, 1, 2] and tuple(C.shape) == (bs, nq, 9)
    from scipy.optimize import linear_sum_assignment
    col = 0
    for b, n in enumerate(sizes):           # each sample is matched inside its own block of columns
        r, c = linear_sum_assignment(C[b, :, col:col + n].numpy())
        assert res[b][0].tolist() == r.tolist() and res[b][1].tolist() == c.tolist()
        assert res[b][0].dtype == torch.int64
        col += n


def test_step_dram_record_is_reproducible_from_the_committed_launch_list():
    """`bench.py` copies the newest profiles/*_step_dram.json into its roofline record (`hbm_view`); that file must be
    what scripts/step_dram.py makes of the committed ncu launch list it names."""
    import subprocess
    prof = os.path.join(REPO, "profiles")
    newest = sorted(f for f in os.listdir(prof) if f.endswith("_step_dram.json"))[-1]
    rec = json.load(open(os.path.join(prof, newest)))
    src = rec["source"].split(" ")[0]
    assert os.path.exists(os.path.join(REPO, src)), src
    out = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "step_dram.py"), os.path.join(REPO, src)],
                         capture_output=True, text=True, check=True).stdout
    again = json.loads(out)
    for key in ("step_launches", "step_kernel_us", "step_dram_bytes"):
        assert again[key] == pytest.approx(rec[key]), key
    assert again["gemm_family"]["launches"] == rec["gemm_family"]["launches"]
    assert again["gemm_family"]["us"] == pytest.approx(rec["gemm_family"]["us"])
    # the family is the step's largest share, and the step is one whole job: its last launch is the rank / top-k kernel
    assert 0.4 < again["gemm_family"]["us"] / again["step_kernel_us"] < 0.7
    assert any(k.startswith("rank_topk_staged_kernel") for k in again["per_kernel"])
