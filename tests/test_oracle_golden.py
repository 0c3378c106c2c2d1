"""Pin the CPU oracle (oracle/made_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/gen_golden.py in the build container)."""
import os

import numpy as np
import torch

from mgsv_b200 import synth
from oracle import made_oracle as O
from oracle.gen_golden import dup_tracks


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_span_doctest_vectors(golden_dir):
    # music_detr/span_utils.py:48-54 and :99-103
    s1 = torch.tensor([[0, 0.2], [0.5, 1.0]])
    s2 = torch.tensor([[0, 0.3], [0.0, 1.0]])
    iou, union = O.temporal_iou(s1, s2)
    assert torch.allclose(iou, torch.tensor([[0.6667, 0.2], [0.0, 0.5]]), atol=1e-4)
    assert torch.allclose(union, torch.tensor([[0.3, 1.0], [0.8, 1.0]]), atol=1e-6)
    g = O.generalized_temporal_iou(s1, s2)
    assert torch.allclose(g, torch.tensor([[0.6667, 0.2], [-0.2, 0.5]]), atol=1e-4)
    gold = _g(golden_dir, "span_pairs.npz")
    assert np.array_equal(iou.numpy(), gold["doctest_iou"])
    assert np.array_equal(g.numpy(), gold["doctest_giou"])


def test_span_pairs_bit_exact(golden_dir):
    gold = _g(golden_dir, "span_pairs.npz")
    a, b, logits = synth.make_span_pairs(64, 48, synth.BASE_SEED + 3)
    giou = O.generalized_temporal_iou(O.span_cw_to_se(a), O.span_cw_to_se(b)).numpy()
    assert np.array_equal(giou, gold["giou"], equal_nan=True)
    assert np.isnan(giou).sum() == 1          # Q10: two zero-width spans → 0/0
    tgt = b[b[:, 1] != 0]
    cost = O.matcher_cost(torch.from_numpy(gold["prob_fg"]), a, tgt).numpy()
    assert np.array_equal(cost, gold["cost"], equal_nan=True)


def test_forward_b8(golden_dir, sd_fp32):
    gold = _g(golden_dir, "forward_b8.npz")
    v, m, ids = synth.make_eval_set(8, 8, synth.BASE_SEED + 100)
    out, loss, feat, _, _ = O.uni_forward(sd_fp32, v["frame_feats"], m["segment_feats"], v["frame_mask"],
                                          m["segment_mask"], m["spans_target"])
    tol = dict(atol=2e-5, rtol=1e-4)
    for k in ("video_feats", "music_feats", "frame_feats", "segment_feats"):
        np.testing.assert_allclose(feat[k].numpy(), gold[k], **tol)
    for k in ("pred_logits", "pred_spans", "proj_queries", "proj_vid_mem"):
        np.testing.assert_allclose(out[k].numpy(), gold[k], **tol)
    for i, aux in enumerate(out["aux_outputs"]):
        np.testing.assert_allclose(aux["pred_logits"].numpy(), gold[f"aux{i}_pred_logits"], **tol)
        np.testing.assert_allclose(aux["pred_spans"].numpy(), gold[f"aux{i}_pred_spans"], **tol)
    np.testing.assert_allclose(float(loss["retrieval_loss"]), float(gold["retrieval_loss"]), rtol=1e-5)
    np.testing.assert_allclose(float(loss["localization_loss"]), float(gold["localization_loss"]), rtol=1e-4)
    ld = loss["localization_loss_dict"]
    assert sorted(ld.keys()) == list(gold["loss_names"])
    for k, val in zip(gold["loss_names"], gold["loss_values"]):
        np.testing.assert_allclose(float(ld[str(k)]), val, rtol=2e-4, atol=1e-5)
    pooled = O.xpool(sd_fp32, feat["video_feats"], feat["segment_feats"], m["segment_mask"])
    np.testing.assert_allclose(pooled.numpy(), gold["xpool_pooled"], atol=2e-5, rtol=1e-4)
    mask = torch.cat([v["frame_mask"], m["segment_mask"]], 1)
    np.testing.assert_allclose(O.position_embedding_sine(mask)[:2].numpy(), gold["detr_pos"], atol=1e-6)


def test_cfg1_eval(golden_dir, sd_fp32):
    gold = _g(golden_dir, "cfg1_256.npz")
    N = 256
    v, m, ids = synth.make_eval_set(N, N, synth.BASE_SEED + 1)
    dup_tracks(m, ids, n_dup=16)
    r = O.evaluate(sd_fp32, v, m, ids["music_ids"], batch=32)
    np.testing.assert_allclose(r["video_feats"].numpy(), gold["video_feats"], atol=2e-6)
    np.testing.assert_allclose(r["single"][:64].numpy(), gold["single"], atol=2e-6)
    np.testing.assert_allclose(r["dual"][:64].numpy(), gold["dual"], atol=2e-6)
    np.testing.assert_allclose(r["total"][:64], gold["total"], atol=3e-6)
    # ranks are integers: exact except where two scores are closer than the fp32 noise above
    tot = gold["total"]
    mism = np.nonzero(r["ind"] != gold["ind"])[0]
    assert len(mism) <= 2, mism
    np.testing.assert_allclose(r["pred_st"].numpy(), gold["pred_st"], atol=2e-3)
    np.testing.assert_allclose(r["pred_ed"].numpy(), gold["pred_ed"], atol=2e-3)
    np.testing.assert_allclose(r["score"].numpy(), gold["pred_score"], atol=1e-5)
    np.testing.assert_allclose(r["iou"].numpy(), gold["iou"], atol=2e-5)
    loc = dict(zip(gold["loc_keys"], gold["loc_vals"]))
    np.testing.assert_allclose(r["loc"]["mIoU"], loc["mIoU"], atol=1e-5)
    if len(mism) == 0:
        ret = dict(zip(gold["ret_keys"], gold["ret_vals"]))
        for k, val in ret.items():
            np.testing.assert_allclose(float(r["ret"][str(k)]), val, rtol=1e-12)
        com = dict(zip(gold["com_keys"], gold["com_vals"]))
        for k, val in com.items():
            np.testing.assert_allclose(float(r["com"][str(k)]), val, atol=1e-4)


def test_metrics_dedup_semantics():
    # two columns share an id; the GT's duplicate must not count as a distinct earlier id
    sim = np.array([[0.9, 0.1, 0.95, 0.5],
                    [0.2, 0.8, 0.1, 0.9],
                    [0.3, 0.2, 0.1, 0.0],
                    [0.0, 0.6, 0.7, 0.1]])
    ids = ["a", "b", "a", "c"]
    m, ind, top1 = O.recall_metrics(sim, ids)
    assert list(ind) == [0, 1, 0, 2]
    assert top1 == ["a", "c", "a", "a"]
    assert m["R1"] == 50.0


def test_forward_b8_ca_variant(golden_dir):
    """mml_fusion "CA": the oracle's CrossTransformer restatement against the unmodified reference's outputs."""
    from mgsv_b200 import synth
    from oracle import made_oracle as O
    g = _g(golden_dir, "forward_b8_ca.npz")
    sd = synth.make_state_dict(0, ca=True)
    v, m, ids = synth.make_eval_set(8, 8, synth.BASE_SEED + 100)
    out, loss, feat, mask, _ = O.uni_forward(sd, v["frame_feats"], m["segment_feats"], v["frame_mask"], m["segment_mask"],
                                            m["spans_target"], mml_fusion="CA")
    fused = O.cross_transformer(sd, feat["segment_feats"], feat["frame_feats"], m["segment_mask"], v["frame_mask"])
    fused = fused.masked_fill(m["segment_mask"].unsqueeze(-1) == 0, 0)
    np.testing.assert_allclose(fused.numpy(), g["fused"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(out["pred_spans"].numpy(), g["pred_spans"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(out["pred_logits"].numpy(), g["pred_logits"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(float(loss["localization_loss"]), float(g["localization_loss"]), rtol=1e-5)
    for i in range(5):
        np.testing.assert_allclose(out["aux_outputs"][i]["pred_spans"].numpy(), g[f"aux{i}_pred_spans"], atol=2e-6, rtol=0)
