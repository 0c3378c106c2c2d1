"""Host-side mirror of music_detr/matcher.py: `HungarianMatcher(args, ...).forward(outputs, targets)`.

The O(bs*Q x sum(targets)) cost matrix — softmax foreground probability, L1 distance on (c, w),
generalized temporal IoU on (start, end) — is one CUDA kernel (`made_matcher_cost`, bit-exact fp32
given the probabilities); the linear-sum-assignment stays scipy on the host exactly like the
reference (matcher.py:89-92: `.cpu()` + `linear_sum_assignment` per sample; with the shipped one
query / one target per sample it is the identity).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class HungarianMatcher(nn.Module):
    def __init__(self, args, cost_class: float = 1, cost_span: float = 1, cost_giou: float = 1,
                 span_loss_type: str = "l1", snippet_num: int = 100):
        super().__init__()
        if span_loss_type != "l1":
            raise ValueError("made_b200 builds the shipped span_loss_type='l1' matcher only")
        assert cost_class != 0 or cost_span != 0 or cost_giou != 0, "all costs cant be 0"     # matcher.py:32
        self.args = args
        self.cost_class, self.cost_span, self.cost_giou = cost_class, cost_span, cost_giou
        self.span_loss_type, self.snippet_num = span_loss_type, snippet_num
        self.foreground_label = 0 if getattr(args, "fb_label", "01") == "01" else 1
        if self.foreground_label != 0:
            raise ValueError("made_b200 builds the shipped fb_label='01' only")

    @torch.no_grad()
    def cost_matrix(self, outputs, targets):
        """→ (C [bs, Q, sum(targets)] on the device, sizes list) — matcher.py:55-89."""
        bs, nq = outputs["pred_spans"].shape[:2]
        logits = outputs["pred_logits"].flatten(0, 1)
        out_spans = outputs["pred_spans"].flatten(0, 1)
        prob_fg = ops.moment_postproc(logits, out_spans)[2]            # softmax(-1)[:, foreground]
        moment_mask = targets[:, :, 1] != 0                             # matcher.py:59
        tgt_spans = targets[moment_mask]
        sizes = moment_mask.sum(dim=1).tolist()
        C = ops.matcher_cost(prob_fg, out_spans, tgt_spans, cost_span=self.cost_span, cost_giou=self.cost_giou,
                             cost_class=self.cost_class)
        return C.view(bs, nq, -1), sizes

    @torch.no_grad()
    def forward(self, outputs, targets):
        from scipy.optimize import linear_sum_assignment
        C, sizes = self.cost_matrix(outputs, targets)
        # the assignment itself stays scipy on the host like the reference (matcher.py:90-92): sample b is
        # matched inside its own block of target columns (identity at one query / one target)
        host = C.cpu().numpy()
        pairs, col = [], 0
        for b, n_tgt in enumerate(sizes):
            rows, cols = linear_sum_assignment(host[b, :, col:col + n_tgt])
            pairs.append((torch.as_tensor(rows, dtype=torch.int64), torch.as_tensor(cols, dtype=torch.int64)))
            col += n_tgt
        return pairs


def build_matcher(args):
    """matcher.py:95-103."""
    return HungarianMatcher(args, cost_span=10, cost_giou=1, cost_class=4, span_loss_type=args.span_loss_type,
                            snippet_num=args.max_snippet_num)
