"""modules/loss.py of the reference: `cal_distance` (COS; util_test.calc_similarity calls it) and the
evaluation-time VALUES of CLIPLoss / InfoNCELoss (there is no backward through these mirrors)."""
from mgsv_b200.ops import cal_distance  # noqa: F401
from mgsv_b200 import losses as _losses


def CLIPLoss(sims, logit_scale):
    """modules/loss.py:5-24."""
    return _losses.clip_loss(sims, float(logit_scale))


def InfoNCELoss(output, logit_scale, audio_id=None, distance_type="COS", args=None, is_train=False):
    """modules/loss.py:66-123 → (loss, logits_per_video, logits_per_audio).  The same-music masking branch needs
    `is_train and args.ignore_same_music == 0` (never true under test-MaDe.py) and is not built."""
    if audio_id is not None and is_train and args is not None and getattr(args, "ignore_same_music", 1) == 0:
        raise ValueError("Error: InfoNCELoss with ignore_same_music=0 in training is not supported by made_b200")
    return _losses.info_nce_loss(output, float(logit_scale))


__all__ = ["cal_distance", "CLIPLoss", "InfoNCELoss"]
