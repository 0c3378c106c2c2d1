"""modules/metrics.py of the reference (names imported by test-MaDe.py:17)."""
from mgsv_b200.ops import sim_matrix_music_pooling, sim_matrix_video_pooling  # noqa: F401


def sim_matrix_both_pooling(video_embeds_pooled, music_embeds_pooled):
    """modules/metrics.py:43-57 — only vmr_loss "single_oneloss" reaches it (model_Uni.py:243); not a shipped branch."""
    raise ValueError("Error: vmr_loss 'single_oneloss' (sim_matrix_both_pooling) is not supported by made_b200")
