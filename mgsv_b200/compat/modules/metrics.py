"""modules/metrics.py of the reference (names imported by test-MaDe.py:17)."""
from mgsv_b200.ops import sim_matrix_music_pooling  # noqa: F401


def sim_matrix_video_pooling(video_embeds_pooled, music_embeds):
    """modules/metrics.py:26-41 — only the `XA-video*` fusions call it; the shipped config is `XA-music`."""
    raise ValueError("Error: vmr_fusion 'XA-video' is not supported by made_b200 (shipped: vmr_fusion='XA-music')")


def sim_matrix_both_pooling(video_embeds_pooled, music_embeds_pooled):
    """modules/metrics.py:43-57 on materialised tensors → `mgsv_b200.variants.sim_matrix_both_pooling`."""
    from mgsv_b200.variants import sim_matrix_both_pooling as f
    return f(video_embeds_pooled, music_embeds_pooled)
