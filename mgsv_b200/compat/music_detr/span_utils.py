"""music_detr/span_utils.py of the reference → bit-exact CUDA kernels (`mgsv_b200.ops`)."""
from mgsv_b200.ops import (span_cw_to_se, span_se_to_cw, temporal_iou, generalized_temporal_iou,  # noqa: F401
                           detr_iou, span_iou)
