"""music_detr/matcher.py of the reference → `mgsv_b200.matcher`."""
from mgsv_b200.matcher import HungarianMatcher, build_matcher  # noqa: F401
