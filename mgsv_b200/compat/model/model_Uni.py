"""model/model_Uni.py of the reference → the made_b200 mirror (same ctor, forward, state_dict keys)."""
from mgsv_b200.model import Uni_model  # noqa: F401
