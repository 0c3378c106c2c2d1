"""Reference module names over the made_b200 mirrors — the zero-edit drop-in for `test-MaDe.py`.

The reference's tree is a set of NAMESPACE packages (no `__init__.py` in `model/`, `modules/`, `utils/`,
`music_detr/`), so putting this directory in front of it on `sys.path` makes

    from model.model_Uni import Uni_model                          # test-MaDe.py:22
    from modules.metrics import sim_matrix_music_pooling, ...      # :17
    from modules.loss import *                                     # :23
    from utils.util_test import calc_similarity, Recall_metrics, IoU_metrics, Composite_metrics, ...   # :16
    from music_detr.span_utils import span_cw_to_se, detr_iou      # :25

resolve to the sm_100a-backed mirrors, while everything this build does not replace (`utils.util_train`,
`utils.scheduler`, `dataloaders.*`) still resolves to the reference's own files.  The only change to the driver:

    import mgsv_b200.compat; mgsv_b200.compat.install()            # before the reference imports

`eval_epoch` (test-MaDe.py:243-447) then runs unmodified: `model(...)`, the per-sample post-processing,
`model.video_guided_to_music_pooling_cross_transformer(...)` (materialised, fp32-exact, capped in size),
`sim_matrix_music_pooling`, `calc_similarity`, `Recall_metrics(sim_matrix, dedup=True, ...)`, `detr_iou`,
`IoU_metrics`, `Composite_metrics`.  For throughput use `GalleryEvaluator.run` instead.
"""
import os
import sys

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))


def install() -> str:
    """Put the compat directory at sys.path[0] (idempotent) and return it."""
    if COMPAT_DIR in sys.path:
        sys.path.remove(COMPAT_DIR)
    sys.path.insert(0, COMPAT_DIR)
    # namespace packages that were imported before install() keep their old search path: extend them
    for name in ("model", "modules", "utils", "music_detr"):
        mod = sys.modules.get(name)
        path = getattr(mod, "__path__", None)
        sub = os.path.join(COMPAT_DIR, name)
        if mod is not None and path is not None and sub not in list(path):
            try:
                path._path.insert(0, sub)      # _NamespacePath
            except AttributeError:
                pass
    return COMPAT_DIR
