"""utils/util_test.py of the reference, with the reference's SIGNATURES (test-MaDe.py:402-420 calls them with a
host float64 similarity matrix), over the CUDA rank/top-k and IoU kernels."""
from mgsv_b200.metrics import (calc_similarity, IoU_metrics, Composite_metrics,  # noqa: F401
                               Recall_metrics_matrix as Recall_metrics)
from mgsv_b200.ingest import save_results_json as uni_save_results_json  # noqa: F401
