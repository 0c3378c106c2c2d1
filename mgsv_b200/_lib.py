"""ctypes binding of libmade_b200.so (include/made_b200.h).  No CPU fallback: if the library is
missing it is built with nvcc; if that fails, importing raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import build as _build

_LIB: Optional[C.CDLL] = None

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED, ESTATE = 0, -1, -2, -3, -4, -5
F32, BF16, F16 = 0, 1, 2
VIDEO, MUSIC = 0, 1
PREC_FP16, PREC_SPLIT, PREC_FP32 = 0, 1, 2
ABI_VERSION = 2

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_f = C.c_float

class Ragged(C.Structure):
    """include/made_b200.h: made_ragged (device pointers into the caller's int32 index workspace)."""
    _fields_ = [("seq_len", _p), ("seq_off", _p), ("total", _p), ("tok_src", _p), ("B", _i64), ("L", C.c_int32)]


# name -> argtypes, in the order of include/made_b200.h
SIGNATURES = {
    "made_abi_version": [],
    "made_device_check": [_i32],
    "made_prof_enable": [_i32],
    "made_prof_collect": [C.POINTER(C.c_double), C.POINTER(_i64), _i32],
    "made_span_cw_to_se": [_p, _p, _i64, _p],
    "made_span_se_to_cw": [_p, _p, _i64, _p],
    "made_span_iou": [_p, _p, _p, _p, _f, _i64, _p, _p],
    "made_giou": [_p, _i64, _p, _i64, _p, _p],
    "made_temporal_iou": [_p, _i64, _p, _i64, _p, _p, _p],
    "made_matcher_cost": [_p, _p, _i64, _p, _i64, _f, _f, _f, _p, _p],
    "made_moment_postproc": [_p, _p, _p, _p, _f, _i64, _p, _p, _p, _p, _p],
    "made_rank_topk": [_p, _p, _i64, _i64, _i64, _p, _p, _p, C.c_int32, _i32, _p, _p, _p, _p, _p],
    "made_topk_merge": [_p, _p, _i64, _i32, _i32, _p, _p, _p],
    "made_cosine_sim": [_p, _i64, _p, _i64, _i32, _p, _i64, _p],
    "made_ctx_create": [C.POINTER(_p), _i32],
    "made_ctx_destroy": [_p],
    "made_ctx_load_weights": [_p, _i32, C.POINTER(C.c_char_p), C.POINTER(_p), C.POINTER(_i64), _p],
    "made_ctx_set_precision": [_p, _i32],
    "made_ctx_operand_width": [_p, _i32],
    "made_h2d_valid_rows": [_p, _i32, _p, _i64, _i32, _i32, _p, _i32, _p, C.POINTER(_i64), _p],
    "made_ragged_build": [_p, _i64, _i32, _p, C.POINTER(Ragged), _p],
    "made_ingest_ragged": [_p, _p, _i32, C.POINTER(Ragged), _i32, _p, _p],
    "made_encode": [_p, _i32, _p, _i32, _p, _i64, _p, _p, _p, _p],
    "made_encode_ragged": [_p, _i32, _p, C.POINTER(Ragged), _p, _p, _p, _p],
    "made_gallery_prepare": [_p, _p, _p, _i64, _p, _p, _p, _p],
    "made_query_prepare": [_p, _p, _i64, _p, _p, _p],
    "made_xpool_score": [_p, _p, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _p],
    "made_xpool_pooled": [_p, _i32, _p, _i64, _p, _p, _i64, _p, _p],
    "made_ca_fuse": [_p, _p, _p, _p, _p, _i64, _p, _p, _p],
    "made_pooled_cosine": [_p, _p, _i64, _i64, _p, _i64, _i64, _p],
    "made_detr_detect": [_p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p],
    "made_detr_losses": [_p, _p, _p, _p, _p, _i64, _i32, _f, _f, _f, _p, _p],
    "made_retrieval_loss": [_p, _p, _i64, _i32, _f, _p, _p],
    "made_gemm_f16": [_p, _p, _i64, _i32, _i32, _p, _p, _i32, _p, _p, _p, _p, _p],
    "made_gemm_f16_split": [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _i32, _p, _p, _p, _p, _p],
    "made_gemm_f16_split_h": [_p, _p, _i64, _i32, _i32, _i32, _p, _i32, _p, _p],
    "made_ffn_fused": [_p, _i64, _p, _p, _p, _p, _i32, _p, _i64, _p, _p, _p, _i64, _i32, _i64, _p],
    "made_mha_core": [_p, _p, _p, _p, _i64, _i32, _p, _p],
}


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if _build.needs_build():
        _build.build()
    lib = C.CDLL(_build.LIB)
    lib.made_last_error_string.restype = C.c_char_p
    lib.made_last_error_string.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.made_ragged_index_words.argtypes = [_i64, _i32]
    lib.made_ragged_index_words.restype = C.c_int64
    _LIB = lib
    return lib


def check(rc: int) -> None:
    """Map MADE_E* to the reference's exception conventions (ValueError for unsupported/invalid
    arguments, RuntimeError for runtime failures)."""
    if rc == OK:
        return
    msg = load().made_last_error_string().decode("utf-8", "replace")
    if rc in (EINVAL, EUNSUPPORTED):
        raise ValueError(f"made_b200: {msg}")
    raise RuntimeError(f"made_b200 (code {rc}): {msg}")


PROF_KINDS = ("gemm", "ffn", "xpool", "attn", "rank")


def prof_enable(on: bool) -> None:
    check(load().made_prof_enable(1 if on else 0))


def prof_collect():
    """→ {family: (total ms, launches)} of the window since the last collect (synchronises the device)."""
    n = len(PROF_KINDS)
    ms = (C.c_double * n)()
    cnt = (_i64 * n)()
    check(load().made_prof_collect(ms, cnt, n))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(PROF_KINDS)}


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("made_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise ValueError("made_b200 needs contiguous tensors")
    return t.data_ptr()


def ptr_any(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device tensor or PINNED host tensor (read in place over PCIe by made_ingest_ragged)."""
    if t is not None and not t.is_cuda:
        if not t.is_pinned():
            raise RuntimeError("host tensors handed to made_b200 must be pinned (tensor.pin_memory())")
        if not t.is_contiguous():
            raise ValueError("made_b200 needs contiguous tensors")
        return t.data_ptr()
    return ptr(t)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("made_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
