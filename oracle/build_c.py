"""Build the oracle's plain-C restatement (oracle/span_oracle.c) with gcc.  Test infrastructure."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "span_oracle.c")
LIB = os.path.join(HERE, "libspan_oracle.so")


def build() -> str:
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def load():
    lib = C.CDLL(build())
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def giou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    load().oracle_giou(_p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]), _p(out))
    return out


def cw_to_se(cw: np.ndarray) -> np.ndarray:
    cw = np.ascontiguousarray(cw, np.float32)
    out = np.empty_like(cw)
    load().oracle_cw_to_se(_p(cw), _p(out), C.c_int64(cw.shape[0]))
    return out


def matcher_cost(prob_fg, out_cw, tgt_cw, w_span=10.0, w_giou=1.0, w_class=4.0) -> np.ndarray:
    p = np.ascontiguousarray(prob_fg, np.float32)
    a, b = np.ascontiguousarray(out_cw, np.float32), np.ascontiguousarray(tgt_cw, np.float32)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    load().oracle_matcher_cost(_p(p), _p(a), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[0]),
                               C.c_float(w_span), C.c_float(w_giou), C.c_float(w_class), _p(out))
    return out


def detr_iou(pred_st, pred_ed, gt_moment, m_duration, max_m_duration=240.0) -> np.ndarray:
    st, ed = np.ascontiguousarray(pred_st, np.float32), np.ascontiguousarray(pred_ed, np.float32)
    gt = np.ascontiguousarray(gt_moment, np.float32).reshape(-1, 2)
    md = np.ascontiguousarray(m_duration, np.float32)
    out = np.empty(st.shape[0], np.float32)
    load().oracle_detr_iou(_p(st), _p(ed), _p(gt), _p(md), C.c_float(max_m_duration), C.c_int64(st.shape[0]), _p(out))
    return out
