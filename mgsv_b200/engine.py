"""made_ctx wrapper: packed weights + the stateful ops (encoders, X-Pool scoring, DETR detection).

Host logic only — every tensor op below is a C-ABI call into libmade_b200.so.  There is no CPU or
eager-PyTorch fallback: constructing an Engine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _lib
from . import config as cfg


PRECISIONS = {"fp16": _lib.PREC_FP16, "split": _lib.PREC_SPLIT, "fp32": _lib.PREC_FP32}


class Engine:
    def __init__(self, device: Optional[torch.device] = None, precision: Optional[str] = None):
        """precision: "split" (default; fp16 (hi, lo) operand pairs where the similarity error is made —
        meets the 1e-3 bar on the full job), "fp16" (single fp16 operands everywhere, fastest) or "fp32" (the
        temporal encoders and X-Pool on the CUDA cores in the reference's fp32 arithmetic: the 1e-5 mode, ~50x
        slower).  MADE_PRECISION overrides the default."""
        _lib.require_cuda()
        precision = precision or os.environ.get("MADE_PRECISION", "split")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("made_b200 needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.made_ctx_create(C.byref(h), self.device.index))
        self._h = h
        _lib.check(self._lib.made_ctx_set_precision(self._h, PRECISIONS[precision]))
        self.loaded = False
        self.mml_fusion = "concat"      # "CA": moment detection runs on made_ca_fuse's output (set by Uni_model / the caller)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.made_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -----------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Reference key names.  Frozen `vit_model.*` / `ast_model.*` entries of
        real checkpoints are ignored (the feature path never uses them)."""
        names, arrs = [], []
        for k, v in sd.items():
            if k.startswith(("vit_model.", "ast_model.")):
                continue
            names.append(k.encode())
            arrs.append(v.detach().to("cpu", torch.float32).contiguous())
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[a.data_ptr() for a in arrs])
        c_num = (C.c_int64 * n)(*[a.numel() for a in arrs])
        with torch.cuda.device(self.device):
            _lib.check(self._lib.made_ctx_load_weights(self._h, n, c_names, c_ptrs, c_num, _lib.stream_ptr()))
        self.loaded = True
        self._weights = dict(zip([k.decode() for k in names], arrs))    # host fp32 masters, for clone()

    def clone(self) -> "Engine":
        """A second context with the same weights and its own workspace arena: calls on one made_ctx
        are serialised by contract, so work that should overlap on another stream (moment detection
        beside scoring) needs its own context (21 MB of packed weights)."""
        if not self.loaded:
            raise RuntimeError("clone() needs loaded weights")
        other = Engine(self.device, self.precision)
        other.load_state_dict(self._weights)
        other.mml_fusion = self.mml_fusion
        return other

    # -----------------------------------------------------------------------------------------
    def h2d_valid_rows(self, feats_host: torch.Tensor, masks_host: torch.Tensor, out: torch.Tensor,
                       host_stage16: Optional[torch.Tensor] = None, n_threads: int = 1) -> int:
        """Copy-engine transfer of the valid rows of a zero-padded host feature tensor [B,L,dim]
        (pinned) into the device staging tensor `out`; rows whose mask is 0 are not transferred (and
        are garbage in `out`).  With `host_stage16` (pinned fp16 [B,L,dim]) the host threads round the
        valid rows to fp16 first and `out` is an fp16 tensor: half the PCIe bytes.  Returns the bytes
        queued."""
        if feats_host.is_cuda or masks_host.is_cuda or not out.is_cuda:
            raise ValueError("h2d_valid_rows: host features + host masks -> device staging")
        want = torch.float16 if host_stage16 is not None else feats_host.dtype
        if feats_host.shape != out.shape or out.dtype != want:
            raise ValueError("h2d_valid_rows: staging must match the host tensor's shape (and dtype)")
        if host_stage16 is not None and (host_stage16.shape != feats_host.shape or host_stage16.dtype != torch.float16
                                         or not host_stage16.is_pinned() or feats_host.dtype != torch.float32):
            raise ValueError("h2d_valid_rows: host_stage16 must be a pinned fp16 tensor shaped like the fp32 features")
        dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}.get(feats_host.dtype)
        if dt is None:
            raise ValueError(f"unsupported feature dtype {feats_host.dtype}")
        B, L, dim = feats_host.shape
        m = masks_host.to(torch.float32).contiguous()
        n = C.c_int64(0)
        _lib.check(self._lib.made_h2d_valid_rows(
            feats_host.contiguous().data_ptr(), dt, m.data_ptr(), B, L, dim,
            None if host_stage16 is None else host_stage16.data_ptr(), int(n_threads), _lib.ptr(out), C.byref(n),
            _lib.stream_ptr()))
        return int(n.value)

    def ragged(self, masks: torch.Tensor):
        """Token-packing descriptor of a batch of masks [B, L] (device, float): → (made_ragged struct,
        index tensor that owns its memory — keep both alive while kernels that use it are pending)."""
        if not masks.is_cuda or masks.dim() != 2:
            raise ValueError("ragged: expected a device mask of shape [B, L]")
        masks = masks.to(torch.float32).contiguous()
        B, L = masks.shape
        words = int(self._lib.made_ragged_index_words(B, L))
        idx = torch.empty(max(words, 1), dtype=torch.int32, device=masks.device)
        rb = _lib.Ragged()
        if B:
            _lib.check(self._lib.made_ragged_build(_lib.ptr(masks), B, L, _lib.ptr(idx), C.byref(rb), _lib.stream_ptr()))
        else:
            rb.B, rb.L = 0, L
        return rb, (idx, masks)

    def ingest(self, modality: int, feats: torch.Tensor, rb, out: Optional[torch.Tensor] = None):
        """Masked cast of raw features (device tensor or PINNED host tensor, fp32/bf16/fp16) into the
        token-packed fp16 operand buffer consumed by `encode(..., ragged=rb)`.  Padded rows are never
        read, so a pinned host tensor costs only its valid rows of PCIe traffic."""
        L, din = (cfg.L_V, cfg.D_VIT) if modality == _lib.VIDEO else (cfg.L_M, cfg.D_AST)
        if feats.dim() != 3 or feats.shape[1] != L or feats.shape[2] != din or feats.shape[0] != rb.B or rb.L != L:
            raise ValueError(f"expected features of shape [{rb.B},{L},{din}], got {tuple(feats.shape)}")
        dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}.get(feats.dtype)
        if dt is None:
            raise ValueError(f"unsupported feature dtype {feats.dtype}")
        B = feats.shape[0]
        width = self.operand_width(din)
        if out is None:
            out = torch.empty((B * L, width), dtype=torch.float16, device=self.device)
        elif out.shape[-1] != width or out.dtype != torch.float16:
            raise ValueError(f"ingest: out must be fp16 with rows of {width} columns")
        _lib.check(self._lib.made_ingest_ragged(self._h, _lib.ptr_any(feats), dt, C.byref(rb), din, _lib.ptr(out),
                                                _lib.stream_ptr()))
        return out

    def operand_width(self, dim: int) -> int:
        """Columns of a packed operand row of `dim` features ((hi | lo) pairs in split precision)."""
        return int(self._lib.made_ctx_operand_width(self._h, dim))

    def encode(self, modality: int, feats: torch.Tensor, masks: torch.Tensor, want_f32: bool = True,
               ragged=None, out=None):
        """forward_{video,audio}_encoder_feature → (seq16 [B,L,256], seq_f32 or None, pooled [B,256]).
        `ragged=rb`: feats is the packed fp16 buffer written by `ingest(..., rb)` (used in place).
        `out=(seq16, pooled)` writes into caller-provided (contiguous slices of) tensors."""
        L, din = (cfg.L_V, cfg.D_VIT) if modality == _lib.VIDEO else (cfg.L_M, cfg.D_AST)
        B = masks.shape[0]
        if tuple(masks.shape) != (B, L):
            raise ValueError(f"expected masks of shape [B,{L}], got {tuple(masks.shape)}")
        dev = masks.device if masks.is_cuda else feats.device
        if out is not None:
            seq, pooled = out
        else:
            seq = torch.empty((B, L, cfg.D_MODEL), dtype=torch.float16, device=dev)
            pooled = torch.empty((B, cfg.D_MODEL), dtype=torch.float32, device=dev)
        seq32 = torch.empty((B, L, cfg.D_MODEL), dtype=torch.float32, device=dev) if want_f32 else None
        if ragged is not None:
            if feats.dtype != torch.float16 or feats.dim() != 2 or feats.shape[1] != self.operand_width(din) \
                    or ragged.B != B:
                raise ValueError("ragged encode takes the packed fp16 output of Engine.ingest")
            _lib.check(self._lib.made_encode_ragged(self._h, modality, _lib.ptr(feats), C.byref(ragged), _lib.ptr(seq),
                                                    _lib.ptr(seq32), _lib.ptr(pooled), _lib.stream_ptr()))
            return seq, seq32, pooled
        if feats.dim() != 3 or feats.shape[0] != B or feats.shape[1] != L or feats.shape[2] != din:
            raise ValueError(f"expected features of shape [B,{L},{din}], got {tuple(feats.shape)}")
        dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}.get(feats.dtype)
        if dt is None:
            raise ValueError(f"unsupported feature dtype {feats.dtype}")
        feats = feats.contiguous()
        masks = masks.to(torch.float32).contiguous()
        _lib.check(self._lib.made_encode(self._h, modality, _lib.ptr(feats), dt, _lib.ptr(masks), B, _lib.ptr(seq),
                                         _lib.ptr(seq32), _lib.ptr(pooled), _lib.stream_ptr()))
        return seq, seq32, pooled

    def gallery_prepare(self, seg16: torch.Tensor, seg_masks: torch.Tensor, out=None):
        """Per-track X-Pool operands: kz [N*96,768] fp16, gram [N*96,112] fp16 (= [G | W5 | 0]), maskbits [N,4] int32.
        `out=(kz, gram, bits)` writes into caller-provided (contiguous slices of) tensors."""
        N = seg16.shape[0]
        dev = seg16.device
        if seg16.dtype != torch.float16:
            raise ValueError("gallery_prepare takes the fp16 encoded segments")
        if out is not None:
            kz, gram, bits = out
        else:
            kz = torch.empty((N * cfg.L_M, 3 * cfg.D_MODEL), dtype=torch.float16, device=dev)
            gram = torch.empty((N * cfg.L_M, cfg.XPOOL_G_COLS), dtype=torch.float16, device=dev)
            bits = torch.empty((N, 4), dtype=torch.int32, device=dev)
        masks = seg_masks.to(torch.float32).contiguous()
        _lib.check(self._lib.made_gallery_prepare(self._h, _lib.ptr(seg16.contiguous()), _lib.ptr(masks), N,
                                                  _lib.ptr(kz), _lib.ptr(gram), _lib.ptr(bits), _lib.stream_ptr()))
        return kz, gram, bits

    def query_prepare(self, video_feats: torch.Tensor):
        N = video_feats.shape[0]
        vf = video_feats.to(torch.float32).contiguous()
        q = torch.empty((N, cfg.D_MODEL), dtype=torch.float16, device=vf.device)
        vhat = torch.empty((N, cfg.D_MODEL), dtype=torch.float32, device=vf.device)
        _lib.check(self._lib.made_query_prepare(self._h, _lib.ptr(vf), N, _lib.ptr(q), _lib.ptr(vhat),
                                                _lib.stream_ptr()))
        return q, vhat

    def xpool_pooled(self, video_feats: torch.Tensor, segment_feats: torch.Tensor, segment_masks: torch.Tensor,
                     out: Optional[torch.Tensor] = None, track_chunk: Optional[int] = None,
                     which: int = _lib.MUSIC) -> torch.Tensor:
        """Transformer_XA.forward MATERIALISED (modules/transformer.py:156-180): video_feats [N_v,256],
        segment_feats [N_m,96,256], segment_masks [N_m,96] → [N_m, N_v, 256] fp32, in the reference's fp32
        arithmetic on the CUDA cores.  The scoring path never forms this tensor (`xpool_score`); this is for callers
        that want it and for the fp32 precision mode.  Tracks are processed `track_chunk` at a time (default:
        ~256 MB of scratch).  `which=_lib.VIDEO` runs the second module of vmr_fusion "XA-music-video"
        (model.music_guided_to_video_pooling_cross_transformer): guides = music_feats [N_m,256] in `video_feats`, keys =
        frame features [N_v,50,256] + masks in `segment_feats` / `segment_masks` → [N_v, N_m, 256]."""
        dev = self.device
        vf = video_feats.to(dev, torch.float32).contiguous()
        sf = segment_feats.to(dev, torch.float32).contiguous()
        sm = segment_masks.to(dev, torch.float32).contiguous()
        n_q, n_m = vf.shape[0], sf.shape[0]
        if out is None:
            out = torch.empty((n_m, n_q, cfg.D_MODEL), dtype=torch.float32, device=dev)
        if track_chunk is None:
            track_chunk = max(1, min(n_m, (256 << 20) // max(1, n_q * 2432)))
        for s in range(0, n_m, track_chunk):
            e = min(n_m, s + track_chunk)
            _lib.check(self._lib.made_xpool_pooled(self._h, which, _lib.ptr(vf), n_q, _lib.ptr(sf[s:e]), _lib.ptr(sm[s:e]), e - s,
                                                   _lib.ptr(out[s:e]), _lib.stream_ptr()))
        return out

    def ca_fuse(self, segment_feats: torch.Tensor, segment_masks: torch.Tensor, frame_feats: torch.Tensor,
                frame_masks: torch.Tensor, want_f32: bool = False):
        """mml_fusion "CA" (model_Uni.py:209-211): CrossTransformer(segment_feats [B,96,256] as queries, frame_feats
        [B,50,256] as keys / values, both masks) followed by the masked_fill of padded segments → (fused fp16
        [B,96,256], fused fp32 or None).  DETR then runs on the fused segments alone."""
        dev = self.device
        sf = segment_feats.to(dev, torch.float32).contiguous()
        ff = frame_feats.to(dev, torch.float32).contiguous()
        sm = segment_masks.to(dev, torch.float32).contiguous()
        fm = frame_masks.to(dev, torch.float32).contiguous()
        B = sf.shape[0]
        if tuple(sf.shape) != (B, cfg.L_M, cfg.D_MODEL) or tuple(ff.shape) != (B, cfg.L_V, cfg.D_MODEL) \
                or tuple(sm.shape) != (B, cfg.L_M) or tuple(fm.shape) != (B, cfg.L_V):
            raise ValueError("ca_fuse: expected segment_feats [B,96,256], frame_feats [B,50,256] and their masks")
        out16 = torch.empty((B, cfg.L_M, cfg.D_MODEL), dtype=torch.float16, device=dev)
        out32 = torch.empty((B, cfg.L_M, cfg.D_MODEL), dtype=torch.float32, device=dev) if want_f32 else None
        _lib.check(self._lib.made_ca_fuse(self._h, _lib.ptr(sf), _lib.ptr(sm), _lib.ptr(ff), _lib.ptr(fm), B,
                                          _lib.ptr(out16), _lib.ptr(out32), _lib.stream_ptr()))
        return out16, out32

    def xpool_score(self, q, vhat, kz, gram, bits, out: Optional[torch.Tensor] = None, col_offset: int = 0):
        n_q, n_m = q.shape[0], bits.shape[0]
        if out is None:
            out = torch.empty((n_q, n_m), dtype=torch.float32, device=q.device)
            col_offset = 0
        _lib.check(self._lib.made_xpool_score(self._h, _lib.ptr(q), _lib.ptr(vhat), n_q, _lib.ptr(kz), _lib.ptr(gram),
                                              _lib.ptr(bits), n_m, _lib.ptr(out), out.stride(0), col_offset,
                                              _lib.stream_ptr()))
        return out

    def detr_detect(self, frame16, frame_masks, seg16, seg_masks, video_feats, track_idx=None,
                    want_proj: bool = False, want_memory: bool = False):
        """→ dict(hs [6,B,256], pred_logits [6,B,2], pred_spans [6,B,2], proj_queries?, proj_vid_mem?, memory?)."""
        B = video_feats.shape[0]
        dev = video_feats.device
        f32 = dict(dtype=torch.float32, device=dev)
        hs = torch.empty((cfg.DETR_DEC_LAYERS, B, cfg.D_MODEL), **f32)
        logits = torch.empty((cfg.DETR_DEC_LAYERS, B, 2), **f32)
        spans = torch.empty((cfg.DETR_DEC_LAYERS, B, 2), **f32)
        pq = torch.empty((cfg.DETR_DEC_LAYERS, B, cfg.D_MODEL), **f32) if want_proj else None
        pv = torch.empty((B, cfg.L_V, cfg.D_MODEL), **f32) if want_proj else None
        mem = torch.empty((B, cfg.L_DETR, cfg.D_MODEL), **f32) if want_memory else None
        if track_idx is not None:
            track_idx = track_idx.to(torch.int32).contiguous()
        _lib.check(self._lib.made_detr_detect(
            self._h, _lib.ptr(frame16.contiguous()), _lib.ptr(frame_masks.to(torch.float32).contiguous()),
            _lib.ptr(seg16.contiguous()), _lib.ptr(seg_masks.to(torch.float32).contiguous()), _lib.ptr(track_idx),
            _lib.ptr(video_feats.to(torch.float32).contiguous()), B, _lib.ptr(hs), _lib.ptr(logits), _lib.ptr(spans),
            _lib.ptr(pq), _lib.ptr(pv), _lib.ptr(mem), _lib.stream_ptr()))
        return dict(hs=hs, pred_logits=logits, pred_spans=spans, proj_queries=pq, proj_vid_mem=pv, memory=mem)
