"""Checkpoint files of the reference (utils/util_train.py:21-60): `pytorch_model.bin.<epoch | best_r1 | best_iou |
best_r1iou05 | best_r1iou07>` = torch.save({"epoch", "loss", "model_state_dict", "optimizer_state_dict"}).

`load_model` keeps the reference's signature and return value, so test-MaDe.py:485-516 works on a
`mgsv_b200.model.Uni_model`; the frozen `vit_model.*` / `ast_model.*` backbones that real checkpoints carry are
skipped (the feature path never reads them).  `load_checkpoint` is the same without the args plumbing, and packs the
weights for the device right away (fp16 operands, folded X-Pool / decoder matrices: `Engine.load_state_dict`) instead
of at the first forward.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch

_FROZEN = ("vit_model.", "ast_model.", "mert_model.", "vivit_model.", "cnclip_model.")      # test-MaDe.py:226


def read_state_dict(path: str) -> Tuple[Dict[str, torch.Tensor], int, float]:
    """→ (state_dict without frozen backbones and without a DDP `module.` prefix, epoch, loss)."""
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    sd = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt   # util_train.py:53
    if not isinstance(sd, dict):
        raise ValueError(f"{path}: not a state_dict checkpoint")
    out = {}
    for k, v in sd.items():
        k = k[7:] if k.startswith("module.") else k
        if k.startswith(_FROZEN) or not isinstance(v, torch.Tensor):
            continue
        out[k] = v
    epoch = int(ckpt.get("epoch", 0)) if isinstance(ckpt, dict) and "model_state_dict" in ckpt else 0
    loss = float(ckpt.get("loss", 0)) if isinstance(ckpt, dict) and "model_state_dict" in ckpt else 0.0
    return out, epoch, loss


def load_checkpoint(model, path: str, strict: bool = True, pack: bool = True):
    """Load a reference checkpoint file into `model` (a mgsv_b200 Uni_model) → (epoch, loss)."""
    sd, epoch, loss = read_state_dict(path)
    model.load_state_dict(sd, strict=strict)
    if pack and torch.cuda.is_available():
        model.engine()          # repack for the device now
    return epoch, loss


def load_model(args, logger, model, stage, optimizer=None):
    """utils/util_train.py:38-60, same arguments and 4-tuple."""
    model = model.module if hasattr(model, "module") else model
    if getattr(args, "resume_path", None) is not None:
        path = args.resume_path
    elif stage == 1:
        path = args.load_retrieval_model_path
    elif stage == 2:
        path = args.load_grounding_model_path
    elif stage == 0:
        path = args.load_uni_model_path
    else:
        raise ValueError("Invalid stage")
    epoch, loss = load_checkpoint(model, path)
    if optimizer is not None:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        if isinstance(ckpt, dict) and "optimizer_state_dict" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if getattr(args, "local_rank", 0) == 0 and logger is not None:
        logger.info("Model loaded from %s", path)
    return model, optimizer, epoch, loss


def save_model(epoch, args, logger, model, optimizer=None, loss=None, best_model=False, best_name="best"):
    """utils/util_train.py:21-36: same arguments, file name pattern and dictionary."""
    if getattr(args, "save_model", 1) == 0:
        return None
    model = model.module if hasattr(model, "module") else model
    name = f"pytorch_model.bin.{best_name}" if best_model else f"pytorch_model.bin.{epoch}"
    path = os.path.join(args.path_log, name)
    torch.save({"epoch": epoch, "loss": loss if loss is not None else "None", "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else "None"}, path)
    if logger is not None:
        logger.info("Model saved to %s", path)
    return path
