"""Checkpoint I/O in the reference's format (ECD.py:342-354, 369-377)."""
from __future__ import annotations

import torch

CHECKPOINT_KEYS = ("epoch", "model_state_dict", "optimizer_state_dict", "best_val_loss",
                   "train_history", "val_history", "param_dim")


def load_best_model(path, model, optimizer=None, map_location=None):
    """ECD.py:369-377.  Accepts the reference's checkpoint dict or a bare ``state_dict``."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
    model.load_state_dict(sd)
    if optimizer is not None and isinstance(ckpt, dict) and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return ckpt


def save_checkpoint(path, model, epoch=0, best_val_loss=float("nan"), optimizer=None,
                    train_history=(), val_history=()):
    """Write the dict the reference writes at ECD.py:345-353."""
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
                "best_val_loss": best_val_loss, "train_history": list(train_history),
                "val_history": list(val_history), "param_dim": model.param_dim}, path)
