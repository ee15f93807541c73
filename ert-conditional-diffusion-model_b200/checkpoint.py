"""Checkpoint I/O in the reference's format (ECD.py:342-354, 369-377)."""
from __future__ import annotations

import torch

CHECKPOINT_KEYS = ("epoch", "model_state_dict", "optimizer_state_dict", "best_val_loss",
                   "train_history", "val_history", "param_dim")


def load_best_model(path, model, optimizer=None, map_location=None, weights_only=True):
    """ECD.py:369-377.  Accepts the reference's checkpoint dict or a bare ``state_dict``.

    The reference's checkpoint holds only tensors, lists, floats and ints, so it loads with
    ``weights_only=True`` (no arbitrary unpickling); pass ``weights_only=False`` explicitly for a
    trusted file that needs it."""
    ckpt = torch.load(path, map_location=map_location, weights_only=weights_only)
    sd = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
    model.load_state_dict(sd)
    if optimizer is not None and isinstance(ckpt, dict) and ckpt.get("optimizer_state_dict"):
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return ckpt


def save_checkpoint(path, model, epoch=0, best_val_loss=float("nan"), optimizer=None,
                    train_history=(), val_history=()):
    """Write the dict the reference writes at ECD.py:345-353.  Without an optimizer the
    ``optimizer_state_dict`` key is left out (the reference's loader would feed an empty dict to
    ``optimizer.load_state_dict`` and fail)."""
    ckpt = {"epoch": epoch, "model_state_dict": model.state_dict(),
            "best_val_loss": best_val_loss, "train_history": list(train_history),
            "val_history": list(val_history), "param_dim": model.param_dim}
    if optimizer is not None:
        ckpt["optimizer_state_dict"] = optimizer.state_dict()
    torch.save(ckpt, path)
