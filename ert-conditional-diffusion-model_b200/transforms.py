"""After the chain: logits -> physical parameters -> bounds check.

Device replacement for the three host steps the reference runs on every realisation
(ERT_Conditional_Diffusion.py:402-406): ``inverse_transform`` (ECD.py:42-53),
``MinMaxScaler.inverse_transform`` and ``check_param_bounds`` (ECD.py:183-218, limits from
Generate_ERT_utils.py:8-59), fused into one elementwise kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _f64(dev, v):
    return None if v is None else torch.as_tensor(np.asarray(v, dtype=np.float64)).to(dev).contiguous()


@torch.no_grad()
def untransform_and_check(u, a=0.0, b=1.0, scaler_min=None, scaler_scale=None, limits=None):
    """u ``(B, P)`` float32 CUDA logits.

    ``scaler_min`` / ``scaler_scale``: sklearn ``MinMaxScaler.min_`` / ``.scale_`` (P,), applied
    as sklearn does on a float32 array (``X -= min_; X /= scale_`` in float64, stored float32).
    ``limits``: ``(P, 2)`` ``[min, max]`` rows (``ParameterLimits().plims``).
    Returns ``(phys (B,P) float32, valid (B,) bool, first_bad (B,) int32)``; ``first_bad`` is the
    first out-of-bounds parameter index, the one ``check_param_bounds`` reports, or -1.
    """
    if u.device.type != "cuda":
        raise _lib.ErtdiffError("untransform_and_check runs on CUDA only")
    u = u.to(torch.float32).contiguous()
    B, P = u.shape
    dev = u.device
    smin, sscale = _f64(dev, scaler_min), _f64(dev, scaler_scale)
    lim = None if limits is None else np.asarray(limits, dtype=np.float64)
    lo = _f64(dev, None if lim is None else lim[:, 0])
    hi = _f64(dev, None if lim is None else lim[:, 1])
    phys = torch.empty_like(u)
    valid = torch.empty(B, device=dev, dtype=torch.uint8)
    first_bad = torch.empty(B, device=dev, dtype=torch.int32)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ertdiff_untransform_bounds(
                _lib.ptr(u), B, P, float(a), float(b), _lib.ptr(smin), _lib.ptr(sscale),
                _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(phys), _lib.ptr(valid), _lib.ptr(first_bad),
                _lib.stream_ptr(dev)), "untransform_bounds")
    return phys, valid.bool(), first_bad


def inverse_transform(u, a, b):
    """ECD.py:42-53 for CUDA tensors: ``a + (b-a)*sigmoid(u)``."""
    return untransform_and_check(u, a, b)[0]


def check_param_bounds(param, limits, verbose=True):
    """ECD.py:183-218 on the device: the rows of ``param`` (B,P) whose parameters all lie inside ``limits``
    (P,2), in order; ``None`` when no row survives (as the reference returns).  Same semantics as the
    reference's loop: a row is dropped when some parameter is ``< min or > max`` (compared in float64, as
    numpy promotes; a NaN never drops a row), and with ``verbose`` the first offending parameter of every dropped
    row is printed in the reference's format.  float32 and float64 inputs keep their dtype."""
    was_numpy = not isinstance(param, torch.Tensor)
    t = torch.as_tensor(np.asarray(param)) if was_numpy else param
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    if t.device.type != "cuda":
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    t = t.contiguous()
    lim = np.asarray(limits, dtype=np.float64)
    B, P = t.shape
    if P > 32:
        raise _lib.ErtdiffError("check_param_bounds: at most 32 parameters")
    dev = t.device
    lo, hi = _f64(dev, lim[:, 0]), _f64(dev, lim[:, 1])
    valid = torch.empty(B, device=dev, dtype=torch.uint8)
    first_bad = torch.empty(B, device=dev, dtype=torch.int32)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ertdiff_check_bounds(
                _lib.ptr(t), _lib.F32 if t.dtype == torch.float32 else _lib.F64, B, P, _lib.ptr(lo), _lib.ptr(hi),
                _lib.ptr(valid), _lib.ptr(first_bad), _lib.stream_ptr(dev)), "check_bounds")
    ok = valid.bool().cpu().numpy()
    if verbose and not ok.all():
        fb = first_bad.cpu().numpy()
        host = param if was_numpy else t.cpu().numpy()
        for i in np.nonzero(~ok)[0]:
            j = int(fb[i])
            print(f"Sample {i} Parameter {j}: {host[i][j]:.4f} (out of bounds [{lim[j, 0]:.4f}, {lim[j, 1]:.4f}])")
    kept = param[ok] if was_numpy else param[torch.from_numpy(ok).to(param.device)]
    return kept if kept.shape[0] else None
