"""Ensemble statistics over members (axis 0): device replacements for the numpy / scipy
calls the reference makes inline (ERT_Conditional_Diffusion.py:747-762, 867-872, 612,
1126-1127, 1199-1200).

Every function takes an ``(N, ...)`` array -- a CUDA ``torch.Tensor`` or a ``numpy.ndarray``
(copied to the device) -- and returns the same kind of object it was given, with the trailing
shape.  float32 and float64 are supported; the result dtypes follow numpy's rules.
Mean/std/var and percentiles are bit-identical to numpy for ``Q > 1`` columns; the KDE mode's
contract is the argmax grid index.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.float64: _lib.F64}


def _to_device(a, device=None):
    """-> (2-D contiguous CUDA tensor (N, Q), trailing shape, was_numpy)."""
    was_numpy = isinstance(a, np.ndarray)
    if was_numpy:
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev)
    else:
        t = a
        if t.device.type != "cuda":
            if device is None:
                raise _lib.ErtdiffError("statistics run on CUDA only: pass a CUDA tensor, a numpy "
                                        "array, or device=")
            t = t.to(device)
        if t.dtype not in _DT:
            t = t.to(torch.float64)
    if t.dim() < 1 or t.size(0) < 1:
        raise ValueError("need an (N, ...) array with N >= 1 members on axis 0")
    trailing = tuple(t.shape[1:])
    t2 = t.reshape(t.size(0), -1).contiguous()
    return t2, trailing, was_numpy


def _finish(t, trailing, was_numpy, lead=()):
    t = t.reshape(tuple(lead) + tuple(trailing))
    return t.cpu().numpy() if was_numpy else t


def ensemble_moments(a, device=None, mean=True, std=True, var=True):
    """``np.mean / np.std / np.var (a, axis=0)`` in one pass pair (ECD.py:867-869).
    Returns a dict with the requested keys."""
    t, trailing, was_numpy = _to_device(a, device)
    N, Q = t.shape
    outs = {k: torch.empty(Q, device=t.device, dtype=t.dtype)
            for k, want in (("mean", mean), ("std", std), ("var", var)) if want}
    if Q > 0:
        with torch.cuda.device(t.device):
            _lib.check(_lib.load().ertdiff_ensemble_moments(
                _lib.ptr(t), _DT[t.dtype], N, Q, _lib.ptr(outs.get("mean")),
                _lib.ptr(outs.get("std")), _lib.ptr(outs.get("var")), _lib.stream_ptr(t.device)),
                "ensemble_moments")
    return {k: _finish(v, trailing, was_numpy) for k, v in outs.items()}


def ensemble_mean(a, device=None):
    return ensemble_moments(a, device, True, False, False)["mean"]


def ensemble_std(a, device=None):
    return ensemble_moments(a, device, False, True, False)["std"]


def ensemble_var(a, device=None):
    return ensemble_moments(a, device, False, False, True)["var"]


def _index_dtype(a_dtype, q):
    """numpy's rule (``np.percentile`` divides q by ``a.dtype.type(100)``): a python int/float q
    is weakly typed and adopts a float32 array's dtype; numpy scalars / sequences stay float64."""
    weak = isinstance(q, (int, float)) and not isinstance(q, np.generic)
    return _lib.F32 if (weak and a_dtype == torch.float32) else _lib.F64


def ensemble_percentile(a, q, device=None):
    """``np.percentile(a, q, axis=0)`` with the default linear method
    (ECD.py:870-872, 612, 1126-1127, 1199-1200).  Scalar q -> trailing shape; sequence q ->
    ``(len(q),) + trailing``."""
    t, trailing, was_numpy = _to_device(a, device)
    N, Q = t.shape
    scalar = np.ndim(q) == 0
    idt = _index_dtype(t.dtype, q)
    qs = np.atleast_1d(np.asarray(q, dtype=np.float64)).ravel()
    if qs.size == 0:
        raise ValueError("q is empty")
    if np.any(qs < 0) or np.any(qs > 100) or np.any(np.isnan(qs)):
        raise ValueError("Percentiles must be in the range [0, 100]")
    out_dtype = torch.float32 if (t.dtype == torch.float32 and idt == _lib.F32) else torch.float64
    out = torch.empty(qs.size, Q, device=t.device, dtype=out_dtype)
    if Q > 0:
        qarr = (C.c_double * qs.size)(*qs.tolist())
        with torch.cuda.device(t.device):
            _lib.check(_lib.load().ertdiff_ensemble_percentiles(
                _lib.ptr(t), _DT[t.dtype], N, Q, qarr, int(qs.size), idt, _lib.ptr(out),
                _lib.stream_ptr(t.device)), "ensemble_percentiles")
    if scalar:
        return _finish(out[0], trailing, was_numpy)
    return _finish(out, trailing, was_numpy, lead=(qs.size,))


def ensemble_kde_mode(a, n_grid=5000, grid_range=None, device=None, return_index=False):
    """Per-column Gaussian-KDE mode on a common grid (ECD.py:747-762): ``grid =
    linspace(a.min(), a.max(), n_grid)`` over the WHOLE array (or ``grid_range=(lo, hi)``), Scott
    bandwidth, first argmax.  Returns float64 modes (and the int64 grid indices)."""
    t, trailing, was_numpy = _to_device(a, device)
    N, Q = t.shape
    if N < 2:
        raise ValueError("KDE needs at least 2 members")
    lib = _lib.load()
    lohi = torch.empty(2, device=t.device, dtype=torch.float64)
    mode = torch.empty(Q, device=t.device, dtype=torch.float64)
    index = torch.empty(Q, device=t.device, dtype=torch.int64)
    with torch.cuda.device(t.device):
        st = _lib.stream_ptr(t.device)
        if grid_range is None:
            if Q > 0:      # range from the data: one entry point (a single fused launch for small ensembles)
                _lib.check(lib.ertdiff_ensemble_kde_mode_auto(_lib.ptr(t), _DT[t.dtype], N, Q, int(n_grid),
                                                              _lib.ptr(lohi), _lib.ptr(mode), _lib.ptr(index), st),
                           "ensemble_kde_mode")
        else:
            if isinstance(grid_range, torch.Tensor):       # (lo, hi) already on the device: no host round trip
                lohi.copy_(grid_range.to(device=t.device, dtype=torch.float64).reshape(2))
            else:
                lohi.copy_(torch.tensor([float(grid_range[0]), float(grid_range[1])], dtype=torch.float64))
            if Q > 0:
                _lib.check(lib.ertdiff_ensemble_kde_mode(_lib.ptr(t), _DT[t.dtype], N, Q, _lib.ptr(lohi),
                                                         int(n_grid), _lib.ptr(mode), _lib.ptr(index), st),
                           "ensemble_kde_mode")
    m = _finish(mode, trailing, was_numpy)
    return (m, _finish(index, trailing, was_numpy)) if return_index else m


def global_minmax(a, device=None):
    """``(np.min(a), np.max(a))`` over the whole array as a 2-element float64 CUDA tensor: the range
    of the KDE grid (ECD.py:749-750)."""
    t, _, _ = _to_device(a, device)
    lohi = torch.empty(2, device=t.device, dtype=torch.float64)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().ertdiff_minmax(_lib.ptr(t), _DT[t.dtype], t.numel(), _lib.ptr(lohi),
                                              _lib.stream_ptr(t.device)), "minmax")
    return lohi


def ensemble_summary_packed(x, percentiles=(25, 50, 75), n_grid=5000, col0=0, ncols=None, out=None, lohi_out=None):
    """One library call (``ertdiff_ensemble_summary``) for everything the path reports about the ensemble
    ``x (N, Q)`` (CUDA, float32 / float64) for the column window ``[col0, col0 + ncols)``: returns the float64 block
    ``(ncols, 5 + len(percentiles))`` with one record per column -- ``[mean, std, var, percentiles..., mode,
    mode grid index]`` (the moments are the array-dtype results widened; the index is integral).  The KDE grid spans
    the min / max of the WHOLE array (ECD.py:749-751).  ``out``: a preallocated contiguous float64 block with at
    least ``ncols`` rows to write into."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.dtype in _DT):
        raise _lib.ErtdiffError("ensemble_summary_packed takes a 2-D float32 / float64 CUDA tensor")
    x = x.contiguous()
    N, Q = x.shape
    ncols = Q - col0 if ncols is None else ncols
    nq = len(percentiles)
    rows = 5 + nq
    if out is None:
        out = torch.empty(ncols, rows, device=x.device, dtype=torch.float64)
    elif out.dtype != torch.float64 or out.dim() != 2 or out.size(1) != rows or out.size(0) < ncols or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 (>= ncols, 5 + len(percentiles)) block")
    qarr = (C.c_double * max(nq, 1))(*[float(q) for q in percentiles])
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ertdiff_ensemble_summary(
            _lib.ptr(x), _DT[x.dtype], N, Q, int(col0), int(ncols), qarr, nq, int(n_grid), _lib.ptr(lohi_out),
            _lib.ptr(out), rows, _lib.stream_ptr(x.device)), "ensemble_summary")
    return out


def summary_views(block, nq):
    """Named views of a packed summary block ``(Q, 5 + nq)``: ``mean, std, var, pct (nq, Q), mode, mode_index``
    (``mode_index`` as the integral float64 values of the block; ``.long()`` them where an index tensor is needed)."""
    return {"mean": block[:, 0], "std": block[:, 1], "var": block[:, 2], "pct": block[:, 3:3 + nq].t(),
            "mode": block[:, 3 + nq], "mode_index": block[:, 4 + nq], "packed": block}


def ensemble_summary(x, percentiles=(25, 50, 75), n_grid=5000):
    """Moments, percentiles and KDE mode of a CUDA tensor ``x (N, Q)`` as one step and one library call; returns
    float64 views of the packed block (see ``ensemble_summary_packed`` / ``summary_views``)."""
    return summary_views(ensemble_summary_packed(x, percentiles, n_grid), len(percentiles))


def ensemble_statistics(a, percentiles=(25, 50, 75), mode=True, n_grid=5000, device=None):
    """Everything ECD.py:747-762 + 867-872 computes for an ensemble of maps, in one call:
    ``{"mean","std","var","percentiles":{q: map},"mode","mode_index"}``."""
    out = ensemble_moments(a, device)
    out["percentiles"] = {q: ensemble_percentile(a, q, device) for q in percentiles}
    if mode:
        out["mode"], out["mode_index"] = ensemble_kde_mode(a, n_grid, None, device, True)
    return out


# ---------------------------------------------------------------------------------------------
# UQ calibration metrics (ECD.py:1089-1137 pooled over the parameters, ECD.py:1191-1214 per parameter)
def _trapezoid(y, x):
    # np.trapz(y, x, dx=...) (ECD.py:1099, 1107, 1112): with x given, dx is ignored
    d = np.diff(x)
    return float(np.sum(d * (y[1:] + y[:-1]) / 2.0))


def _calibration_scores(avg, prob):
    a_p = (avg >= prob).astype(np.int64)                                  # ECD.py:1089-1096
    accuracy = _trapezoid(a_p.astype(np.float64), prob)                   # ECD.py:1098-1100
    precision = 0.0 if accuracy == 0 else 1 - 2 * _trapezoid(a_p * (avg - prob), prob)   # ECD.py:1102-1109
    goodness = 1 - _trapezoid((3 * a_p - 2) * (avg - prob), prob)         # ECD.py:1111-1115
    return accuracy, precision, goodness


def uq_calibration(generated, true, n_prob=30, device=None):
    """The reference's uncertainty-calibration metrics for an ensemble ``generated (N, M, P)``
    against ``true (M, P)`` (ECD.py:1115-1137 and 1191-1214): for ``n_prob`` nested probability
    intervals p, the 2*n_prob ensemble percentiles ``(1 -/+ p)/2*100`` over the members and the
    coverage counts run on the device (one multi-quantile selection pass + one counting kernel);
    the 30-point trapezoid integrals are host arithmetic.  Returns numpy values:
    ``prob_array, avg_proportion, accuracy, precision, goodness`` and ``param_*`` per parameter."""
    t = generated if isinstance(generated, torch.Tensor) else np.asarray(generated)
    if t.ndim != 3:
        raise ValueError("generated must be (N, M, P)")
    N, M, P = t.shape
    tr = true if isinstance(true, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(true)))
    if tuple(tr.shape) != (M, P):
        raise ValueError(f"true must be ({M}, {P})")
    if P > 32:
        raise _lib.ErtdiffError("uq_calibration: at most 32 parameters")
    prob = np.linspace(0.01, 0.99, int(n_prob))                           # ECD.py:1119
    q = np.concatenate([(1 - prob) / 2 * 100, (1 + prob) / 2 * 100])      # ECD.py:1123-1126 (float64 scalars)
    flat, _, _ = _to_device(t.reshape(N, M * P), device)
    bounds = ensemble_percentile(flat, list(q))                           # (2K, M*P) float64 on the device
    truth = tr.to(device=flat.device, dtype=torch.float64).reshape(-1).contiguous()
    K = int(n_prob)
    counts = torch.empty(K, P + 1, device=flat.device, dtype=torch.int32)
    with torch.cuda.device(flat.device):
        _lib.check(_lib.load().ertdiff_interval_coverage(
            _lib.ptr(bounds[:K]), _lib.ptr(bounds[K:]), _lib.ptr(truth), K, M * P, P, _lib.ptr(counts),
            _lib.stream_ptr(flat.device)), "interval_coverage")
    c = counts.cpu().numpy().astype(np.float64)
    avg = c[:, 0] / (M * P)
    pavg = (c[:, 1:] / M).T.copy()                                        # (P, K)
    out = {"prob_array": prob, "avg_proportion": avg}
    out["accuracy"], out["precision"], out["goodness"] = _calibration_scores(avg, prob)
    scores = np.array([_calibration_scores(pavg[j], prob) for j in range(P)])
    out.update(param_avg_proportion=pavg, param_accuracy=scores[:, 0], param_precision=scores[:, 1],
               param_goodness=scores[:, 2])
    return out


def pack_rows_f64(rows, out):
    """``out[c, r] = float64(rows[r][c])``: pack 1-D CUDA row vectors (float32 / float64 / int64, equal length)
    into the first ``len(rows[0])`` records of the float64 block ``out (>= ncols, len(rows))`` with ONE launch
    (``ertdiff_pack_rows_f64``) instead of a cat and a conversion per row."""
    codes = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.int64: 2}
    rows = [r.contiguous() for r in rows]
    n, ncols = len(rows), rows[0].numel()
    if out.dtype != torch.float64 or out.dim() != 2 or out.size(1) != n or out.size(0) < ncols or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 (>= ncols, n_rows) block")
    if any(r.numel() != ncols or r.dtype not in codes for r in rows):
        raise ValueError("rows must be float32/float64/int64 vectors of equal length")
    ptrs = (C.c_void_p * n)(*[r.data_ptr() for r in rows])
    dts = (C.c_int32 * n)(*[codes[r.dtype] for r in rows])
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().ertdiff_pack_rows_f64(ptrs, dts, n, ncols, n, _lib.ptr(out), _lib.stream_ptr(out.device)),
                   "pack_rows_f64")
    return out


def argsort_stable(v):
    """``np.argsort(v, kind="stable")`` of a 1-D CUDA tensor on the device (ECD.py:786 ranks the members by
    their total WSSE); NaN sorts last, as in numpy."""
    v = v.contiguous()
    if v.dtype not in _DT:
        v = v.to(torch.float64)
    order = torch.empty(v.numel(), device=v.device, dtype=torch.int64)
    if v.numel():
        with torch.cuda.device(v.device):
            _lib.check(_lib.load().ertdiff_argsort_stable(_lib.ptr(v), _DT[v.dtype], v.numel(), _lib.ptr(order),
                                                          _lib.stream_ptr(v.device)), "argsort_stable")
    return order


def misfit_metrics(sim_data, observed, A=0.1, B=0.01, device=None):
    """Per-member data misfit of the simulated maps ``sim_data (N, L, C)`` against the observed map
    ``observed (L, C)`` (the conditional ERT sample), as the reference computes it inline:

    * ``wsse (N, C)``: ``WSSE_metric(A, B, sim_data[i][:, es], observed[:, es])`` for every member and
      survey (ECD.py:764-783), ``wsse_total (N)`` = ``wsse.sum(axis=1)`` and ``order`` = its argsort
      (ECD.py:785-786);
    * ``mse (N)``: ``mean_squared_error(observed.flatten(), sim_data[i].flatten())`` (ECD.py:927-930;
      with a single map, e.g. the ensemble mean or mode, this is ECD.py:939-940).

    Computed on the device in the arrays' dtype (float32 or float64) along numpy's pairwise-summation
    tree: the values are bit-identical to numpy's.  numpy in -> numpy out, CUDA tensors in -> CUDA tensors."""
    was_numpy = isinstance(sim_data, np.ndarray)
    sims = torch.from_numpy(np.ascontiguousarray(sim_data)) if was_numpy else sim_data
    obs = torch.from_numpy(np.ascontiguousarray(np.asarray(observed))) if not isinstance(observed, torch.Tensor) else observed
    if sims.dim() == 2:
        sims = sims[None]
    if sims.dim() != 3 or tuple(obs.shape) != tuple(sims.shape[1:]):
        raise ValueError("sim_data must be (N, L, C) and observed (L, C)")
    if sims.dtype not in _DT:
        sims = sims.to(torch.float64)
    if sims.device.type != "cuda":
        dev = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if was_numpy else None)
        if dev is None:
            raise _lib.ErtdiffError("misfit metrics run on CUDA only: pass CUDA tensors, numpy arrays, or device=")
        sims = sims.to(dev)
    sims = sims.contiguous()
    obs = obs.to(device=sims.device, dtype=sims.dtype).contiguous()
    N, L, C = sims.shape
    wsse = torch.empty(N, C, device=sims.device, dtype=sims.dtype)
    total = torch.empty(N, device=sims.device, dtype=sims.dtype)
    mse = torch.empty(N, device=sims.device, dtype=sims.dtype)
    with torch.cuda.device(sims.device):
        _lib.check(_lib.load().ertdiff_misfit_metrics(
            _lib.ptr(sims), _lib.ptr(obs), _DT[sims.dtype], N, L, C, float(A), float(B), _lib.ptr(wsse),
            _lib.ptr(total), _lib.ptr(mse), _lib.stream_ptr(sims.device)), "misfit_metrics")
    out = {"wsse": wsse, "wsse_total": total, "order": argsort_stable(total), "mse": mse}
    return {k: v.cpu().numpy() for k, v in out.items()} if was_numpy else out


def wasserstein_distance(u_values, v_values, device=None):
    """``scipy.stats.wasserstein_distance(u.flatten(), v.flatten())`` on the device, as the reference calls it
    for a simulated / ensemble-mean / ensemble-mode map against the observed map (ECD.py:860, 898-899).
    ``u_values`` may carry a leading member axis over ``v_values``' shape -- ``(N, L, C)`` against ``(L, C)``
    -- and then yields the ``N`` distances in one call.  float64 arithmetic as in scipy; scipy's last step is
    a BLAS dot product, so values agree to ~1e-12 relative, not bitwise.  numpy in -> float / numpy out."""
    was_numpy = not isinstance(u_values, torch.Tensor)
    u = torch.from_numpy(np.ascontiguousarray(np.asarray(u_values))) if was_numpy else u_values
    v = v_values if isinstance(v_values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(v_values)))
    batched = u.dim() == v.dim() + 1 and tuple(u.shape[1:]) == tuple(v.shape) and v.dim() >= 1
    if u.numel() == 0 or v.numel() == 0:
        raise ValueError("Distribution can't be empty.")          # scipy's message
    if u.dtype not in _DT:
        u = u.to(torch.float64)
    if u.device.type != "cuda":
        dev = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if was_numpy else None)
        if dev is None:
            raise _lib.ErtdiffError("wasserstein_distance runs on CUDA only: pass CUDA tensors, numpy arrays, or device=")
        u = u.to(dev)
    u = u.reshape(u.size(0), -1).contiguous() if batched else u.reshape(1, -1).contiguous()
    v = v.to(device=u.device, dtype=u.dtype).reshape(-1).contiguous()
    out = torch.empty(u.size(0), device=u.device, dtype=torch.float64)
    with torch.cuda.device(u.device):
        _lib.check(_lib.load().ertdiff_wasserstein_distance(
            _lib.ptr(u), _lib.ptr(v), _DT[u.dtype], u.size(0), u.size(1), v.numel(), _lib.ptr(out),
            _lib.stream_ptr(u.device)), "wasserstein_distance")
    if not batched:
        return float(out.item()) if was_numpy else out[0]
    return out.cpu().numpy() if was_numpy else out
