"""numpy restatement of the reference's ensemble statistics.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference computes its ensemble maps inline with numpy / scipy
(``ECD.py:747-762`` KDE mode, ``ECD.py:867-872`` mean/std/var/percentiles, further
percentile call sites ``ECD.py:612, 1126-1127, 1199-1200``).  The arithmetic lives in
third-party code that is not under ``/root/reference``: numpy 2.3.5
(``numpy/lib/_function_base_impl.py``: ``_quantile``, ``_lerp``; ``numpy/_core/_methods.py``:
``_mean``, ``_var``) and scipy 1.18.1 (``scipy/stats/_kde.py`` + compiled
``gaussian_kernel_estimate``).  Both libraries are installed wherever the tests run, so
each restated function here is pinned by calling the library itself on the same input
(``tests/test_stats_oracle.py``); the restatements exist to spell out the exact operation
order the CUDA kernels must reproduce, element by element.

Members are on axis 0 throughout: ``a`` has shape ``(N, Q)``.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- moments
def seq_mean(a: np.ndarray) -> np.ndarray:
    """``np.mean(a, axis=0)`` (ECD.py:867): for an axis-0 reduction of a C-contiguous array
    numpy adds row after row into the output, i.e. a plain left-to-right sum per column in
    the array's dtype, then one true-divide by N."""
    acc = a[0].copy()
    for i in range(1, a.shape[0]):
        acc = acc + a[i]
    return acc / a.dtype.type(a.shape[0])


def seq_var(a: np.ndarray) -> np.ndarray:
    """``np.var(a, axis=0)`` (ECD.py:869), ddof=0: mean as above, then the left-to-right sum
    of ``(a_i - mean) * (a_i - mean)`` (separate multiply and add, no FMA), divided by N."""
    m = seq_mean(a)
    d = a[0] - m
    acc = d * d
    for i in range(1, a.shape[0]):
        d = a[i] - m
        acc = acc + d * d
    return acc / a.dtype.type(a.shape[0])


def seq_std(a: np.ndarray) -> np.ndarray:
    """``np.std(a, axis=0)`` (ECD.py:868) = sqrt of the above."""
    return np.sqrt(seq_var(a))


# --------------------------------------------------------------------------- quantiles
def percentile_index_dtype(a_dtype, q) -> np.dtype:
    """The dtype numpy does the index arithmetic in: a python int/float
    ``q`` is weakly typed and adopts the array's float dtype; a numpy float64 scalar, list or
    array ``q`` stays float64."""
    if isinstance(q, (int, float)) and not isinstance(q, np.generic):
        return np.dtype(a_dtype) if np.issubdtype(a_dtype, np.floating) else np.dtype(np.float64)
    return np.dtype(np.float64)


def lerp_linear(A, Bv, gamma, out_dtype):
    """numpy ``_lerp``: ``d = Bv - A`` in the ARRAY's dtype; then, in the result dtype,
    ``A + d*gamma`` and, where ``gamma >= 0.5``, ``Bv - d*(1-gamma)``; separate multiply and
    add/subtract, no FMA."""
    d = (Bv - A).astype(out_dtype)
    r = A.astype(out_dtype) + d * gamma
    alt = Bv.astype(out_dtype) - d * (1 - gamma)
    return np.where(gamma >= 0.5, alt, r)


def percentile_linear(a: np.ndarray, q, index_dtype=None) -> np.ndarray:
    """``np.percentile(a, q, axis=0)`` with the default ``method='linear'``
    (ECD.py:870-872, 612, 1126-1127, 1199-1200).

    ``q`` scalar -> result shape ``(Q,)``; ``q`` sequence -> ``(len(q), Q)``.
    ``quant = q/100``; ``vi = (N-1)*quant``; ``lo = floor(vi)``, ``hi = lo+1`` (both ``N-1`` when
    ``vi >= N-1``), ``gamma = vi - lo``, all in ``index_dtype``; the result dtype is
    ``result_type(a, gamma)``.  Columns containing NaN give NaN.
    """
    N = a.shape[0]
    idt = np.dtype(index_dtype) if index_dtype is not None else percentile_index_dtype(a.dtype, q)
    scalar = np.ndim(q) == 0
    qs = np.atleast_1d(np.asarray(q, dtype=idt))
    quant = qs / idt.type(100)
    s = np.sort(a, axis=0)
    has_nan = np.isnan(a).any(axis=0)
    out_dt = np.result_type(a.dtype, idt)
    out = np.empty((len(qs),) + a.shape[1:], dtype=out_dt)
    for k, qq in enumerate(quant):
        vi = idt.type(N - 1) * qq
        lo_f = np.floor(vi)
        gamma = vi - lo_f                    # index dtype
        lo = int(lo_f)
        hi = lo + 1
        if vi >= N - 1:                      # both indexes are the last element
            lo = hi = N - 1
        if vi < 0:
            lo = hi = 0
        r = lerp_linear(s[lo], s[hi], gamma, out_dt)
        out[k] = np.where(has_nan, np.nan, r)
    return out[0] if scalar else out


# --------------------------------------------------------------------------- KDE mode
def kde_grid(a: np.ndarray, n_grid: int = 5000) -> np.ndarray:
    """ECD.py:749-751: ``np.linspace(global_min, global_max, 5000)`` over the whole array."""
    return np.linspace(np.min(a), np.max(a), n_grid)


def kde_pdf_closed_form(col: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """What ``scipy.stats.gaussian_kde(col)(grid)`` evaluates (ECD.py:758-759), written out:
    Scott bandwidth ``h^2 = var_ddof1(col) * N^(-2/5)``;
    ``pdf(g) = sum_i exp(-(g - x_i)^2 / (2 h^2)) / (N sqrt(2 pi h^2))``, float64, the sum taken
    over members in order.  Values agree with scipy's compiled kernel to ~1e-15 relative but
    not bit for bit (scipy whitens with a Cholesky factor first); the argmax index is the
    contract."""
    col = np.asarray(col, dtype=np.float64)
    N = col.shape[0]
    h2 = np.var(col, ddof=1) * float(N) ** (-0.4)
    acc = np.zeros_like(grid, dtype=np.float64)
    for xi in col:
        acc = acc + np.exp(-((grid - xi) ** 2) / (2.0 * h2))
    return acc / (N * np.sqrt(2.0 * np.pi * h2))


def kde_mode(a: np.ndarray, grid: np.ndarray | None = None, n_grid: int = 5000):
    """ECD.py:747-762: per column, ``grid[argmax(pdf)]`` (first maximum).  Returns
    ``(mode_values (Q,), argmax_index (Q,) int64)``."""
    a2 = a.reshape(a.shape[0], -1)
    if grid is None:
        grid = kde_grid(a2, n_grid)
    idx = np.empty(a2.shape[1], dtype=np.int64)
    for j in range(a2.shape[1]):
        idx[j] = int(np.argmax(kde_pdf_closed_form(a2[:, j], grid)))
    return grid[idx].reshape(a.shape[1:]), idx.reshape(a.shape[1:])


def kde_mode_scipy(a: np.ndarray, grid: np.ndarray | None = None, n_grid: int = 5000):
    """The reference's own call sequence (ECD.py:753-762) through scipy, for pinning."""
    from scipy import stats
    a2 = a.reshape(a.shape[0], -1)
    if grid is None:
        grid = kde_grid(a2, n_grid)
    idx = np.empty(a2.shape[1], dtype=np.int64)
    pdfs = []
    for j in range(a2.shape[1]):
        vals = stats.gaussian_kde(a2[:, j])(grid)
        idx[j] = int(np.argmax(vals))
        pdfs.append(vals)
    return grid[idx].reshape(a.shape[1:]), idx.reshape(a.shape[1:]), np.stack(pdfs, axis=1)


def kde_coarse_to_fine_check(col: np.ndarray, grid: np.ndarray, max_stride: int = 32, tol: float = 1e-3):
    """Numpy restatement of the CUDA scan's coarse-to-fine skip rule (csrc/stats.cuh, "coarse-to-fine scan"), used by the
    CPU tests to check the RULE, not the kernel: every ``s``-th grid point is evaluated; an interval between two
    evaluated points is skipped when ``max(S(c), S(c')) + 0.17329 N (sc s step)^2 < M (1 - 2 tol)``, ``M`` the largest
    coarse value.  Returns ``(stride, evaluated_fraction, worst)`` with ``worst`` = the largest skipped value divided by
    the candidates' threshold ``max(S) (1 - tol)``: the rule is sound iff ``worst < 1`` (and the argmax is never
    skipped)."""
    col = np.asarray(col, dtype=np.float64)
    N, G = col.shape[0], grid.shape[0]
    step = (grid[-1] - grid[0]) / (G - 1)
    h = np.sqrt(np.var(col, ddof=1)) * float(N) ** (-0.2)
    s = 1
    while 2 * s <= max_stride and 6.0 * s * step <= h:
        s *= 2
    sc = np.sqrt(np.log2(np.e) / (2.0 * h * h))
    S = np.zeros(G)
    for i in range(0, N, 256):
        S += np.exp2(-((grid[:, None] - col[None, i:i + 256]) * sc) ** 2).sum(axis=1)
    if s == 1:
        return 1, 1.0, 0.0
    cidx = np.arange(0, G, s)
    if cidx[-1] != G - 1:
        cidx = np.append(cidx, G - 1)
    cv = S[cidx]
    M = cv.max()
    thr = M * (1.0 - 2.0 * tol) - 0.17329 * N * (sc * s * step) ** 2 * 1.001
    evaluated, worst = len(cidx), 0.0
    for k in range(len(cidx) - 1):
        a, b = cidx[k], cidx[k + 1]
        if max(cv[k], cv[k + 1]) < thr:
            if b - a > 1:
                worst = max(worst, S[a + 1:b].max() / (S.max() * (1.0 - tol)))
        else:
            evaluated += b - a - 1
    return s, evaluated / G, worst


# ---------------------------------------------------------------------------------------------
# UQ calibration metrics (ECD.py:1089-1137 for all parameters pooled, ECD.py:1191-1214 per
# parameter).  `generated` is (N realisations, M conditions, P) -- Uncertainty_params.npy at
# ECD.py:1082 -- and `true` is (M, P).
def _trapezoid(y, x):
    # np.trapz(y, x, dx=...) at ECD.py:1099, 1107, 1112: with x given, dx is ignored
    y = np.asarray(y, dtype=np.float64)
    d = np.diff(np.asarray(x, dtype=np.float64))
    return float(np.sum(d * (y[1:] + y[:-1]) / 2.0))


def avg_prop_indicator(avg_proportion, prob_array):          # ECD.py:1089-1096
    return np.array([1 if avg_proportion[i] >= prob_array[i] else 0 for i in range(prob_array.shape[0])])


def accuracy_score(a_p, prob_array):                         # ECD.py:1098-1100
    return _trapezoid(a_p, prob_array)


def precision_score(accuracy, avg_proportion, prob_array, a_p):   # ECD.py:1102-1109
    if accuracy == 0:
        return 0.0
    return 1 - 2 * _trapezoid(a_p * (avg_proportion - prob_array), prob_array)


def goodness_score(a_p, avg_proportion, prob_array):         # ECD.py:1111-1115
    return 1 - _trapezoid((3 * a_p - 2) * (avg_proportion - prob_array), prob_array)


def coverage_curve(dist, true, prob_array):
    """avg_proportion[k] = mean over the trailing axes of (low < true) & (true <= upp), with
    low/upp = np.percentile(dist, (1 -/+ p_k)/2 * 100, axis=0) (ECD.py:1121-1132, 1195-1206; the
    percentile argument is an np.float64, so the index arithmetic runs in float64)."""
    avg = np.zeros(len(prob_array))
    for k, p in enumerate(prob_array):
        p_low, p_upp = (1 - p) / 2, (1 + p) / 2
        low = np.percentile(dist, p_low * 100, axis=0)
        upp = np.percentile(dist, p_upp * 100, axis=0)
        avg[k] = np.mean(((low < true) & (true <= upp)).astype(int))
    return avg


def uq_calibration(generated, true, n_prob=30):
    """Returns {"prob_array", "avg_proportion", "accuracy", "precision", "goodness"} for the pooled
    parameters and the same keys with a ``param_`` prefix holding one entry per parameter."""
    prob = np.linspace(0.01, 0.99, n_prob)                   # ECD.py:1119
    out = {"prob_array": prob}
    avg = coverage_curve(generated, true, prob)
    a_p = avg_prop_indicator(avg, prob)
    acc = accuracy_score(a_p, prob)
    out.update(avg_proportion=avg, accuracy=acc, precision=precision_score(acc, avg, prob, a_p),
               goodness=goodness_score(a_p, avg, prob))
    P = generated.shape[2]
    pavg = np.zeros((P, n_prob))
    pacc, pprec, pgood = np.zeros(P), np.zeros(P), np.zeros(P)
    for j in range(P):                                        # ECD.py:1191-1214
        pavg[j] = coverage_curve(generated[:, :, j], true[:, j], prob)
        a_p = avg_prop_indicator(pavg[j], prob)
        pacc[j] = accuracy_score(a_p, prob)
        pprec[j] = precision_score(pacc[j], pavg[j], prob, a_p)
        pgood[j] = goodness_score(a_p, pavg[j], prob)
    out.update(param_avg_proportion=pavg, param_accuracy=pacc, param_precision=pprec, param_goodness=pgood)
    return out


# ---- per-member misfit metrics (ECD.py:764-785, 927-940) -------------------------------------------
def pairwise_sum(a: np.ndarray):
    """numpy's add-reduction of a contiguous 1-D float array, restated (numpy 2.3
    ``_core/src/umath/loops_utils.h.src``, ``@TYPE@_pairwise_sum``): n < 8 in order from 0; n <= 128
    through 8 interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the n % 8
    tail in order; longer arrays split at n/2 rounded down to a multiple of 8.  Pure-Python loops: for
    small cases only (tests check it against ``np.add.reduce`` itself)."""
    T = a.dtype.type
    n = len(a)
    if n < 8:
        r = T(0)
        for v in a:
            r = T(r + v)
        return r
    if n <= 128:
        r = [a[k] for k in range(8)]
        body = n - n % 8
        for i in range(8, body, 8):
            for k in range(8):
                r[k] = T(r[k] + a[i + k])
        res = T(T(T(r[0] + r[1]) + T(r[2] + r[3])) + T(T(r[4] + r[5]) + T(r[6] + r[7])))
        for i in range(body, n):
            res = T(res + a[i])
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return T(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


def wsse_metric(A, B, predictions, observations):               # ECD.py:764-770
    sd = A * np.abs(observations) + B
    wse = (predictions - observations) ** 2 / (sd) ** 2
    return np.average(wse), wse


def misfit_metrics(sim_data, observed, A=0.1, B=0.01):
    """The reference's inline misfit loops: WSSE per (member, survey) and its per-member total and
    ranking (ECD.py:773-786), MSE per member over the flattened maps (ECD.py:927-930 -- sklearn's
    ``mean_squared_error`` is ``np.average((y_true - y_pred) ** 2)`` for 1-D inputs)."""
    N, L, C = sim_data.shape
    wsse = np.array([[wsse_metric(A, B, sim_data[i][:, es], observed[:, es])[0] for es in range(C)]
                     for i in range(N)])
    total = wsse.sum(axis=1)
    yt = observed.flatten()
    mse = np.array([np.average((yt - sim_data[i].flatten()) ** 2) for i in range(N)])
    return {"wsse": wsse, "wsse_total": total, "order": np.argsort(total, kind="stable"), "mse": mse}


def wasserstein_distance(u_values, v_values):
    """scipy 1.18 ``stats.wasserstein_distance`` = ``_cdf_distance(1, ...)`` (scipy/stats/_stats_py.py),
    restated with numpy for unweighted samples; the reference calls it at ECD.py:860, 898-899."""
    u = np.asarray(u_values, dtype=float)
    v = np.asarray(v_values, dtype=float)
    u_sorted, v_sorted = np.sort(u), np.sort(v)
    all_values = np.concatenate((u, v))
    all_values.sort(kind="mergesort")
    deltas = np.diff(all_values)
    u_cdf = u_sorted.searchsorted(all_values[:-1], "right") / u.size
    v_cdf = v_sorted.searchsorted(all_values[:-1], "right") / v.size
    return np.vecdot(np.abs(u_cdf - v_cdf), deltas)
