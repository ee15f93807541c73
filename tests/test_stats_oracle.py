"""The statistics restatements against numpy / scipy themselves (the third-party code the
reference calls at ECD.py:747-762, 867-872) and against the golden maps."""
import numpy as np
import pytest

from oracle import stats_oracle as so


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N", [2, 3, 16, 50, 255, 1024])
def test_moments_restatement_is_numpy(dtype, N):
    a = np.random.default_rng(N).lognormal(size=(N, 11)).astype(dtype)
    assert np.array_equal(so.seq_mean(a), np.mean(a, axis=0))
    assert np.array_equal(so.seq_var(a), np.var(a, axis=0))
    assert np.array_equal(so.seq_std(a), np.std(a, axis=0))


QS = [25, 50, 75, 0, 100, 2.5, 97.5, 33.3, np.float64(50.5), [2.5, 97.5],
      list((1 + np.linspace(0.01, 0.99, 30)) / 2 * 100)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N", [2, 5, 50, 256, 1000, 8192])
def test_percentile_restatement_is_numpy(dtype, N):
    a = np.random.default_rng(N + 7).normal(size=(N, 5)).astype(dtype)
    for q in QS:
        ref = np.percentile(a, q, axis=0)
        got = so.percentile_linear(a, q)
        assert ref.dtype == got.dtype, (q, ref.dtype, got.dtype)
        assert np.array_equal(ref, got), q


def test_percentile_nan_column():
    a = np.random.default_rng(0).normal(size=(20, 4))
    a[3, 2] = np.nan
    ref = np.percentile(a, [10, 50], axis=0)
    got = so.percentile_linear(a, [10, 50])
    assert np.array_equal(ref, got, equal_nan=True)


def test_kde_closed_form_argmax_is_scipy():
    rng = np.random.default_rng(4)
    a = rng.lognormal(size=(50, 20))
    grid = so.kde_grid(a, 5000)
    _, idx = so.kde_mode(a, grid)
    _, idx_sp, pdfs = so.kde_mode_scipy(a, grid)
    same = idx == idx_sp
    # near-tie rule: a differing index is tolerated only if scipy's pdf at
    # the two indices agrees to 1e-13 relative
    for j in np.nonzero(~same)[0]:
        p = pdfs[:, j]
        assert abs(p[idx[j]] - p[idx_sp[j]]) <= 1e-13 * p[idx_sp[j]]
    assert same.mean() >= 0.9


def test_golden_maps(golden):
    g = golden("stats_maps.npz")
    sim = g["sim"]
    a = sim.reshape(sim.shape[0], -1)
    assert np.array_equal(so.seq_mean(a).reshape(sim.shape[1:]), g["mean"])
    assert np.array_equal(so.seq_std(a).reshape(sim.shape[1:]), g["std"])
    assert np.array_equal(so.seq_var(a).reshape(sim.shape[1:]), g["var"])
    for q, key in ((25, "p25"), (50, "p50"), (75, "p75")):
        assert np.array_equal(so.percentile_linear(a, q).reshape(sim.shape[1:]), g[key])
    ci = so.percentile_linear(a.astype(np.float32), [2.5, 97.5])
    assert ci.dtype == np.float64 and np.array_equal(ci.reshape(g["ci95"].shape), g["ci95"])
    mode, idx = so.kde_mode(sim)
    assert np.array_equal(idx, g["mode_index"])
    assert np.array_equal(mode, g["mode"])


@pytest.mark.parametrize("N", [50, 256, 2048, 20000])
def test_kde_coarse_to_fine_rule_never_skips_a_candidate(N):
    # the CUDA scan evaluates every s-th grid point first and skips intervals that provably hold no candidate of the
    # float64 selection (csrc/stats.cuh); this checks the RULE on seeded samples: no skipped point reaches the
    # candidates' threshold, and a useful share of the grid is skipped
    rng = np.random.default_rng(N)
    samples = {
        "normal": rng.normal(0.0, 270.0, N),
        "bimodal": np.concatenate([rng.normal(-2.0, 0.3, N // 2), rng.normal(1.5, 0.6, N - N // 2)]),
        "lognormal": rng.lognormal(0.0, 0.8, N),
        "uniform": rng.uniform(-1.0, 1.0, N),
        "ties": np.concatenate([rng.normal(0.0, 1.0, N - 5), np.full(5, 0.25)]),
        "two far clusters": np.concatenate([rng.normal(0.0, 1.0, N - N // 4), rng.normal(40.0, 0.05, N // 4)]),
    }
    fractions = []
    for name, col in samples.items():
        span = col.max() - col.min()
        grid = np.linspace(col.min() - 0.3 * span, col.max() + 0.2 * span, 5000)      # the common grid is wider than a column
        s, frac, worst = so.kde_coarse_to_fine_check(col, grid)
        assert worst < 1.0, (name, s, worst)
        if s > 1:
            fractions.append(frac)
    assert fractions and np.median(fractions) < 0.5
