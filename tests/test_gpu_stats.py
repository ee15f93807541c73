"""GPU parity of the ensemble statistics against numpy / scipy (the reference's own calls,
ECD.py:747-762, 867-872) and the golden maps.  Moments and percentiles: bit-exact.  KDE mode:
exact argmax index, with the counted near-tie rule of SURVEY.md §8 a7."""
import numpy as np
import pytest
import torch

import ertdiff_b200 as eb
from oracle import stats_oracle as so

pytestmark = pytest.mark.gpu


def same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N,Q", [(2, 2), (3, 29), (16, 29), (50, 4693 * 14), (256, 29), (1000, 257), (8192, 29), (1024, 3001)])
def test_moments_bit_exact(cuda_dev, dtype, N, Q):
    a = np.random.default_rng(N * 7 + Q).lognormal(size=(N, Q)).astype(dtype)
    m = eb.ensemble_moments(a)
    assert same(m["mean"], np.mean(a, axis=0))
    assert same(m["var"], np.var(a, axis=0))
    assert same(m["std"], np.std(a, axis=0))


QS = [25, 50, 75, 0, 100, 2.5, 97.5, 33.3, np.float64(50.5), [2.5, 97.5],
      list((1 - np.linspace(0.01, 0.99, 30)) / 2 * 100) + list((1 + np.linspace(0.01, 0.99, 30)) / 2 * 100)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N,Q", [(2, 3), (5, 29), (16, 29), (33, 9), (50, 1000), (64, 300), (65, 17), (255, 64), (256, 29), (513, 70),
                                 (1000, 33), (1024, 29), (1025, 40), (1500, 40), (2048, 33), (3000, 1500), (4096, 7), (8192, 29)])
def test_percentiles_bit_exact(cuda_dev, dtype, N, Q, monkeypatch):
    # three kernels, all bit-identical to numpy: one CTA per column group sorting in shared memory; one warp per column
    # sorting in registers (short columns of many-column arrays -- forced here for every N <= 1024 as well); longer
    # columns of few-column arrays: sorted runs + exact multi-run selection
    if N <= 1024:
        a = np.random.default_rng(N + Q).normal(size=(N, Q)).astype(dtype)
        monkeypatch.setenv("ERTDIFF_PCTL_WARP", "1")
        for q in QS:
            assert same(eb.ensemble_percentile(a, q), np.percentile(a, q, axis=0)), ("warp kernel", q)
        monkeypatch.delenv("ERTDIFF_PCTL_WARP")
    a = np.random.default_rng(N + Q).normal(size=(N, Q)).astype(dtype)
    for q in QS:
        ref = np.percentile(a, q, axis=0)
        got = eb.ensemble_percentile(a, q)
        assert same(got, ref), (q, ref.dtype, got.dtype, np.abs(got - ref).max())


def test_percentiles_edge_cases(cuda_dev):
    rng = np.random.default_rng(1)
    a = rng.normal(size=(1, 5))                       # one member
    assert same(eb.ensemble_percentile(a, 50), np.percentile(a, 50, axis=0))
    a = rng.normal(size=(20, 6))
    a[3, 2] = np.nan                                  # NaN column -> NaN, others untouched
    assert same(eb.ensemble_percentile(a, [10, 50]), np.percentile(a, [10, 50], axis=0))
    a = np.repeat(rng.normal(size=(1, 9)), 40, axis=0)   # all members equal (ties)
    assert same(eb.ensemble_percentile(a, [0, 37.5, 100]), np.percentile(a, [0, 37.5, 100], axis=0))
    a = rng.integers(0, 4, size=(64, 11)).astype(np.float32)   # heavy ties
    assert same(eb.ensemble_percentile(a, 30), np.percentile(a, 30, axis=0))
    a = rng.normal(size=(50, 12, 3))                  # trailing shape is kept
    assert eb.ensemble_percentile(a, 25).shape == (12, 3)
    assert eb.ensemble_percentile(a, [25, 75]).shape == (2, 12, 3)
    with pytest.raises(ValueError):
        eb.ensemble_percentile(a, 101)
    t = torch.from_numpy(a).to(cuda_dev)              # CUDA tensor in -> CUDA tensor out
    r = eb.ensemble_percentile(t, 25)
    assert r.is_cuda and same(r.cpu().numpy(), np.percentile(a, 25, axis=0))
    a = np.where(rng.random((30, 8)) < 0.1, np.inf, rng.normal(size=(30, 8)))   # infinities
    assert same(eb.ensemble_percentile(a, [5, 50, 99]), np.percentile(a, [5, 50, 99], axis=0))


def test_golden_maps(cuda_dev, golden):
    g = golden("stats_maps.npz")
    sim = g["sim"]
    st = eb.ensemble_statistics(sim, percentiles=(25, 50, 75))
    assert same(st["mean"], g["mean"]) and same(st["std"], g["std"]) and same(st["var"], g["var"])
    for q, key in ((25, "p25"), (50, "p50"), (75, "p75")):
        assert same(st["percentiles"][q], g[key])
    assert same(eb.ensemble_percentile(sim.astype(np.float32), [2.5, 97.5]), g["ci95"])
    assert np.array_equal(st["mode_index"], g["mode_index"])
    assert np.array_equal(st["mode"], g["mode"])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N", [16, 50, 256, 1000, 3000])      # 3000 x 24 values: beyond the fused small-ensemble launch
def test_kde_mode_index_is_scipy(cuda_dev, dtype, N):
    a = np.random.default_rng(N).lognormal(size=(N, 24)).astype(dtype)
    # the reference's maps are float64 (ECD.py:716); float32 input is defined as promoted first
    a64 = a.astype(np.float64)
    grid = so.kde_grid(a64, 5000)
    mode, idx = eb.ensemble_kde_mode(a, 5000, return_index=True)
    _, idx_sp, pdfs = so.kde_mode_scipy(a64, grid)
    differ = np.nonzero(idx != idx_sp)[0]
    for j in differ:                                   # near-tie rule, counted
        p = pdfs[:, j]
        assert abs(p[idx[j]] - p[idx_sp[j]]) <= 1e-13 * p[idx_sp[j]], (j, idx[j], idx_sp[j])
    assert len(differ) <= 1, f"near-tie rule invoked {len(differ)} times out of 24"
    assert np.array_equal(mode, grid[idx])             # the grid itself is np.linspace, bit for bit


def test_kde_mode_full_grid_pixels_property(cuda_dev):
    # reference-sized map subsample: mode lies inside the data range and moves with a shift
    rng = np.random.default_rng(8)
    a = rng.lognormal(size=(50, 2000))
    mode, idx = eb.ensemble_kde_mode(a, 5000, return_index=True)
    assert mode.min() >= a.min() and mode.max() <= a.max()
    lo, hi = a.min(), a.max()
    mode2, idx2 = eb.ensemble_kde_mode(a + 3.0, 5000, grid_range=(lo + 3.0, hi + 3.0), return_index=True)
    assert np.abs(idx2 - idx).max() <= 1               # same grid, shifted data: same argmax (+-1 ulp effects)


def test_statistics_of_large_ensembles_properties(cuda_dev):
    # BASELINE config 4 size: 8192 members; selection must return members of the column
    a = torch.randn(8192, 29, device=cuda_dev, generator=torch.Generator(cuda_dev).manual_seed(0))
    q = eb.ensemble_percentile(a, [0, 50, 100])
    assert torch.equal(q[0].float(), a.min(dim=0).values) and torch.equal(q[2].float(), a.max(dim=0).values)
    srt = a.sort(dim=0).values
    assert torch.equal(q[1].float(), ((srt[4095].double() + (srt[4096] - srt[4095]).double() * 0.5)).float())
    perm = a[torch.randperm(8192, device=cuda_dev)]
    assert torch.equal(eb.ensemble_percentile(perm, [0, 50, 100]), q)   # order of members is irrelevant


def test_kde_fused_and_staged_paths_agree(cuda_dev):
    # the same data through the one-launch small-ensemble kernel (range from the data) and through the
    # staged kernels (explicit range): same grid, same argmax
    a = np.random.default_rng(12).normal(size=(400, 29)).astype(np.float32) * np.linspace(0.5, 30, 29, dtype=np.float32)
    m1, i1 = eb.ensemble_kde_mode(a, 5000, return_index=True)
    m2, i2 = eb.ensemble_kde_mode(a, 5000, grid_range=(float(a.min()), float(a.max())), return_index=True)
    assert np.array_equal(i1, i2) and np.array_equal(m1, m2)
    # repeated calls reuse the self-cleaning tickets
    m3, i3 = eb.ensemble_kde_mode(a, 5000, return_index=True)
    assert np.array_equal(i1, i3)


def test_sharded_statistics_single_rank_equals_plain(cuda_dev):
    x = torch.randn(700, 29, device=cuda_dev, generator=torch.Generator(cuda_dev).manual_seed(3)) * 5
    qs = (2.5, 50.0, 97.5)
    sh = eb.parallel.sharded_statistics(x, qs, 1000)
    st = eb.ensemble_statistics(x, percentiles=(), n_grid=1000)
    pct = eb.ensemble_percentile(x, list(qs))
    for k in ("mean", "std", "var"):
        assert torch.equal(sh[k].to(st[k].dtype), st[k])
    assert torch.equal(sh["pct"], pct.double()) and torch.equal(sh["mode"], st["mode"])
    assert torch.equal(sh["mode_index"].long(), st["mode_index"])


# ---- ensembles beyond one CTA's shared memory (round 2: no size cap) -------------------------------
@pytest.fixture
def env_override(monkeypatch):
    def set_(name, value):
        monkeypatch.setenv(name, str(value))
    return set_


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N,Q", [(40000, 29), (160000, 29), (33000, 3), (70001, 64)])
def test_percentiles_any_size_bit_exact(cuda_dev, dtype, N, Q):
    rng = np.random.default_rng(N + Q)
    a = rng.normal(size=(N, Q)).astype(dtype)
    a[:, 0] = np.round(a[:, 0] * 4) / 4                  # heavy ties
    if Q > 2:
        a[rng.integers(0, N, 5), 2] = np.inf
    for q in (50, [2.5, 25.0, 50.0, 75.0, 97.5], [0, 100, 33.3], np.float64(12.5)):
        ref = np.percentile(a, q, axis=0)
        got = eb.ensemble_percentile(a, q)
        assert same(got, ref), (q, np.abs(got - ref).max())
    a[N // 2, 1] = np.nan                                # a NaN column is NaN, its neighbours untouched
    assert same(eb.ensemble_percentile(a, [10, 50]), np.percentile(a, [10, 50], axis=0))
    b = np.ascontiguousarray(a[:, [0, 2]])               # C order: numpy then adds row after row, as the kernel does
    m = eb.ensemble_moments(b)                           # (a 1-column or Fortran-ordered array is summed pairwise instead)
    assert same(m["mean"], np.mean(b, axis=0)) and same(m["std"], np.std(b, axis=0))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("run_len", [2, 16, 256])
@pytest.mark.parametrize("N,Q", [(1, 3), (2, 2), (5, 29), (17, 1), (255, 7), (1000, 33), (4097, 5)])
def test_percentiles_run_path_small_runs(cuda_dev, env_override, dtype, run_len, N, Q):
    # the sorted-runs / multi-run selection path forced onto small columns: ragged last runs, one-member runs,
    # more runs than lanes (4097 / 16 = 257 runs), ties, infinities, every query form
    env_override("ERTDIFF_PCTL_RUN_LEN", run_len)
    if N / run_len > 4096:
        pytest.skip("more runs than the selection kernel's window table holds")
    rng = np.random.default_rng(N * 31 + Q + run_len)
    a = rng.integers(-3, 4, size=(N, Q)).astype(dtype) if (N + Q) % 2 else rng.normal(size=(N, Q)).astype(dtype)
    if N > 4:
        a[1, 0], a[3, 0] = -np.inf, np.inf
        a[2, Q - 1] = -0.0
    for q in QS:
        ref = np.percentile(a, q, axis=0)
        got = eb.ensemble_percentile(a, q)
        assert same(got, ref), (q, ref.dtype, got.dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N", [30000, 70000])
def test_kde_mode_any_size_index_is_scipy(cuda_dev, dtype, N):
    a = np.random.default_rng(N).lognormal(size=(N, 6)).astype(dtype)
    a64 = a.astype(np.float64)
    grid = so.kde_grid(a64, 2000)
    mode, idx = eb.ensemble_kde_mode(a, 2000, return_index=True)
    _, idx_sp, pdfs = so.kde_mode_scipy(a64, grid)
    for j in np.nonzero(idx != idx_sp)[0]:
        p = pdfs[:, j]
        assert abs(p[idx[j]] - p[idx_sp[j]]) <= 1e-13 * p[idx_sp[j]], (j, idx[j], idx_sp[j])
    assert (idx != idx_sp).sum() <= 1
    assert np.array_equal(mode, grid[idx])


@pytest.mark.parametrize("tile", [64, 448, 4096])
def test_kde_tiled_path_equals_resident_path(cuda_dev, env_override, tile):
    # members streamed through shared memory in tiles vs the whole column resident: same argmax, and the fp32
    # scan is the same sum in the same order (tiles are multiples of its 64-member blocks)
    rng = np.random.default_rng(tile)
    a = (rng.normal(size=(3000, 29)) * np.linspace(0.5, 30, 29)).astype(np.float32)
    a[:, 5] = 2.0                                        # constant column: no KDE (NaN, -1)
    lohi = (float(a.min()), float(a.max()))
    m1, i1 = eb.ensemble_kde_mode(a, 5000, grid_range=lohi, return_index=True)
    env_override("ERTDIFF_KDE_TILE", tile)
    m2, i2 = eb.ensemble_kde_mode(a, 5000, grid_range=lohi, return_index=True)
    assert np.array_equal(i1, i2) and np.array_equal(m1, m2, equal_nan=True)
    assert i2[5] == -1 and np.isnan(m2[5])
    b = rng.lognormal(size=(2500, 300))                  # many columns: one CTA walks its column's whole grid
    env_override("ERTDIFF_KDE_TILE", 0)
    r1 = eb.ensemble_kde_mode(b, 1000, grid_range=(b.min(), b.max()), return_index=True)[1]
    env_override("ERTDIFF_KDE_TILE", tile)
    r2 = eb.ensemble_kde_mode(b, 1000, grid_range=(b.min(), b.max()), return_index=True)[1]
    assert np.array_equal(r1, r2)


def test_statistics_full_size_151k_members(cuda_dev):
    # 8 GPUs x 18,944 members (one full wave of 128-member tiles each): the gathered ensemble
    N = 8 * 18944
    a = torch.randn(N, 29, device=cuda_dev, generator=torch.Generator(cuda_dev).manual_seed(1)) * 3 + 1
    q = eb.ensemble_percentile(a, [0.0, 50.0, 100.0])
    assert torch.equal(q[0].float(), a.min(dim=0).values) and torch.equal(q[2].float(), a.max(dim=0).values)
    srt = a.sort(dim=0).values
    d = (srt[N // 2] - srt[N // 2 - 1]).double()         # numpy: subtract in the array's dtype, then promote
    assert torch.equal(q[1], srt[N // 2].double() - d * 0.5)      # gamma = 0.5 exactly: the `b - d*(1-gamma)` branch
    mode, idx = eb.ensemble_kde_mode(a, 5000, return_index=True)
    assert ((mode - 1.0).abs() < 0.5).all()              # N(1, 3): the KDE mode sits near the mean
    m = eb.ensemble_moments(a)
    assert ((m["mean"] - 1.0).abs() < 0.05).all() and ((m["std"] - 3.0).abs() < 0.05).all()


def test_argsort_stable_is_numpy(cuda_dev):
    rng = np.random.default_rng(5)
    for n in (1, 2, 50, 1025, 5000):
        v = np.round(rng.normal(size=n), 1)              # ties
        if n > 10:
            v[7] = np.nan
            v[3] = -np.inf
        for dt in (np.float32, np.float64):
            got = eb.stats.argsort_stable(torch.from_numpy(v.astype(dt)).to(cuda_dev)).cpu().numpy()
            assert np.array_equal(got, np.argsort(v.astype(dt), kind="stable")), (n, dt)


@pytest.mark.parametrize("N,Q", [(50, 4693 * 14), (200, 3000), (256, 2400)])
def test_percentiles_many_short_columns(cuda_dev, env_override, N, Q):
    # the reference's own shape (50 realisations of a 4693 x 14 map, ECD.py:870-872) takes the warp-per-column kernel;
    # the shared-memory kernel forced onto the same data must agree with it and with numpy, NaN pixels included
    a = np.random.default_rng(N + Q + 5).lognormal(size=(N, Q))
    a[N // 2, 7] = np.nan
    ref = np.percentile(a, [25, 50, 75], axis=0)
    assert same(eb.ensemble_percentile(a, [25, 50, 75]), ref)
    env_override("ERTDIFF_PCTL_NO_WARP", 1)
    assert same(eb.ensemble_percentile(a, [25, 50, 75]), ref)


@pytest.mark.parametrize("N", [3000, 30000])
def test_kde_columns_much_narrower_than_the_grid_step(cuda_dev, N):
    # the grid spans the GLOBAL min..max (ECD.py:749-751): a column whose spread is far below the grid step has a
    # KDE that is exactly zero on almost every grid point -- in float32 for the scan (beyond 13.2 bandwidths) and in
    # float64 (beyond 38.6); the kernels skip those points and must still return scipy's argmax, index 0 when
    # every grid point is zero
    rng = np.random.default_rng(N)
    sig = np.array([1e-3, 1e-2, 0.1, 0.5, 3.0, 40.0, 1000.0, 2e-4])
    off = np.array([0.0, 17.3, -250.0, 1200.0, -3.0, 90.0, 0.0, -3999.0])
    a = rng.normal(size=(N, sig.size)) * sig + off
    grid = so.kde_grid(a, 5000)
    mode, idx = eb.ensemble_kde_mode(a, 5000, return_index=True)
    _, idx_sp, pdfs = so.kde_mode_scipy(a, grid)
    for j in np.nonzero(idx != idx_sp)[0]:
        p = pdfs[:, j]
        assert abs(p[idx[j]] - p[idx_sp[j]]) <= 1e-13 * p[idx_sp[j]], (j, idx[j], idx_sp[j], p[idx[j]], p[idx_sp[j]])
    assert (idx != idx_sp).sum() <= 1
    assert (pdfs.max(axis=0) == 0).any() and (pdfs.max(axis=0) > 0).any()      # both kinds of column are present


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("N,Q,col0,ncols", [(300, 29, 0, 29), (300, 29, 8, 4), (5000, 29, 25, 4), (2000, 64, 3, 17)])
def test_fused_summary_equals_the_separate_calls(cuda_dev, dtype, N, Q, col0, ncols):
    # ertdiff_ensemble_summary (one call, packed records, a column window of the array) against the separate
    # entry points on the sliced columns with the whole array's grid range: identical bits
    x = (torch.randn(N, Q, device=cuda_dev, generator=torch.Generator(cuda_dev).manual_seed(N + Q)) * 7 + 1).to(dtype)
    qs = [2.5, 50.0, 97.5]
    lohi = torch.empty(2, device=cuda_dev, dtype=torch.float64)
    block = eb.ensemble_summary_packed(x, qs, 777, col0=col0, ncols=ncols, lohi_out=lohi)
    cols = x[:, col0:col0 + ncols].contiguous()
    m = eb.ensemble_moments(cols)
    pct = eb.ensemble_percentile(cols, qs)
    assert torch.equal(lohi, eb.stats.global_minmax(x))
    mode, idx = eb.ensemble_kde_mode(cols, 777, grid_range=lohi, return_index=True)
    v = eb.stats.summary_views(block, 3)
    for k in ("mean", "std", "var"):
        assert torch.equal(v[k], m[k].double()), k
    assert torch.equal(v["pct"], pct) and torch.equal(v["mode"], mode) and torch.equal(v["mode_index"].long(), idx)


@pytest.mark.parametrize("N,Q", [(5000, 3), (2048, 4), (30000, 2), (60000, 1)])
def test_kde_few_columns_split_the_members_between_lanes(cuda_dev, N, Q):
    # a rank's share of the chain's output is a handful of columns of a long ensemble: the scan then lets several
    # lanes share a grid point and split the members (resident and tiled kernels); same argmax as scipy
    a = np.random.default_rng(N + Q).normal(size=(N, Q)) * np.linspace(1.0, 50.0, Q) + 3.0
    grid = so.kde_grid(a, 3000)
    mode, idx = eb.ensemble_kde_mode(a, 3000, return_index=True)
    _, idx_sp, pdfs = so.kde_mode_scipy(a, grid)
    for j in np.nonzero(idx != idx_sp)[0]:
        p = pdfs[:, j]
        assert abs(p[idx[j]] - p[idx_sp[j]]) <= 1e-13 * p[idx_sp[j]], (j, idx[j], idx_sp[j])
    assert (idx != idx_sp).sum() <= 1 and np.array_equal(mode, grid[idx])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N,Q", [(2, 3), (5, 29), (257, 64), (1000, 33), (1024, 2000), (4097, 40), (8192, 1100), (16000, 5)])
def test_percentiles_radix_selection_bit_exact(cuda_dev, env_override, dtype, N, Q):
    # exact selection by radix in shared memory (k_percentiles_select; the natural path of many medium-length columns
    # with <= 4 quantiles, forced here onto every shape): histogram + compaction rounds, buckets too large for the
    # compaction buffer (narrowed in place), ties, all-equal columns, infinities, signed zeros, NaN columns, keys that
    # differ only in their low bits
    env_override("ERTDIFF_PCTL_SELECT", 1)
    rng = np.random.default_rng(N * 13 + Q)
    a = rng.lognormal(sigma=2.0, size=(N, Q)).astype(dtype)
    a[:, 0] = rng.integers(-2, 3, size=N)                         # heavy ties: one bucket holds a fifth of the column
    if Q > 2:
        a[:, 1] = 7.25                                            # all members equal
        a[:, 2] = 1.0 + np.arange(N) * (1e-6 if dtype == np.float32 else 1e-13)     # only the low bits differ
    if Q > 4 and N > 6:
        a[[1, N // 2, N - 1], 3] = [np.inf, -np.inf, -0.0]
        a[N // 3, 4] = np.nan
    if Q > 5:
        a[:, 5] = -a[:, 5]                                        # negative values: the key transform's other branch
    for q in (50, [25, 50, 75], [0, 100], [2.5, 97.5, 33.3], np.float64(12.5), 99.9):
        ref = np.percentile(a, q, axis=0)
        got = eb.ensemble_percentile(a, q)
        assert same(got, ref), (q, np.nonzero(~((got == ref) | (np.isnan(got) & np.isnan(ref)))))


def _mixture_columns(rng, N, Q, dtype):
    """Columns of different shapes on one common grid: normal, bimodal (equal and unequal heights), log-normal, uniform,
    heavy ties, a narrow spike far from a broad bulk -- the shapes the coarse-to-fine scan has to get right."""
    cols = []
    for j in range(Q):
        k = j % 7
        if k == 0:
            c = rng.normal(0.0, 1.0 + j, N)
        elif k == 1:
            c = np.concatenate([rng.normal(-4.0, 0.5, N // 2), rng.normal(3.0, 0.5, N - N // 2)])
        elif k == 2:
            c = np.concatenate([rng.normal(-6.0, 1.5, (2 * N) // 3), rng.normal(5.0, 0.4, N - (2 * N) // 3)])
        elif k == 3:
            c = rng.lognormal(0.0, 0.7, N)
        elif k == 4:
            c = rng.uniform(-8.0, 8.0, N)
        elif k == 5:
            c = np.concatenate([rng.normal(0.0, 2.0, N - N // 8), np.full(N // 8, 1.25)])
        else:
            c = np.concatenate([rng.normal(0.0, 3.0, N - N // 5), rng.normal(25.0, 0.02, N // 5)])
        cols.append(c[:N])
    return np.stack(cols, axis=1).astype(dtype)


@pytest.mark.parametrize("N,Q,tile", [(1024, 29, 0), (3000, 7, 0), (2048, 4, 0), (2500, 700, 0), (6000, 14, 448), (30000, 3, 0), (70000, 2, 0)])
def test_kde_coarse_to_fine_scan_equals_full_scan(cuda_dev, env_override, N, Q, tile):
    # the scan evaluates every s-th grid point first and skips the intervals that provably hold no candidate
    # (ERTDIFF_KDE_COARSE=1 scans every point): same argmax index for every column shape, one or many CTAs per
    # column, members split between lanes, resident and tiled kernels -- and scipy's index
    a = _mixture_columns(np.random.default_rng(N + Q), N, Q, np.float32 if Q % 2 else np.float64)
    if tile:
        env_override("ERTDIFF_KDE_TILE", tile)
    env_override("ERTDIFF_KDE_COARSE", 1)
    m_full, i_full = eb.ensemble_kde_mode(a, 5000, return_index=True)
    for stride in (4, 32, 64):
        env_override("ERTDIFF_KDE_COARSE", stride)
        for rep in range(2):                   # (the shared-maximum cells are tagged per launch: a second call must not see the first's)
            m, i = eb.ensemble_kde_mode(a, 5000, return_index=True)
            assert np.array_equal(i, i_full) and np.array_equal(m, m_full), (stride, rep, np.nonzero(i != i_full)[0])
    if N <= 6000 and Q <= 29:
        grid = so.kde_grid(a, 5000)
        _, idx_sp, pdfs = so.kde_mode_scipy(a.astype(np.float64), grid)
        for j in np.nonzero(i_full != idx_sp)[0]:
            p = pdfs[:, j]
            assert abs(p[i_full[j]] - p[idx_sp[j]]) <= 1e-13 * p[idx_sp[j]], (j, i_full[j], idx_sp[j])
