"""Per-member misfit metrics (ECD.py:764-785 WSSE per survey / total / ranking, ECD.py:927-930 MSE):
the oracle against a fixture made by the reference's own ``WSSE_metric`` and sklearn's
``mean_squared_error``; the device path against both, bit for bit, in float64 and float32."""
import contextlib
import io

import numpy as np
import pytest

from oracle import stats_oracle as so
from oracle.reference_loader import load_reference, reference_available


@pytest.mark.parametrize("tag,dt", [("f64", np.float64), ("f32", np.float32)])
def test_oracle_matches_reference_fixture(golden, tag, dt):
    g = golden("misfit.npz")
    o = so.misfit_metrics(g["sim_data"].astype(dt), g["observed"].astype(dt))
    assert o["wsse"].dtype == dt and o["mse"].dtype == dt
    assert np.array_equal(o["wsse"], g[f"wsse_{tag}"])
    assert np.array_equal(o["wsse_total"], g[f"wsse_total_{tag}"])
    assert np.array_equal(o["mse"].astype(np.float64), g[f"mse_{tag}"])     # sklearn returns python floats


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_pairwise_sum_restatement_is_numpys_reduction(dt):
    rng = np.random.default_rng(1)
    for n in (1, 5, 7, 8, 9, 14, 100, 128, 129, 255, 256, 1000, 4693):
        a = (rng.standard_normal(n) ** 2).astype(dt)
        assert so.pairwise_sum(a) == np.add.reduce(a), n
        assert dt(so.pairwise_sum(a) / dt(n)) == np.average(a), n


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_equals_reference_live():
    ref = load_reference()
    rng = np.random.default_rng(2)
    obs = rng.normal(size=(97, 3))
    sims = obs[None] + rng.normal(scale=0.3, size=(4, 97, 3))
    o = so.misfit_metrics(sims, obs, A=0.2, B=0.05)
    with contextlib.redirect_stdout(io.StringIO()):
        w = np.array([[ref.WSSE_metric(0.2, 0.05, sims[i][:, es], obs[:, es])[0] for es in range(3)] for i in range(4)])
    assert np.array_equal(o["wsse"], w)
    from sklearn.metrics import mean_squared_error
    assert [mean_squared_error(obs.flatten(), sims[i].flatten()) for i in range(4)] == list(o["mse"])


# ---- device path ------------------------------------------------------------------------------------
def _check(res, o):
    for k in ("wsse", "wsse_total", "mse", "order"):
        assert res[k].shape == o[k].shape, k
        assert np.array_equal(res[k], o[k]), k
    assert res["wsse"].dtype == o["wsse"].dtype and res["mse"].dtype == o["mse"].dtype


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dt", [("f64", np.float64), ("f32", np.float32)])
def test_device_matches_reference_fixture(cuda_dev, golden, tag, dt):
    import ertdiff_b200 as eb
    g = golden("misfit.npz")
    res = eb.misfit_metrics(g["sim_data"].astype(dt), g["observed"].astype(dt), device=cuda_dev)
    assert np.array_equal(res["wsse"], g[f"wsse_{tag}"])
    assert np.array_equal(res["wsse_total"], g[f"wsse_total_{tag}"])
    assert np.array_equal(res["mse"].astype(np.float64), g[f"mse_{tag}"])


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("N,L,C", [(1, 1, 1), (3, 5, 2), (2, 8, 9), (4, 129, 14), (2, 1000, 7), (3, 4693, 14), (2, 300, 128)])
def test_device_is_bit_identical_to_numpy(cuda_dev, dt, N, L, C):
    # sizes walk the summation tree's cases: n < 8, one leaf, a tail, several levels, the reference's grid
    import ertdiff_b200 as eb
    rng = np.random.default_rng(N * 1000 + L + C)
    obs = rng.normal(2.0, 1.5, size=(L, C)).astype(dt)
    sims = (obs[None] + rng.normal(scale=0.5, size=(N, L, C))).astype(dt)
    _check(eb.misfit_metrics(sims, obs, device=cuda_dev), so.misfit_metrics(sims, obs))
    _check(eb.misfit_metrics(sims, obs, A=0.3, B=0.5, device=cuda_dev), so.misfit_metrics(sims, obs, A=0.3, B=0.5))


@pytest.mark.gpu
def test_device_tensors_and_single_map(cuda_dev):
    # CUDA tensors in -> CUDA tensors out; a single (L, C) map (ensemble mean / mode vs the observation,
    # ECD.py:939-940); an exact copy of the observation has zero misfit and ranks first
    import torch
    import ertdiff_b200 as eb
    rng = np.random.default_rng(5)
    obs = rng.normal(size=(4693, 14)).astype(np.float32)
    sims = obs[None] + rng.normal(scale=0.2, size=(50, 4693, 14)).astype(np.float32)
    sims[17] = obs
    res = eb.misfit_metrics(torch.from_numpy(sims).to(cuda_dev), torch.from_numpy(obs).to(cuda_dev))
    assert all(v.is_cuda for v in res.values())
    o = so.misfit_metrics(sims, obs)
    _check({k: v.cpu().numpy() for k, v in res.items()}, o)
    assert res["mse"][17].item() == 0.0 and res["wsse_total"][17].item() == 0.0 and res["order"][0].item() == 17
    one = eb.misfit_metrics(sims.mean(axis=0), obs, device=cuda_dev)
    assert one["mse"].shape == (1,) and one["mse"][0] == np.average((obs.flatten() - sims.mean(axis=0).flatten()) ** 2)
    with pytest.raises(ValueError):
        eb.misfit_metrics(sims, obs[:, :3], device=cuda_dev)


# ---- Wasserstein distance (ECD.py:860, 898-899) -------------------------------------------------------
def test_wasserstein_oracle_is_scipys():
    from scipy.stats import wasserstein_distance
    rng = np.random.default_rng(3)
    for n, m in ((1, 1), (5, 3), (100, 100), (1000, 777), (65702, 65702)):
        u, v = rng.normal(size=n), rng.normal(0.3, 1.2, size=m)
        assert so.wasserstein_distance(u, v) == wasserstein_distance(u, v)
    u = rng.integers(0, 5, size=200).astype(np.float32)               # heavy ties, float32 input
    v = rng.integers(0, 5, size=150).astype(np.float32)
    assert so.wasserstein_distance(u, v) == wasserstein_distance(u, v)


@pytest.mark.gpu
def test_wasserstein_device_matches_scipy(cuda_dev, golden):
    import ertdiff_b200 as eb
    from scipy.stats import wasserstein_distance
    rng = np.random.default_rng(4)
    for n, m, dt in ((1, 1, np.float64), (5, 3, np.float64), (100, 129, np.float32), (1000, 777, np.float64),
                     (4096, 4096, np.float32)):
        u, v = rng.normal(size=n).astype(dt), rng.normal(0.3, 1.2, size=m).astype(dt)
        d = eb.wasserstein_distance(u, v, device=cuda_dev)
        assert isinstance(d, float) and d == pytest.approx(wasserstein_distance(u, v), rel=1e-12, abs=1e-300)
    u = rng.integers(0, 5, size=200).astype(np.float64)               # runs of equal values in and across samples
    v = rng.integers(0, 5, size=150).astype(np.float64)
    assert eb.wasserstein_distance(u, v, device=cuda_dev) == pytest.approx(wasserstein_distance(u, v), rel=1e-12)
    assert eb.wasserstein_distance(u, u, device=cuda_dev) == 0.0
    # the reference's use: every simulated map (and the ensemble mean) against the observed map, full grid
    obs = rng.normal(size=(4693, 14)).astype(np.float32)
    sims = obs[None] + rng.normal(scale=0.2, size=(6, 4693, 14)).astype(np.float32)
    d = eb.wasserstein_distance(sims, obs, device=cuda_dev)
    ref = np.array([wasserstein_distance(sims[i].flatten(), obs.flatten()) for i in range(6)])
    assert d.shape == (6,) and np.allclose(d, ref, rtol=1e-12, atol=0)
    dm = eb.wasserstein_distance(sims.mean(axis=0), obs, device=cuda_dev)
    assert dm == pytest.approx(wasserstein_distance(sims.mean(axis=0).flatten(), obs.flatten()), rel=1e-12)
    with pytest.raises(ValueError):
        eb.wasserstein_distance(np.zeros(0), obs, device=cuda_dev)
