"""logits -> physical parameters -> bounds, against the reference's own host
steps (ECD.py:42-53, 402-406, 183-218) restated with torch / sklearn / numpy."""
import numpy as np
import pytest
import torch

import ertdiff_b200 as eb

pytestmark = pytest.mark.gpu


def test_untransform_and_bounds(cuda_dev):
    from sklearn.preprocessing import MinMaxScaler
    rng = np.random.default_rng(0)
    P, B = 29, 500
    raw = rng.uniform(-3, 40, size=(200, P)) * rng.uniform(0.1, 100, size=(1, P))
    scaler = MinMaxScaler().fit(raw)
    u = torch.from_numpy(rng.normal(scale=1.6, size=(B, P)).astype(np.float32))
    # the reference's sequence: torch sigmoid (fp32) -> numpy -> sklearn inverse -> python bounds loop
    s = (0.0 + (1.0 - 0.0) * torch.sigmoid(u)).numpy()
    phys_ref = scaler.inverse_transform(s.copy())
    limits = np.stack([raw.min(0) + 0.02 * np.ptp(raw, axis=0), raw.max(0) - 0.02 * np.ptp(raw, axis=0)], axis=1)
    valid_ref = np.array([all(lo <= v <= hi for v, (lo, hi) in zip(row, limits)) for row in phys_ref])
    first_ref = np.array([next((i for i, (v, (lo, hi)) in enumerate(zip(row, limits)) if v < lo or v > hi), -1)
                          for row in phys_ref])
    phys, valid, first_bad = eb.untransform_and_check(u.to(cuda_dev), 0.0, 1.0, scaler.min_, scaler.scale_, limits)
    assert phys.dtype == torch.float32 and phys_ref.dtype == np.float32
    # sigmoid differs from torch's CPU kernel by <= 2 ulp of fp32 (|ds| <= 2.4e-7); the scaler
    # then maps ds to ds/scale_ and rounds once more, so the bound is per parameter, in units of
    # its data range -- not relative to the value (values near zero are differences of large terms)
    tol = 3e-7 / scaler.scale_ + 2e-7 * np.abs(phys_ref).max(axis=0)
    dmax = np.abs(phys.cpu().numpy().astype(np.float64) - phys_ref) / tol[None, :]
    assert dmax.max() <= 1.0, dmax.max()
    # rows with a parameter sitting within the sigmoid tolerance of a limit may flip
    near = (np.minimum(np.abs(phys_ref - limits[None, :, 0]), np.abs(phys_ref - limits[None, :, 1]))
            <= tol[None, :]).any(axis=1)
    assert near.mean() < 0.02, near.mean()
    v, fb = valid.cpu().numpy(), first_bad.cpu().numpy()
    assert (v[~near] == valid_ref[~near]).all(), np.nonzero(v != valid_ref)
    assert (fb[~near] == first_ref[~near]).all(), (np.nonzero(fb != first_ref), fb[fb != first_ref], first_ref[fb != first_ref])
    assert valid_ref.sum() > 0 and (~valid_ref).sum() > 0
    # sigmoid only
    s_gpu = eb.inverse_transform(u.to(cuda_dev), 0.0, 1.0)
    np.testing.assert_allclose(s_gpu.cpu().numpy(), s, rtol=3e-7, atol=1e-38)
    # bounds filter keeps whole rows, None when nothing survives (ECD.py:211-218)
    kept = eb.check_param_bounds(phys_ref, limits)
    assert kept.shape[0] == valid_ref.sum() and np.array_equal(kept, phys_ref[valid_ref])
    assert eb.check_param_bounds(phys_ref, np.stack([limits[:, 1] + 1, limits[:, 1] + 2], 1)) is None


def test_f1_epilogue_against_the_reference_fixture(cuda_dev, golden):
    # tests/golden/transforms_f1.npz: the reference's own inverse_transform (ECD.py:42-53, torch branch) ->
    # sklearn MinMaxScaler.inverse_transform -> the reference's own check_param_bounds (ECD.py:183-218) against
    # ParameterLimits().plims (GEU.py:8-59), run unmodified by oracle/make_golden.py round2; first_bad is parsed
    # from what check_param_bounds printed
    g = golden("transforms_f1.npz")
    u = torch.from_numpy(g["u"]).to(cuda_dev)
    phys, valid, first_bad = eb.untransform_and_check(u, 0.0, 1.0, g["scaler_min"], g["scaler_scale"], g["limits"])
    phys_ref, lim = g["phys"], g["limits"]
    assert phys.dtype == torch.float32 and phys_ref.dtype == np.float32
    # the device sigmoid differs from torch's CPU kernel by <= 2 ulp of fp32; the scaler maps that to
    # 2.4e-7 / scale_ per parameter and rounds once more (values near zero are differences of large terms)
    tol = 3e-7 / g["scaler_scale"] + 2e-7 * np.abs(phys_ref).max(axis=0)
    got = phys.cpu().numpy().astype(np.float64)
    assert (np.abs(got - phys_ref) / tol[None, :]).max() <= 1.0
    np.testing.assert_allclose(eb.inverse_transform(u, 0.0, 1.0).cpu().numpy(), g["sigmoid"], rtol=3e-7, atol=1e-38)
    # rows with a parameter within that tolerance of a limit may legitimately flip; there must be few of them
    near = (np.minimum(np.abs(phys_ref - lim[None, :, 0]), np.abs(phys_ref - lim[None, :, 1])) <= tol[None, :]).any(axis=1)
    assert near.mean() < 0.02
    v, fb = valid.cpu().numpy(), first_bad.cpu().numpy()
    assert np.array_equal(v[~near], g["valid"][~near]) and np.array_equal(fb[~near], g["first_bad"][~near])
    assert 0 < g["valid"].sum() < g["valid"].size
    # the bounds filter alone, fed the reference's own values: exact, float32 and float64, row order kept
    for key, fbk in (("phys", "first_bad"), ("phys64", "first_bad64")):
        vals = g[key]
        kept = eb.check_param_bounds(vals, lim, verbose=False)
        assert kept.dtype == vals.dtype and np.array_equal(kept, vals[g[fbk] < 0])
        kept_t = eb.check_param_bounds(torch.from_numpy(vals).to(cuda_dev), lim, verbose=False)
        assert kept_t.is_cuda and np.array_equal(kept_t.cpu().numpy(), vals[g[fbk] < 0])


def test_check_param_bounds_semantics(cuda_dev, capsys):
    lim = np.array([[0.0, 1.0], [-1.0, 1.0], [1e-12, 1e-8]])
    vals = np.array([[0.5, 0.0, 1e-9],
                     [np.nan, 0.0, 1e-9],            # NaN is neither < min nor > max: the reference keeps the row
                     [0.5, 1.0000000000000002, 1e-9],    # float64 values one ulp outside a limit are out
                     [1.0, -1.0, 1e-8],              # the limits themselves are inside
                     [0.5, 2.0, 1.0]])               # the FIRST offending parameter is reported
    kept = eb.check_param_bounds(vals, lim)
    assert np.array_equal(kept, vals[[0, 1, 3]], equal_nan=True)
    out = capsys.readouterr().out
    assert "Sample 2 Parameter 1: 1.0000 (out of bounds [-1.0000, 1.0000])" in out
    assert "Sample 4 Parameter 1: 2.0000" in out and "Sample 4 Parameter 2" not in out
    assert eb.check_param_bounds(vals[[2, 4]], lim, verbose=False) is None
