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
