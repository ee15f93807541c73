"""UQ calibration metrics (ECD.py:1089-1137, 1191-1214): the oracle restatement against a fixture
produced by the reference's own helper functions, and the device path against the oracle."""
import warnings

import numpy as np
import pytest

from oracle import stats_oracle as so
from oracle.reference_loader import load_reference, reference_available

KEYS = ("avg_proportion", "accuracy", "precision", "goodness", "param_avg_proportion", "param_accuracy",
        "param_precision", "param_goodness")


def test_oracle_matches_reference_fixture(golden):
    g = golden("uq_calibration.npz")
    o = so.uq_calibration(g["generated"], g["true"])
    assert np.array_equal(o["prob_array"], g["prob_array"])
    for k in KEYS:
        assert np.array_equal(np.asarray(o[k]), g[k]), k


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_helpers_equal_reference_live():
    ref = load_reference()
    rng = np.random.default_rng(0)
    prob = np.linspace(0.01, 0.99, 30)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)          # the reference calls np.trapz
        for _ in range(20):
            avg = np.clip(prob + rng.normal(scale=0.2, size=30), 0, 1)
            a_ref = ref.avg_prop_indicator_function(avg, prob)
            a = so.avg_prop_indicator(avg, prob)
            assert np.array_equal(a, a_ref)
            acc = ref.accuracy_score(a_ref, prob)
            assert so.accuracy_score(a, prob) == acc
            assert so.precision_score(acc, avg, prob, a) == ref.preccision_score(acc, avg, prob, a_ref)
            assert so.goodness_score(a, avg, prob) == ref.goodness_score(a_ref, avg, prob)
        zero = np.zeros(30)                                          # accuracy == 0 branch
        assert so.precision_score(0, zero, prob, zero.astype(int)) == ref.preccision_score(0, zero, prob, zero.astype(int))


@pytest.mark.gpu
@pytest.mark.parametrize("N,M,dtype", [(50, 8, np.float32), (50, 32, np.float32), (7, 3, np.float64), (256, 5, np.float32)])
def test_device_uq_calibration_is_bit_exact(cuda_dev, golden, N, M, dtype):
    import ertdiff_b200 as eb
    if (N, M) == (50, 8):
        g = golden("uq_calibration.npz")
        gen, truth = g["generated"], g["true"]
    else:
        rng = np.random.default_rng(N * 100 + M)
        truth = rng.normal(size=(M, 29)).astype(dtype)
        gen = (truth[None] + rng.normal(scale=1.3, size=(N, M, 29))).astype(dtype)
        gen[:, 0, 3] = truth[0, 3]                                   # ties: low == true is NOT inside, true == upp is
    o = so.uq_calibration(gen, truth)
    d = eb.uq_calibration(gen, truth, device=cuda_dev)
    assert np.array_equal(d["prob_array"], o["prob_array"])
    for k in KEYS:
        assert np.array_equal(np.asarray(d[k]), np.asarray(o[k])), k
