"""Host-side adapter (C++ behind the C ABI) against the oracle's restatement of src/core/ynormalize.rs,
estimate_amplitude (gpr.rs:429-450), expected_improvement (acquisition.rs:141-171) and the reference's own
ynormalize tests (ynormalize.rs:322-524)."""
import math

import numpy as np
import pytest

import hbetune_rs_b200 as h
from oracle import adapter as oad


@pytest.mark.parametrize("A", [np.float64, np.float32])
@pytest.mark.parametrize("proj", ["linear", "logarithmic"])
@pytest.mark.parametrize("ko", [None, -3.0, 50.0])
def test_ynormalize_matches_oracle(A, proj, ko):
    rng = np.random.default_rng(3)
    y = (rng.random(37) * 10 + 2).astype(A)
    code = h.LINEAR if proj == "linear" else h.LOGARITHMIC
    yn, cfg = h.YNormalize.new_project_into_normalized(y, code, ko, A)
    yo, ocfg = oad.YNormalize.new_project_into_normalized(y, proj, ko, A)
    tol = 1e-14 if A == np.float64 else 1e-6
    np.testing.assert_allclose(yn, yo, rtol=tol)
    assert abs(cfg.amplitude - float(ocfg.amplitude)) <= tol * abs(float(ocfg.amplitude))
    assert abs(cfg.expected - float(ocfg.expected)) <= tol * max(1.0, abs(float(ocfg.expected)))
    mean = (rng.random(9) + 0.1).astype(A)
    var = (rng.random(9) * 0.2 + 0.01).astype(A)
    ytest = (rng.random(9) * 10 + 60).astype(A)
    np.testing.assert_allclose(cfg.project_into_normalized(ytest), ocfg.project_into_normalized(ytest), rtol=tol * 10)
    np.testing.assert_allclose(cfg.project_location_from_normalized(mean), ocfg.project_location_from_normalized(mean), rtol=tol * 10)
    np.testing.assert_allclose(cfg.project_mean_from_normalized(mean, var), ocfg.project_mean_from_normalized(mean, var), rtol=tol * 10)
    np.testing.assert_allclose(cfg.project_std_from_normalized(mean, var), ocfg.project_std_from_normalized(mean, var), rtol=tol * 10)
    np.testing.assert_allclose(cfg.project_cv_from_normalized(mean, var), ocfg.project_cv_from_normalized(mean, var), rtol=tol * 10)
    # inverse property (ynormalize.rs:358-380)
    back = cfg.project_location_from_normalized(cfg.project_into_normalized(ytest))
    np.testing.assert_allclose(back, ytest, rtol=1e-12 if A == np.float64 else 2e-5)


def test_linear_normalisation_has_unit_mean_plus_fudge():
    y = np.array([3.0, 5.0, 10.0])
    yn, cfg = h.YNormalize.new_project_into_normalized(y)
    assert cfg.expected == 3.0 and abs(yn.mean() - 1.05) < 1e-15 and yn.min() == 0.05
    yn, cfg = h.YNormalize.new_project_into_normalized(np.array([2.0, 2.0]))  # no deviation: amplitude 1
    assert cfg.amplitude == 1.0 and (yn == 0.05).all()


@pytest.mark.parametrize("A", [np.float64, np.float32])
def test_estimate_amplitude_matches_oracle(A):
    rng = np.random.default_rng(0)
    for n in (1, 2, 10, 11, 99, 1000):
        y = (rng.random(n) * 3 + 0.05).astype(A)
        a, b = h.estimate_amplitude(y), oad.estimate_amplitude(y)
        assert a.min == pytest.approx(b.min, rel=1e-6 if A == np.float32 else 1e-14)
        assert a.max == pytest.approx(b.max, rel=1e-6 if A == np.float32 else 1e-14)
        assert a.value == pytest.approx(b.value, rel=1e-6 if A == np.float32 else 1e-14)
    assert h.estimate_amplitude(np.array([1e-3, 1.0, 1.0, 1.0, 1.0])).min == 1e-5  # floor of 2e-5 / 2
    with pytest.raises(h.HbegpError):  # all-zero y: start = 0 < lo, the reference's BoundedValue::new(..).unwrap() panics
        h.estimate_amplitude(np.zeros(5))
    fixed = h.estimate_amplitude(np.ones(5), (0.5, 8.0))
    assert (fixed.min, fixed.max, fixed.value) == (0.5, 8.0, pytest.approx(2.0))


def test_expected_improvement_matches_oracle():
    rng = np.random.default_rng(1)
    for _ in range(200):
        mean, std, fmin = rng.normal(), abs(rng.normal()) * 2, rng.normal()
        assert h.expected_improvement(mean, std, fmin) == pytest.approx(oad.expected_improvement(mean, std, fmin), rel=1e-13, abs=1e-300)
    assert h.expected_improvement(1.0, 0.0, 2.0) == 1.0  # std == 0: guaranteed improvement
    assert h.expected_improvement(3.0, 0.0, 2.0) == 0.0
    assert h.expected_improvement(3.0, -1.0, 2.0) == 0.0
    assert h.expected_improvement(0.0, 1.0, 0.0) == pytest.approx(1 / math.sqrt(2 * math.pi))


def test_normal_inverse_cdf():
    from scipy.stats import norm
    for p in (1e-12, 1e-5, 0.01, 0.25, 0.5, 0.75, 0.99, 1 - 1e-9):
        assert h.lib.hbegp_normal_inverse_cdf(p, 0.0, 1.0) == pytest.approx(norm.ppf(p), rel=1e-13, abs=1e-15)
    assert h.lib.hbegp_normal_inverse_cdf(0.75, 2.0, 3.0) == pytest.approx(2.0 + 3.0 * 0.6744897501960817)
