"""Restatement of the thin adapter around ``src/gpr`` (TEST INFRASTRUCTURE, see oracle/__init__.py):
``src/core/gpr.rs`` (EstimatorGPR / SurrogateModelGPR), ``src/core/ynormalize.rs`` and
``expected_improvement`` from ``src/core/acquisition.rs:141-171``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import gpr
from .gpr import BoundedValue, BoundsError, ConstantKernel, Matern, Product

FUDGE_MIN = 0.05  # ynormalize.rs:5

LINEAR = "linear"
LOGARITHMIC = "logarithmic"


def _guess_min(known_optimum, y, minimum):
    # ynormalize.rs:291-304
    mn = y.min() - minimum
    if known_optimum is not None and known_optimum < mn:
        return known_optimum
    return mn


def _nd_mean(y: np.ndarray):
    """ndarray ``mean_axis`` = ``sum_axis / n``; for a 1-D array ``sum_axis`` takes the generic
    branch (``res = res + view_i``), i.e. sequential accumulation from zero."""
    dt = y.dtype.type
    acc = dt(0)
    for v in y:
        acc = dt(acc + v)
    return dt(acc / dt(len(y)))


def _guess_amplitude(y):
    # ynormalize.rs:307-320
    amplitude = _nd_mean(y)
    return amplitude if amplitude > 0 else y.dtype.type(1.0)


@dataclass
class YNormalize:
    """``src/core/ynormalize.rs:7-12, 158-288``."""

    amplitude: float
    expected: float
    projection: str
    A: type = np.float64

    @classmethod
    def new_project_into_normalized(cls, y, projection=LINEAR, known_optimum=None, A=np.float64):
        y = np.asarray(y, dtype=A)
        ko = None if known_optimum is None else A(known_optimum)
        if projection == LINEAR:
            expected = _guess_min(ko, y, A(0))
            y = y - expected
            amplitude = _guess_amplitude(y)
            y = y / amplitude + A(FUDGE_MIN)
        else:
            expected = _guess_min(ko, y, A(1.0))
            y = np.log(y - expected)
            amplitude = _guess_amplitude(y)
            y = y / amplitude
        return y, cls(amplitude, expected, projection, A)

    def project_into_normalized(self, y):
        A = self.A
        y = np.asarray(y, dtype=A)
        if self.projection == LINEAR:
            return (y - self.expected) / self.amplitude + A(FUDGE_MIN)
        return np.log(y - self.expected) / self.amplitude

    def project_location_from_normalized(self, y):
        A = self.A
        y = np.asarray(y, dtype=A)
        if self.projection == LINEAR:
            return (y - A(FUDGE_MIN)) * self.amplitude + self.expected
        return np.exp(y * self.amplitude) + self.expected

    def project_mean_from_normalized(self, mean, variance):
        A = self.A
        mean, variance = np.asarray(mean, dtype=A), np.asarray(variance, dtype=A)
        if self.projection == LINEAR:
            return (mean - A(FUDGE_MIN)) * self.amplitude + self.expected
        mean_amp = mean * self.amplitude
        var_amp = variance * self.amplitude * self.amplitude
        return np.exp(mean_amp + var_amp / A(2)) + self.expected

    def project_std_from_normalized(self, mean, variance):
        A = self.A
        mean, variance = np.asarray(mean, dtype=A), np.asarray(variance, dtype=A)
        if self.projection == LINEAR:
            return np.sqrt(variance) * self.amplitude
        mu = mean * self.amplitude
        sigma2 = variance * self.amplitude * self.amplitude
        # logwarp::project_variance_from (ynormalize.rs:116-122): exp(2 mu + s^2) * (exp(s^2) - 1)
        return np.sqrt(np.exp(mu * A(2) + sigma2) * (np.exp(sigma2) - A(1)))

    def project_cv_from_normalized(self, mean, variance):
        A = self.A
        mean, variance = np.asarray(mean, dtype=A), np.asarray(variance, dtype=A)
        if self.projection == LINEAR:
            return np.sqrt(variance) * self.amplitude / ((mean - A(FUDGE_MIN)) * self.amplitude + self.expected)
        return np.sqrt(np.exp(variance * self.amplitude ** 2) - A(1))


def _norm_cdf(z: float) -> float:
    # statrs 0.12 Normal::cdf = 0.5 * erfc((mean - x) / (std * sqrt(2)))
    return 0.5 * math.erfc(-z / math.sqrt(2.0))


def _norm_pdf(z: float) -> float:
    return math.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)


def expected_improvement(mean: float, std: float, fmin: float) -> float:
    """``src/core/acquisition.rs:141-171`` (f64)."""
    assert math.isfinite(mean) and math.isfinite(std) and math.isfinite(fmin)
    # `ulps_eq!(std, 0.0)` (approx 0.3): abs_diff_eq with epsilon = f64::EPSILON comes first, then <= 4 ULPs
    if std <= 0.0 or abs(std) <= 2.220446049250313e-16:
        return -(mean - fmin) if mean < fmin else 0.0
    z = -(mean - fmin) / std
    ei = -(mean - fmin) * _norm_cdf(z) + std * _norm_pdf(z)
    assert math.isfinite(ei), f"EI must be finite: {ei}"
    return ei


def estimate_amplitude(y: np.ndarray, bounds: Optional[Tuple[float, float]] = None) -> BoundedValue:
    """``src/core/gpr.rs:429-450``.  Quantile 0.1 with ``interpolate::Lower`` (ndarray-stats 0.3):
    the element at index floor((n - 1) * q) of the sorted data."""
    if bounds is None:
        y64 = np.asarray(y)
        hi = float(gpr.nd_sum(y64 * y64))
        srt = np.sort(y64.astype(np.float64))
        q = float(srt[int(math.floor((len(srt) - 1) * 0.1))])
        lo = q ** 2 * len(y64)
        assert lo >= 0.0
        lo = lo if lo > 2e-5 else 2e-5
        lo, hi = lo / 2.0, hi * 2.0
    else:
        lo, hi = bounds
    start = math.exp((math.log(lo) + math.log(hi)) / 2.0)
    return BoundedValue(start, lo, hi)


class SurrogateModelGPR:
    """``src/core/gpr.rs:53-213``."""

    def __init__(self, kernel, noise, x_train, y_train, alpha, k_inv, y_norm, lml, A):
        self.kernel, self.noise = kernel, noise
        self.x_train, self.y_train, self.alpha, self.k_inv = x_train, y_train, alpha, k_inv
        self.y_norm, self.lml, self.A = y_norm, lml, A

    def length_scales(self):
        return [b.value for b in self.kernel.k2.length_scale]

    def predict_mean_a(self, x):
        y = gpr.predict(self.kernel, self.alpha, x, self.x_train, self.k_inv, None, self.A)
        return self.y_norm.project_location_from_normalized(y)

    def predict_normalized(self, x):
        var = np.zeros(np.asarray(x).shape[0], dtype=self.A)
        mean = gpr.predict(self.kernel, self.alpha, x, self.x_train, self.k_inv, var, self.A)
        return mean, var

    def predict_confidence_bound(self, x, cb):
        mnorm, vnorm = self.predict_normalized(np.asarray(x, dtype=self.A)[None, :])
        return self.y_norm.project_location_from_normalized(mnorm + np.sqrt(vnorm) * self.A(cb))[0]

    def predict_mean_ei_a(self, x, fmin):
        y, y_var = self.predict_normalized(x)
        fmin_n = self.y_norm.project_into_normalized(np.array([fmin], dtype=self.A))[0]
        ei = np.array([self.A(expected_improvement(float(m), float(np.sqrt(v)), float(fmin_n)))
                       for m, v in zip(y, y_var)], dtype=self.A)
        return self.y_norm.project_location_from_normalized(y), ei


class EstimatorGPR:
    """``src/core/gpr.rs:215-400``."""

    def __init__(self, n_features: int):
        self.noise_bounds = (1e-5, 1e5)
        self.length_scale_bounds = [(1e-3, 1e3)] * n_features
        self.n_restarts_optimizer = 2
        self.matern_nu = 5.0 / 2.0
        self.amplitude_bounds = None
        self.y_projection = LINEAR
        self.known_optimum = None

    def default_kernel(self, amplitude: BoundedValue):
        # gpr.rs:402-427
        noise = BoundedValue(1.0, *self.noise_bounds)
        ls = [BoundedValue(math.exp((math.log(lo) + math.log(hi)) / 2.0), lo, hi) for lo, hi in self.length_scale_bounds]
        return Product(ConstantKernel(amplitude), Matern(self.matern_nu, ls)), noise

    def estimate(self, x, y, prior: Optional[SurrogateModelGPR], rng, minimize_by_gradient, A=np.float64):
        # gpr.rs:238-291
        x = np.asarray(x, dtype=A)
        assert len(y) == x.shape[0]
        y_train, y_norm = YNormalize.new_project_into_normalized(y, self.y_projection, self.known_optimum, A)
        amplitude = estimate_amplitude(y_train, self.amplitude_bounds)
        if prior is not None:
            kernel, noise = prior.kernel, prior.noise
        else:
            kernel, noise = self.default_kernel(amplitude)
        fk = gpr.fit_kernel(kernel, x, y_train, rng.fork_random_state(), self.n_restarts_optimizer, noise,
                            minimize_by_gradient, A)
        return SurrogateModelGPR(fk.kernel, fk.noise, x, y_train, fk.alpha, fk.k_inv, y_norm, fk.lml, A)

    def extend(self, x, y, prior: SurrogateModelGPR, A=np.float64):
        # gpr.rs:293-337
        x = np.asarray(x, dtype=A)
        y_train, y_norm = YNormalize.new_project_into_normalized(y, self.y_projection, self.known_optimum, A)
        fk = gpr.fitted_kernel_extend(prior.kernel, x, y_train, prior.noise, A)
        return SurrogateModelGPR(fk.kernel, fk.noise, x, y_train, fk.alpha, fk.k_inv, y_norm, fk.lml, A)
