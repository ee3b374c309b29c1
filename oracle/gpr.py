"""Restatement of the reference's ``src/gpr`` numerics (TEST INFRASTRUCTURE, see oracle/__init__.py).

All arrays are row-major ``(n_samples, n_features)`` like the reference's ``Array2<A>``
(``src/gpr/kernel.rs:10-12``).  ``A`` is ``numpy.float64`` (default) or ``numpy.float32``
(``--use-32``, ``src/gpr/scalar.rs:3-30``).  Kernel hyper-parameters stay f64 in both modes.
"""
from __future__ import annotations

import math
import sys
from dataclasses import dataclass, replace
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import lapack


# --------------------------------------------------------------------------- util
class BoundsError(ValueError):
    """``src/util/bounded_value.rs:72-77``."""

    def __init__(self, value, lo, hi):
        super().__init__(f"value {value} violated bounds [{lo}, {hi}]")
        self.value, self.min, self.max = value, lo, hi


@dataclass(frozen=True)
class BoundedValue:
    """``src/util/bounded_value.rs:3-56`` (inclusive bounds)."""

    value: float
    min: float
    max: float

    def __post_init__(self):
        if not (self.min <= self.value <= self.max):
            raise BoundsError(self.value, self.min, self.max)

    def with_value(self, value: float) -> "BoundedValue":
        return BoundedValue(value, self.min, self.max)

    def with_clamped_value(self, value: float) -> "BoundedValue":
        # bounded_value.rs:43-56: `value < min -> min; max < value -> max`
        if value < self.min:
            value = self.min
        elif self.max < value:
            value = self.max
        return BoundedValue(value, self.min, self.max)


def nd_sum(a: np.ndarray):
    """ndarray 0.13 ``ArrayBase::sum`` on a contiguous array: ``numeric_util::unrolled_fold``
    with eight running partial sums combined as ((p0+p4)+(p1+p5)+(p2+p6)+(p3+p7)) and a
    sequential tail (crate source absent from /root/reference; restated from the published
    crate).  Used by ``lml.rs:58`` and ``lml.rs:69``."""
    flat = np.ascontiguousarray(a).reshape(-1)
    dt = flat.dtype.type
    n8 = (flat.size // 8) * 8
    if n8:
        lanes = np.add.reduce(flat[:n8].reshape(-1, 8), axis=0, dtype=flat.dtype)  # sequential per lane
    else:
        lanes = np.zeros(8, dtype=flat.dtype)
    acc = dt(0)
    for i in range(4):
        acc = dt(acc + dt(lanes[i] + lanes[i + 4]))
    for x in flat[n8:]:
        acc = dt(acc + x)
    return acc


def cdist(xa: np.ndarray, xb: np.ndarray) -> np.ndarray:
    """``src/gpr/matern_kernel.rs:262-283``: Euclidean distances, accumulated left to right
    over the feature index in ``A``, full (na, nb) rectangle."""
    assert xa.shape[1] == xb.shape[1]
    acc = np.zeros((xa.shape[0], xb.shape[0]), dtype=xa.dtype)
    for k in range(xa.shape[1]):
        diff = xa[:, k, None] - xb[None, :, k]
        acc += diff * diff  # `(xa_i - xb_i).powi(2)`
    return np.sqrt(acc)


def outer(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``src/gpr/lml.rs:85-103``: out[i, j] = b[j] * a[i]."""
    out = np.broadcast_to(b[None, :], (a.shape[0], b.shape[0])).copy()
    out *= a[:, None]
    return out


# --------------------------------------------------------------------------- kernels
class ConstantKernel:
    """``src/gpr/constant_kernel.rs:9-67``."""

    def __init__(self, constant: BoundedValue):
        self.constant = constant

    def kernel(self, x1, x2, A=np.float64):
        return np.full((x1.shape[0], x2.shape[0]), A(self.constant.value), dtype=A)

    def theta_grad(self, x, A=np.float64):
        n = x.shape[0]
        return self.kernel(x, x, A), np.full((n, n, 1), A(self.constant.value), dtype=A)

    def diag(self, x, A=np.float64):
        return np.full(x.shape[0], A(self.constant.value), dtype=A)

    def n_params(self):
        return 1

    def theta(self) -> List[float]:
        return [math.log(self.constant.value)]

    def with_theta(self, theta):
        (t,) = theta
        return ConstantKernel(self.constant.with_value(math.exp(t)))

    def with_clamped_theta(self, theta):
        (t,) = theta
        return ConstantKernel(self.constant.with_clamped_value(math.exp(t)))

    def bounds(self):
        return [(math.log(self.constant.min), math.log(self.constant.max))]


def _nu_is(nu: float, target: float) -> bool:
    return abs(nu - target) <= sys.float_info.epsilon  # approx::abs_diff_eq! default epsilon


class Matern:
    """``src/gpr/matern_kernel.rs:12-187`` (nu in {0.5, 1.5, 2.5}; anything else is unimplemented)."""

    def __init__(self, nu: float, length_scale: Sequence[BoundedValue]):
        self.nu = nu
        self.length_scale = list(length_scale)

    def _ls(self, A):
        return np.array([b.value for b in self.length_scale], dtype=np.float64).astype(A)

    def kernel(self, x1, x2, A=np.float64):
        # matern_kernel.rs:37-81
        assert x1.shape[1] == self.n_params(), "number of x1 columns must match number of features"
        assert x2.shape[1] == self.n_params(), "number of x2 columns must match number of features"
        ls = self._ls(A)
        x1 = np.asarray(x1, dtype=A) / ls[None, :]  # divide first (:51-60)
        x2 = np.asarray(x2, dtype=A) / ls[None, :]
        dists = cdist(x1, x2)
        if _nu_is(self.nu, 0.5):
            return np.exp(-dists)
        if _nu_is(self.nu, 1.5):
            k = dists * A(math.sqrt(3.0))
            return (k + A(1)) * np.exp(-k)
        if _nu_is(self.nu, 2.5):
            k = dists * A(math.sqrt(5.0))
            return (A(1) + k + k * k / A(3)) * np.exp(-k)  # ((1 + k) + k^2/3) * exp(-k)  (:75)
        raise NotImplementedError("Matern kernel with arbitrary values for nu")

    def theta_grad(self, x, A=np.float64):
        # matern_kernel.rs:83-135.  Materialises the (n, n, d) tensor exactly like the reference.
        x = np.asarray(x, dtype=A)
        kernel = self.kernel(x, x, A)
        ls = self._ls(A)
        s2 = ls * ls  # `powi(2)`
        diff = x[:, None, :] - x[None, :, :]
        d = diff * diff  # raw x, subtract, square ...
        d = d / s2[None, None, :]  # ... then divide row-wise (:95-98)
        if _nu_is(self.nu, 0.5):
            with np.errstate(divide="ignore", invalid="ignore"):
                grad = kernel[:, :, None] * d / np.sqrt(_sum_axis2(d))[:, :, None]
            grad[~np.isfinite(grad)] = A(0)
        elif _nu_is(self.nu, 1.5):
            tmp = np.exp(-np.sqrt(_sum_axis2(d) * A(3)))
            grad = d * tmp[:, :, None] * A(3)
        elif _nu_is(self.nu, 2.5):
            tmp = np.sqrt(_sum_axis2(d) * A(5))[:, :, None]
            grad = np.exp(-tmp) * (tmp + A(1)) * d * A(5.0 / 3.0)  # multiplication order of :123-130
        else:
            raise NotImplementedError("Matern kernel gradient with arbitrary values for nu")
        return kernel, grad.astype(A, copy=False)

    def diag(self, x, A=np.float64):
        return np.ones(x.shape[0], dtype=A)

    def n_params(self):
        return len(self.length_scale)

    def theta(self):
        return [math.log(b.value) for b in self.length_scale]

    def with_theta(self, theta):
        assert len(theta) == self.n_params()
        return Matern(self.nu, [b.with_value(math.exp(t)) for t, b in zip(theta, self.length_scale)])

    def with_clamped_theta(self, theta):
        assert len(theta) == self.n_params()
        return Matern(self.nu, [b.with_clamped_value(math.exp(t)) for t, b in zip(theta, self.length_scale)])

    def bounds(self):
        return [(math.log(b.min), math.log(b.max)) for b in self.length_scale]


def _sum_axis2(d: np.ndarray) -> np.ndarray:
    """ndarray 0.13 ``sum_axis`` on a 3-D array: ``res = zeros; for i in 0..n { res = res + view_i }``
    i.e. left-to-right accumulation starting from 0 (used at ``matern_kernel.rs:104,115,120``)."""
    res = np.zeros(d.shape[:2], dtype=d.dtype)
    for k in range(d.shape[2]):
        res = res + d[:, :, k]
    return res


class Product:
    """``src/gpr/product_kernel.rs:8-109``."""

    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2

    def kernel(self, xa, xb, A=np.float64):
        return self.k1.kernel(xa, xb, A) * self.k2.kernel(xa, xb, A)

    def theta_grad(self, x, A=np.float64):
        kernel1, gradient1 = self.k1.theta_grad(x, A)
        kernel2, gradient2 = self.k2.theta_grad(x, A)
        kernel = kernel1 * kernel2
        g1k2 = gradient1 * kernel2[:, :, None]
        g2k1 = gradient2 * kernel1[:, :, None]
        return kernel, np.concatenate([g1k2, g2k1], axis=2)

    def diag(self, x, A=np.float64):
        return self.k1.diag(x, A) * self.k2.diag(x, A)

    def n_params(self):
        return self.k1.n_params() + self.k2.n_params()

    def theta(self):
        return self.k1.theta() + self.k2.theta()

    def with_theta(self, theta):
        assert len(theta) == self.n_params()
        n1 = self.k1.n_params()
        return Product(self.k1.with_theta(theta[:n1]), self.k2.with_theta(theta[n1:]))

    def with_clamped_theta(self, theta):
        assert len(theta) == self.n_params()
        n1 = self.k1.n_params()
        return Product(self.k1.with_clamped_theta(theta[:n1]), self.k2.with_clamped_theta(theta[n1:]))

    def bounds(self):
        return self.k1.bounds() + self.k2.bounds()


# --------------------------------------------------------------------------- LAPACK
def _lapack(name: str, A):
    return getattr(lapack, ("d" if A == np.float64 else "s") + name)


class CholeskyFactorized:
    """ndarray-linalg 0.12 ``CholeskyFactorized`` (UPLO::Lower): ``factor`` holds L in its lower triangle."""

    def __init__(self, factor: np.ndarray, A):
        self.factor, self.A = factor, A

    def solvec(self, b):
        x, info = _lapack("potrs", self.A)(self.factor, np.asarray(b, dtype=self.A), lower=1)
        assert info == 0
        return x

    def invc(self):
        inv, info = _lapack("potri", self.A)(self.factor, lower=1)
        assert info == 0
        inv = np.tril(inv)
        return inv + np.tril(inv, -1).T  # hermitian fill

    def ln_diag_sum(self):
        return nd_sum(np.log(np.ascontiguousarray(np.diag(self.factor))))


def factorizec(k: np.ndarray, A) -> Optional[CholeskyFactorized]:
    c, info = _lapack("potrf", A)(k, lower=1, clean=0)
    if info != 0:
        return None
    return CholeskyFactorized(c, A)


# --------------------------------------------------------------------------- lml
@dataclass
class LmlWithGradient:
    """``src/gpr/lml.rs:8-13``."""

    lml: float
    lml_gradient: List[float]
    alpha: np.ndarray
    factorization: CholeskyFactorized


def lml_with_gradient(kernel, noise, x_train, y_train, A=np.float64, want_gradient=True) -> Optional[LmlWithGradient]:
    """``src/gpr/lml.rs:29-79``.  ``noise`` is already an ``A`` value."""
    x_train = np.asarray(x_train, dtype=A)
    y_train = np.asarray(y_train, dtype=A)
    n = x_train.shape[0]
    kernel_matrix, kernel_gradient = kernel.theta_grad(x_train, A)
    noise = A(noise)
    kernel_matrix = kernel_matrix.copy()
    kernel_matrix[np.diag_indices(n)] += noise  # lml.rs:44
    fact = factorizec(kernel_matrix, A)
    if fact is None:
        return None
    alpha = fact.solvec(y_train)
    lml = -0.5 * float(np.dot(y_train, alpha)) - float(fact.ln_diag_sum()) - n / 2.0 * math.log(2.0 * math.pi)
    grad: List[float] = []
    if want_gradient:
        tmp = outer(alpha, alpha) - fact.invc()
        # noise slice is `eye * noise` (lml.rs:41): the product with tmp keeps the diagonal only
        noise_gradient = np.zeros((n, n), dtype=A)
        noise_gradient[np.diag_indices(n)] = noise
        grad.append(0.5 * float(nd_sum(tmp * noise_gradient)))
        for k in range(kernel_gradient.shape[2]):
            grad.append(0.5 * float(nd_sum(tmp * kernel_gradient[:, :, k])))
    return LmlWithGradient(lml, grad, alpha, fact)


# --------------------------------------------------------------------------- predict
def clamp_negative_variance(variances: np.ndarray, warning_level) -> Optional[list]:
    """``src/gpr/predict.rs:104-127`` (in place)."""
    below = [v for v in variances if v < warning_level]
    variances[variances < 0] = 0
    return below or None


def predict(kernel, alpha, x, x_train, k_inv, want_variance: Optional[np.ndarray] = None, A=np.float64):
    """``src/gpr/predict.rs:7-52``.  ``want_variance`` (length m) is filled in place when given."""
    x = np.asarray(x, dtype=A)
    k_trans = kernel.kernel(x, np.asarray(x_train, dtype=A), A)
    y = k_trans.dot(alpha)
    if want_variance is not None:
        min_noise = A(1e-5)
        prod = k_trans.dot(k_inv)
        rows = np.einsum("ij,ij->i", prod, k_trans).astype(A)
        y_var = kernel.diag(x, A) + min_noise - rows
        below = clamp_negative_variance(y_var, -np.sqrt(min_noise))
        if below is not None:
            print("Variances below 0 were predicted and will be corrected: "
                  + ", ".join(f"{v:.2e}" for v in below), file=sys.stderr)
        want_variance[...] = y_var
    return y


# --------------------------------------------------------------------------- fit
@dataclass
class FittedKernel:
    """``src/gpr/fit.rs:6-12``."""

    kernel: Product
    noise: BoundedValue
    alpha: np.ndarray
    k_inv: np.ndarray
    lml: float
    n_evals: int = 0
    trace: Optional[list] = None


def fitted_kernel_extend(kernel, x_train, y_train, noise: BoundedValue, A=np.float64) -> FittedKernel:
    """``src/gpr/fit.rs:33-68``: one evaluation at the prior's theta, no optimisation."""
    res = lml_with_gradient(kernel, A(noise.value), x_train, y_train, A)
    if res is None:
        raise RuntimeError("Kernel matrix must be invertible.")
    return FittedKernel(kernel, noise, res.alpha, res.factorization.invc(), res.lml, 1)


def fit_kernel(kernel, x_train, y_train, rng, n_restarts_optimizer: int, noise: BoundedValue,
               minimize_by_gradient, A=np.float64, keep_trace=False) -> FittedKernel:
    """``src/gpr/fit.rs:71-176`` with ``src/util/gradmin.rs:7-33`` inlined.

    ``minimize_by_gradient(objective, x0, bounds) -> (x, f)`` stands in for NLopt's L-BFGS
    (``gradmin.rs:35-60``; source absent): ``objective(theta) -> (f, grad)``.
    """
    x_train = np.asarray(x_train, dtype=A)
    y_train = np.asarray(y_train, dtype=A)
    assert y_train.shape[0] == x_train.shape[0]
    capture = {}
    trace = [] if keep_trace else None
    n_evals = [0]

    def obj_func(theta):
        noise_theta, kernel_theta = theta[0], list(theta[1:])
        k = kernel.with_clamped_theta(kernel_theta)
        nz = A(math.exp(noise_theta))
        n_evals[0] += 1
        res = lml_with_gradient(k, nz, x_train, y_train, A)
        if res is None:
            if trace is not None:
                trace.append((np.array(theta, dtype=np.float64), math.inf, None))
            return math.inf, np.zeros(len(theta))
        if not capture or res.lml > capture["lml"]:  # strict `>` (fit.rs:116-117)
            capture.update(lml=res.lml, theta=kernel_theta, noise=float(nz), fact=res.factorization, alpha=res.alpha)
        if trace is not None:
            trace.append((np.array(theta, dtype=np.float64), res.lml, np.array(res.lml_gradient)))
        return -res.lml, -np.array(res.lml_gradient, dtype=np.float64)

    bounds = [(math.log(noise.min), math.log(noise.max))] + kernel.bounds()
    theta0 = np.array([math.log(noise.value)] + kernel.theta(), dtype=np.float64)

    # gradmin.rs:19-30
    minimize_by_gradient(obj_func, theta0, bounds)
    for _ in range(n_restarts_optimizer):
        start = np.array([rng.uniform_inclusive(lo, hi) for lo, hi in bounds], dtype=np.float64)
        minimize_by_gradient(obj_func, start, bounds)

    if not capture:
        raise RuntimeError("called `Option::unwrap()` on a `None` value")  # fit.rs:161
    k = kernel.with_clamped_theta(capture["theta"])
    nz = noise.with_clamped_value(capture["noise"])
    return FittedKernel(k, nz, capture["alpha"], capture["fact"].invc(), capture["lml"], n_evals[0], trace)
