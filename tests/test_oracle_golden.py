"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md section 8c-3): sklearn-derived kernel/gradient goldens, cdist, outer,
clamp_negative_variance.  Tolerances are the reference's own (1e-3) tightened to the printed digits."""
import math

import numpy as np
import pytest

from oracle import gpr
from oracle.gpr import BoundedValue, ConstantKernel, Matern, Product


def bv(v, lo, hi):
    return BoundedValue(v, lo, hi)


X3 = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, 2.0]])


@pytest.mark.parametrize("A,tol", [(np.float64, 6e-9), (np.float32, 2e-7)])
def test_matern_nu_3_2_golden(A, tol):
    # src/gpr/matern_kernel.rs:189-220
    kernel = Matern(1.5, [bv(1.0, 0.05, 20.0), bv(1.0, 0.05, 20.0)])
    km = np.array([[1.0, 0.29782077, 0.1013397], [0.29782077, 1.0, 0.48335772], [0.1013397, 0.48335772, 1.0]])
    gm = np.array([
        [[0.0, 0.0], [0.25901289, 0.25901289], [0.0623887, 0.24955481]],
        [[0.25901289, 0.25901289], [0.0, 0.0], [0.0, 0.53076362]],
        [[0.0623887, 0.24955481], [0.0, 0.53076362], [0.0, 0.0]],
    ])
    k, g = kernel.theta_grad(X3.astype(A), A)
    assert k.dtype == A and g.dtype == A
    np.testing.assert_allclose(k, km, atol=tol)
    np.testing.assert_allclose(g, gm, atol=tol)
    np.testing.assert_allclose(kernel.diag(X3, A), np.diag(k), atol=1e-7)


@pytest.mark.parametrize("A,tol", [(np.float64, 6e-9), (np.float32, 2e-7)])
def test_matern_nu_5_2_golden(A, tol):
    # src/gpr/matern_kernel.rs:222-253
    kernel = Matern(2.5, [bv(1.0, 0.05, 20.0), bv(1.0, 0.05, 20.0)])
    km = np.array([[1.0, 0.31728336, 0.09657724], [0.31728336, 1.0, 0.52399411], [0.09657724, 0.52399411, 1.0]])
    gm = np.array([
        [[0.0, 0.0], [0.29364328, 0.29364328], [0.06737947, 0.26951788]],
        [[0.29364328, 0.29364328], [0.0, 0.0], [0.0, 0.57644039]],
        [[0.06737947, 0.26951788], [0.0, 0.57644039], [0.0, 0.0]],
    ])
    k, g = kernel.theta_grad(X3.astype(A), A)
    np.testing.assert_allclose(k, km, atol=tol)
    np.testing.assert_allclose(g, gm, atol=tol)


@pytest.mark.parametrize("A,rtol", [(np.float64, 6e-9), (np.float32, 3e-6)])
def test_product_golden(A, rtol):
    # src/gpr/product_kernel.rs:120-169
    kernel = Product(ConstantKernel(bv(2.0, 1.0, 5.0)),
                     Matern(2.5, [bv(1.0, 0.05, 20.0), bv(1.0, 0.05, 20.0)]))
    x = np.array([[0.5, 7.8], [3.3, 1.4], [3.9, 5.6]])
    km = np.array([
        [2.00000000e+00, 3.22221679e-05, 8.73105609e-03],
        [3.22221679e-05, 2.00000000e+00, 6.14136045e-03],
        [8.73105609e-03, 6.14136045e-03, 2.00000000e+00],
    ])
    gm = np.array([
        [[2.00000000e+00, 0.0, 0.0], [3.22221679e-05, 7.14401245e-05, 3.73238201e-04],
         [8.73105609e-03, 4.52409267e-02, 1.89417029e-02]],
        [[3.22221679e-05, 7.14401245e-05, 3.73238201e-04], [2.00000000e+00, 0.0, 0.0],
         [6.14136045e-03, 9.54435058e-04, 4.67673178e-02]],
        [[8.73105609e-03, 4.52409267e-02, 1.89417029e-02], [6.14136045e-03, 9.54435058e-04, 4.67673178e-02],
         [2.00000000e+00, 0.0, 0.0]],
    ])
    k, g = kernel.theta_grad(x.astype(A), A)
    np.testing.assert_allclose(k, km, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(g, gm, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(kernel.diag(x, A), np.diag(k), atol=1e-7)
    assert kernel.n_params() == 3
    assert kernel.theta() == [math.log(2.0), 0.0, 0.0]


def test_cdist_exact():
    # src/gpr/matern_kernel.rs:285-305
    assert gpr.cdist(np.array([[1.0, 3.0]]), np.array([[2.0, 5.0]]))[0, 0] == math.sqrt(5.0)
    a = np.array([[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]])
    b = np.array([[1.0, 2.0], [3.0, 4.0]])
    exp = np.sqrt(np.array([[5.0, 25.0], [1.0, 13.0], [1.0, 5.0]]))
    assert (gpr.cdist(a, b) == exp).all()


def test_outer_exact():
    # src/gpr/lml.rs:105-119
    assert (gpr.outer(np.array([-1.0, 1.0]), np.array([1.0, 2.0, 3.0]))
            == np.array([[-1.0, -2.0, -3.0], [1.0, 2.0, 3.0]])).all()
    assert (gpr.outer(np.array([-1.0, 1.0]), np.array([3.0, 7.0])) == np.array([[-3.0, -7.0], [3.0, 7.0]])).all()


def test_clamp_negative_variance():
    # src/gpr/predict.rs:129-149
    v = np.array([1.0, -2.0, -0.5])
    assert gpr.clamp_negative_variance(v, -1.0) == [-2.0]
    assert (v == np.array([1.0, 0.0, 0.0])).all()
    v = np.array([1.0, 2.0, -0.5])
    assert gpr.clamp_negative_variance(v, -1.0) is None
    assert (v == np.array([1.0, 2.0, 0.0])).all()


def test_bounded_value():
    # src/util/bounded_value.rs
    b = bv(1.0, 0.5, 2.0)
    assert b.with_clamped_value(3.0).value == 2.0
    assert b.with_clamped_value(0.1).value == 0.5
    assert b.with_clamped_value(1.5).value == 1.5
    with pytest.raises(gpr.BoundsError):
        b.with_value(2.5)
    assert b.with_value(2.0).value == 2.0  # inclusive


def _problem(n=40, d=3, seed=0, A=np.float64):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1) + 0.1 * rng.standard_normal(n)
    kernel = Product(ConstantKernel(bv(1.3, 1e-3, 1e3)),
                     Matern(2.5, [bv(0.4 + 0.2 * k, 1e-3, 1e3) for k in range(d)]))
    return kernel, x.astype(A), y.astype(A)


@pytest.mark.parametrize("nu", [0.5, 1.5, 2.5])
def test_lml_gradient_matches_finite_differences(nu):
    # self-consistency (the reference pins neither LML nor its gradient: SURVEY.md 8c-3)
    kernel, x, y = _problem()
    kernel = Product(kernel.k1, Matern(nu, kernel.k2.length_scale))
    theta = np.array([math.log(0.05)] + kernel.theta())

    def f(t):
        k = kernel.with_clamped_theta(list(t[1:]))
        return gpr.lml_with_gradient(k, math.exp(t[0]), x, y)

    res = f(theta)
    for i in range(len(theta)):
        h = 1e-6
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h
        tm[i] -= h
        fd = (f(tp).lml - f(tm).lml) / (2 * h)
        assert abs(fd - res.lml_gradient[i]) <= 2e-6 * max(1.0, abs(fd)), (i, fd, res.lml_gradient[i])


def test_kinv_is_inverse_and_predict_interpolates():
    kernel, x, y = _problem()
    res = gpr.lml_with_gradient(kernel, 1e-6, x, y)
    kinv = res.factorization.invc()
    k = kernel.kernel(x, x) + 1e-6 * np.eye(len(y))
    np.testing.assert_allclose(kinv @ k, np.eye(len(y)), atol=1e-6)
    var = np.zeros(len(y))
    mean = gpr.predict(kernel, res.alpha, x, x, kinv, var)
    np.testing.assert_allclose(mean, y, atol=1e-4)
    assert (var >= 0).all() and (var < 1e-3).all()


def test_nd_sum_matches_plain_sum():
    rng = np.random.default_rng(1)
    for n in [0, 1, 7, 8, 9, 1000, 1003]:
        a = rng.standard_normal(n)
        assert abs(gpr.nd_sum(a) - a.sum()) <= 1e-12 * max(1.0, np.abs(a).sum())
