"""Independent pin of the oracle AND of the CUDA path against scikit-learn's GaussianProcessRegressor.

The reference's gpr module restates sklearn's GP (its own golden vectors were produced with sklearn,
``src/gpr/matern_kernel.rs:189-253``): the same kernel ``ConstantKernel * Matern(nu) + WhiteKernel``, the same
log-space parameters, hence the same log-marginal likelihood and gradient; predictions differ only in the
variance's noise term (the reference adds 1e-5 where sklearn's WhiteKernel adds the fitted noise, predict.rs:25-37).
sklearn is part of this image (here and on the GPU box); the tests are skipped where it is missing."""
import math

import numpy as np
import pytest

from oracle import gpr as ogpr
from tests.util import oracle_kernel, oracle_lml, synth

skl = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel  # noqa: E402

CASES = [(25, 1, 0.5), (60, 3, 1.5), (60, 3, 2.5), (130, 5, 2.5), (200, 8, 0.5)]


def _theta(d, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([[math.log(rng.uniform(0.02, 0.5)), math.log(rng.uniform(0.5, 3.0))],
                           np.log(rng.uniform(0.3, 2.5, d))])


def _sklearn(theta, nu, x, y):
    k = ConstantKernel(math.exp(theta[1])) * Matern(length_scale=np.exp(theta[2:]), nu=nu) + WhiteKernel(math.exp(theta[0]))
    gp = skl.GaussianProcessRegressor(kernel=k, optimizer=None, alpha=0.0).fit(x, y)
    lml, g = gp.log_marginal_likelihood(np.concatenate([theta[1:], theta[:1]]), eval_gradient=True)
    return gp, lml, np.concatenate([g[-1:], g[:-1]])  # our order: noise first


@pytest.mark.parametrize("n,d,nu", CASES)
def test_oracle_lml_gradient_and_prediction_match_sklearn(n, d, nu):
    x, y = synth(n, d, seed=n + d)
    theta = _theta(d, n)
    gp, lml, grad = _sklearn(theta, nu, x, y)
    res = oracle_lml(theta, x, y, nu=nu)
    assert res.lml == pytest.approx(lml, rel=1e-11, abs=1e-11)
    np.testing.assert_allclose(res.lml_gradient, grad, rtol=1e-8, atol=1e-9 * np.abs(grad).max())
    xs = np.random.default_rng(1).random((40, d))
    mean, std = gp.predict(xs, return_std=True)
    var = np.empty(len(xs))
    omean = ogpr.predict(oracle_kernel(theta, nu), res.alpha, xs, x, res.factorization.invc(), var)
    np.testing.assert_allclose(omean, mean, rtol=0, atol=1e-10 * max(1.0, np.abs(mean).max()))
    # sklearn: var = c + noise - k* K^-1 k*^T; reference: c + 1e-5 - k* K^-1 k*^T (both clamped at 0)
    want = np.maximum(std ** 2 - math.exp(theta[0]) + 1e-5, 0.0)
    np.testing.assert_allclose(var, want, rtol=0, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,nu", CASES)
def test_cuda_lml_gradient_and_prediction_match_sklearn(n, d, nu):
    import hbetune_rs_b200 as h
    x, y = synth(n, d, seed=n + d)
    theta = _theta(d, n)
    gp, lml, grad = _sklearn(theta, nu, x, y)
    with h.Context() as ctx:
        ctx.set_data(x, y)
        got_lml, got_grad, status = ctx.lml_grad_batch(theta[None, :], nu)
        assert status[0] == 0
        assert got_lml[0] == pytest.approx(lml, rel=1e-9, abs=1e-9)
        np.testing.assert_allclose(got_grad[0], grad, rtol=1e-7, atol=1e-9 * np.abs(grad).max())
        model = ctx.model(theta, nu)
        xs = np.random.default_rng(1).random((40, d))
        mean, std = gp.predict(xs, return_std=True)
        gmean, gvar = model.predict(xs)
        np.testing.assert_allclose(gmean, mean, rtol=0, atol=1e-9 * max(1.0, np.abs(mean).max()))
        want = np.maximum(std ** 2 - math.exp(theta[0]) + 1e-5, 0.0)
        np.testing.assert_allclose(gvar, want, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_fit_reaches_the_optimum_sklearn_finds():
    """The fitted hyper-parameters cannot be compared run for run (NLopt / L-BFGS-B / this library's L-BFGS are three
    different optimisers), but with the same bounds and enough restarts all must reach the same maximum of the LML."""
    import hbetune_rs_b200 as h
    n, d = 80, 2
    x, y = synth(n, d, seed=21)
    bv = h.BoundedValue
    kernel = h.Product(h.ConstantKernel(bv(1.0, 1e-2, 1e2)), h.Matern(2.5, [bv(1.0, 1e-2, 1e2)] * d))
    with h.Context() as ctx:
        fk = h.FittedKernel.new(ctx, kernel, x, y, h.RNG.new_with_seed(5), 10, bv(1.0, 1e-3, 1e1))
    k = (ConstantKernel(1.0, (1e-2, 1e2)) * Matern(length_scale=[1.0] * d, length_scale_bounds=(1e-2, 1e2), nu=2.5)
         + WhiteKernel(1.0, (1e-3, 1e1)))
    gp = skl.GaussianProcessRegressor(kernel=k, alpha=0.0, n_restarts_optimizer=10, random_state=0).fit(x, y)
    sk_lml = gp.log_marginal_likelihood_value_
    assert fk.lml >= sk_lml - 1e-5 * abs(sk_lml)
    assert fk.lml <= sk_lml + 1e-3 * abs(sk_lml) + 1e-3  # same maximum, not a different basin
    th_sk = gp.kernel_.theta  # [ln c, ln l_1.., ln noise]
    ours = np.array([math.log(fk.kernel.k1.constant.value)] + [math.log(l.value) for l in fk.kernel.k2.length_scale]
                    + [math.log(fk.noise.value)])
    np.testing.assert_allclose(ours, th_sk, atol=5e-3)
