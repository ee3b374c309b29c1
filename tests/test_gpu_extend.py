"""hbegp_model_extend: FittedKernel::extend (src/gpr/fit.rs:33-68) as a block append to the prior factorisation.

The append must give the same model as the full evaluation (and as the oracle's extend) whenever it is taken, and
must not be taken when the prior's rows are not an unchanged prefix of the new data."""
import math

import numpy as np
import pytest

from oracle import gpr as ogpr
from oracle.gpr import BoundedValue
from tests.util import oracle_kernel, synth

pytestmark = pytest.mark.gpu


def _ctx(A=np.float64):
    import hbetune_rs_b200 as h
    return h.Context(0, h.F64 if A == np.float64 else h.F32)


def _theta(d, noise=0.05, c=1.3, ls=0.7):
    return np.array([math.log(noise), math.log(c)] + [math.log(ls * (1 + 0.1 * k)) for k in range(d)])


def _check_same_model(ext, full, xs, A, tol):
    assert ext.lml == pytest.approx(full.lml, rel=tol, abs=tol)
    scale = np.abs(full.alpha).max()
    np.testing.assert_allclose(ext.alpha, full.alpha, rtol=0, atol=tol * scale)
    m1, v1 = ext.predict(xs)
    m2, v2 = full.predict(xs)
    np.testing.assert_allclose(m1, m2, rtol=0, atol=tol * max(1.0, np.abs(m2).max()))
    np.testing.assert_allclose(v1, v2, rtol=0, atol=tol * max(1.0, np.abs(v2).max()))


@pytest.mark.parametrize("A,tol", [(np.float64, 1e-9), (np.float32, 2e-3)])
@pytest.mark.parametrize("n_old,k_new", [(200, 10), (256, 64), (300, 500), (130, 1), (640, 0), (1000, 25)])
def test_append_equals_full_evaluation(A, tol, n_old, k_new):
    import hbetune_rs_b200 as h
    d = 3
    n = n_old + k_new
    x, y = synth(n, d, seed=11, A=A)
    th = _theta(d)
    ctx = _ctx(A)
    ctx.set_data(x[:n_old], y[:n_old])
    prior = ctx.model(th)
    # the adapter renormalises y on extend (gpr.rs:293-337): the targets of the old rows change, X does not
    y2 = (y * A(1.25) + A(0.1)).astype(A)
    ctx.set_data(x, y2)
    ext = h.Model(ctx, want_kinv=True, prior=prior)
    assert ext.appended
    full = ctx.model(th, want_kinv=True)
    xs = np.random.default_rng(5).random((77, d)).astype(A)
    _check_same_model(ext, full, xs, A, tol)
    np.testing.assert_allclose(ext.k_inv, full.k_inv, rtol=0, atol=tol * np.abs(full.k_inv).max())
    if A == np.float64:
        nz = math.exp(th[0])
        ofk = ogpr.fitted_kernel_extend(oracle_kernel(th), x, y2, BoundedValue(nz, nz / 2, nz * 2), A)
        assert ext.lml == pytest.approx(ofk.lml, rel=1e-9)
        np.testing.assert_allclose(ext.alpha, ofk.alpha, rtol=0, atol=1e-9 * np.abs(ofk.alpha).max())
        ov = np.empty(len(xs), dtype=A)
        om = ogpr.predict(ofk.kernel, ofk.alpha, xs, x, ofk.k_inv, ov, A)
        m, v = ext.predict(xs)
        np.testing.assert_allclose(m, om, rtol=0, atol=1e-9)
        np.testing.assert_allclose(v, ov, rtol=0, atol=1e-9)
    # the prior model is untouched
    ctx.set_data(x[:n_old], y[:n_old])
    again = ctx.model(th)
    m0, v0 = prior.predict(xs)
    m1, v1 = again.predict(xs)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)


def test_append_chain_of_generations():
    """Repeated extends (each from the previous result) stay on the append path and on the full result."""
    import hbetune_rs_b200 as h
    d, A = 4, np.float64
    x, y = synth(700, d, seed=3)
    th = _theta(d, noise=0.02)
    ctx = _ctx(A)
    ctx.set_data(x[:150], y[:150])
    model = ctx.model(th)
    for n in (200, 330, 331, 700):
        ctx.set_data(x[:n], y[:n])
        model = h.Model(ctx, prior=model)
        assert model.appended
    full = ctx.model(th)
    xs = np.random.default_rng(8).random((50, d))
    _check_same_model(model, full, xs, A, 1e-9)


@pytest.mark.parametrize("nu", [0.5, 1.5])
def test_append_other_matern_orders(nu):
    import hbetune_rs_b200 as h
    d, A = 2, np.float64
    x, y = synth(300, d, seed=21)
    th = _theta(d, noise=0.1)
    ctx = _ctx(A)
    ctx.set_data(x[:220], y[:220])
    prior = ctx.model(th, nu)
    ctx.set_data(x, y)
    ext = h.Model(ctx, prior=prior)
    assert ext.appended
    full = ctx.model(th, nu)
    _check_same_model(ext, full, np.random.default_rng(2).random((30, d)), A, 1e-9)


def test_no_append_when_rows_differ_or_prior_is_small():
    import hbetune_rs_b200 as h
    d, A = 3, np.float64
    x, y = synth(260, d, seed=9)
    th = _theta(d)
    xs = np.random.default_rng(4).random((20, d))
    ctx = _ctx(A)
    # (a) one old row changed in the last bit
    ctx.set_data(x[:200], y[:200])
    prior = ctx.model(th)
    x2 = x.copy()
    x2[17, 1] = np.nextafter(x2[17, 1], 2.0)
    ctx.set_data(x2, y)
    ext = h.Model(ctx, prior=prior)
    assert not ext.appended
    full = ctx.model(th)
    assert ext.lml == full.lml and np.array_equal(ext.alpha, full.alpha)
    # (b) same rows in another order
    perm = np.random.default_rng(1).permutation(200)
    x3 = np.concatenate([x[:200][perm], x[200:]])
    y3 = np.concatenate([y[:200][perm], y[200:]])
    ctx.set_data(x3, y3)
    ext = h.Model(ctx, prior=prior)
    assert not ext.appended
    m1, _ = ext.predict(xs)
    ctx.set_data(x, y)
    m2, _ = ctx.model(th).predict(xs)
    np.testing.assert_allclose(m1, m2, rtol=0, atol=1e-9)
    # (c) prior with fewer than 64 rows: no complete leaf to keep
    ctx.set_data(x[:50], y[:50])
    small = ctx.model(th)
    ctx.set_data(x, y)
    ext = h.Model(ctx, prior=small)
    assert not ext.appended
    # (d) fewer rows than the prior
    ctx.set_data(x[:100], y[:100])
    ext = h.Model(ctx, prior=prior)
    assert not ext.appended and ext.n == 100


def test_extend_argument_errors():
    import hbetune_rs_b200 as h
    from hbetune_rs_b200._lib import HbegpError
    x, y = synth(100, 2, seed=2)
    th = _theta(2)
    ctx = _ctx()
    ctx.set_data(x, y)
    prior = ctx.model(th)
    other = _ctx()
    other.set_data(x, y)
    with pytest.raises(HbegpError):
        h.Model(other, prior=prior)  # model of another context
    x3, y3 = synth(100, 3, seed=2)
    ctx.set_data(x3, y3)
    with pytest.raises(HbegpError):
        h.Model(ctx, prior=prior)  # feature count changed


def test_estimator_extend_appends_validation_samples():
    """The reference's call site (minimize.rs:629-644): history + validation samples, y renormalised."""
    import hbetune_rs_b200 as h
    from oracle.rng import RNG
    rng = np.random.default_rng(31)
    n, d = 150, 2
    x = rng.random((n + 6, d))
    y = ((x - 0.4) ** 2).sum(axis=1) * 30 + 5 + 0.3 * rng.standard_normal(n + 6)
    est = h.EstimatorGPR(d).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(1)
    model = est.estimate(x[:n], y[:n], None, RNG.new_with_seed(7))
    ext = est.extend(x, y, model)
    assert ext.fitted.model.appended
    est2 = h.EstimatorGPR(d).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(1)
    ref = h.FittedKernel.extend(est2.ctx, model.kernel, x, ext.y_train, model.noise)  # full evaluation
    assert not ref.model.appended
    xs = rng.random((25, d))
    assert ext.lml == pytest.approx(ref.lml, rel=1e-9)
    m1, v1 = ext.fitted.model.predict(xs)
    m2, v2 = ref.model.predict(xs)
    np.testing.assert_allclose(m1, m2, rtol=0, atol=1e-9)
    np.testing.assert_allclose(v1, v2, rtol=0, atol=1e-9)


def test_model_buffers_are_recycled_not_leaked():
    """The tuner replaces its model every generation: destroyed models hand their buffers to the context's pool, so
    device memory stays flat over many generations and the pool goes away with the context."""
    import torch
    import hbetune_rs_b200 as h
    d, A = 4, np.float64
    x, y = synth(900, d, seed=13)
    th = _theta(d)
    xs = np.random.default_rng(1).random((300, d))
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    ctx = _ctx(A)
    model, low = None, None
    for gen in range(30):
        n = 600 + 10 * gen
        ctx.set_data(x[:n], y[:n])
        model = ctx.model(th) if model is None or gen % 3 else h.Model(ctx, prior=model)
        model.predict(xs)
        free, _ = torch.cuda.mem_get_info()
        if gen == 5:
            low = free
    assert low - free < 64 << 20  # no growth between generation 5 and 29 beyond the slowly growing n x n factor
    model.close()
    ctx.close()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20  # everything (workspaces, pool) returned with the context


def test_size_limits_are_reported():
    import hbetune_rs_b200 as h
    from hbetune_rs_b200._lib import HbegpError
    ctx = _ctx()
    with pytest.raises(HbegpError):
        ctx.set_data(np.zeros((2, 70000)), np.zeros(2))  # the only cap on the feature count (chunks of 64 handle the rest)
    with pytest.raises(HbegpError):
        ctx.set_data(np.zeros((0, 3)), np.zeros(0))  # no rows
    x, y = synth(50, 2)
    ctx.set_data(x, y)  # the context stays usable
    assert np.isfinite(ctx.lml_grad_batch(_theta(2)[None, :])[0][0])
