"""BASELINE-size checks through size-independent properties (the oracle needs ~10 s and ~15 GB per
evaluation at n = 4096, so it is not run here): the north-star shape n = 4096, d = 16."""
import math

import numpy as np
import pytest

from oracle import gpr as ogpr
from tests.util import oracle_kernel, random_thetas, synth

pytestmark = pytest.mark.gpu

N, D = 4096, 16


@pytest.fixture(scope="module")
def problem():
    import hbetune_rs_b200 as h
    x, y = synth(N, D)
    ctx = h.Context()
    ctx.set_data(x, y)
    yield ctx, x, y
    ctx.close()


def test_gradient_matches_finite_differences_of_the_gpu_lml(problem):
    ctx, x, y = problem
    theta = np.array([math.log(0.05), math.log(1.2)] + [math.log(0.6 + 0.05 * k) for k in range(D)])
    lml, grad, status = ctx.lml_grad_batch(theta[None])
    assert status[0] == 0
    h = 1e-5
    probes = []
    for i in (0, 1, 2, D + 1):
        for s in (+1, -1):
            t = theta.copy()
            t[i] += s * h
            probes.append(t)
    vals, _, st = ctx.lml_grad_batch(np.array(probes), want_grad=False)
    assert (st == 0).all()
    for j, i in enumerate((0, 1, 2, D + 1)):
        fd = (vals[2 * j] - vals[2 * j + 1]) / (2 * h)
        assert abs(fd - grad[0, i]) <= 1e-5 * max(1.0, abs(grad[0, i])), (i, fd, grad[0, i])


def test_batch_composition_does_not_change_results(problem):
    ctx, x, y = problem
    thetas = random_thetas(6, D, seed=4)
    lml_all, grad_all, _ = ctx.lml_grad_batch(thetas)
    lml_one, grad_one, _ = ctx.lml_grad_batch(thetas[3:4])
    assert lml_one[0] == lml_all[3] and (grad_one[0] == grad_all[3]).all()  # bit-identical: fixed-order reductions


def test_inverse_round_trip_and_alpha(problem):
    ctx, x, y = problem
    theta = np.array([math.log(0.05), 0.0] + [math.log(0.7)] * D)
    model = ctx.model(theta, want_kinv=True)
    k = oracle_kernel(theta).kernel(x, x) + 0.05 * np.eye(N)
    assert np.abs(model.k_inv @ k - np.eye(N)).max() < 1e-8
    assert np.abs(k @ model.alpha - y).max() < 1e-8
    sign, logdet = np.linalg.slogdet(k)
    lml_ref = -0.5 * float(y @ model.alpha) - 0.5 * logdet - N / 2 * math.log(2 * math.pi)
    assert abs(model.lml - lml_ref) <= 1e-10 * abs(lml_ref)
    model.close()


def test_predict_chunking_and_subset_against_oracle(problem):
    ctx, x, y = problem
    theta = np.array([math.log(0.1), 0.0] + [math.log(0.5)] * D)  # the C4 model
    model = ctx.model(theta, want_kinv=True)
    m = 70_001  # crosses the 32768-row chunk boundary twice, ragged tail
    xs = np.random.default_rng(2).random((m, D))
    mean, var = model.predict(xs)
    pick = np.array([0, 1, 32767, 32768, 65535, 65536, m - 1])
    mean_s, var_s = model.predict(xs[pick])
    # same rows through the throughput path (chunked GEMM) and through the small-batch latency path: the
    # summation orders differ, the values agree to rounding
    np.testing.assert_allclose(mean[pick], mean_s, rtol=0, atol=1e-13 * np.abs(mean).max())
    np.testing.assert_allclose(var[pick], var_s, rtol=0, atol=1e-13)
    again_mean, again_var = model.predict(xs)
    np.testing.assert_array_equal(again_mean, mean)  # run-to-run bit reproducible
    np.testing.assert_array_equal(again_var, var)
    kern = oracle_kernel(theta)
    var_ref = np.zeros(len(pick))
    mean_ref = ogpr.predict(kern, model.alpha, xs[pick], x, model.k_inv, var_ref)
    np.testing.assert_allclose(mean_s, mean_ref, rtol=0, atol=1e-9 * np.abs(mean_ref).max())
    np.testing.assert_allclose(var_s, var_ref, rtol=0, atol=1e-9 * (1.0 + 1e-5))
    assert (var >= 0).all() and (var <= 1.0 + 1e-5 + 1e-12).all()
    model.close()
