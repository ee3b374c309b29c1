"""EstimatorGPR / SurrogateModelGPR (src/core/gpr.rs) on the GPU against the oracle's adapter restatement,
both driven by the same bounded L-BFGS and the same RNG stream (seeded like tests/gpr_tests.rs)."""
import math

import numpy as np
import pytest

from oracle import adapter as oad
from oracle.rng import RNG
from tests.util import lib_minimizer

pytestmark = pytest.mark.gpu


def _data(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = ((x - 0.4) ** 2).sum(axis=1) * 30 + 5 + 0.3 * rng.standard_normal(n)  # noisy sphere, natural units
    return x, y


@pytest.mark.parametrize("proj", ["linear", "logarithmic"])
def test_estimate_predict_extend_match_oracle(proj):
    import hbetune_rs_b200 as h
    n, d = 80, 2
    x, y = _data(n, d, 4531)
    est = h.EstimatorGPR(d).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(1)
    est.y_projection(h.LINEAR if proj == "linear" else h.LOGARITHMIC)
    oest = oad.EstimatorGPR(d)
    oest.noise_bounds, oest.n_restarts_optimizer, oest.y_projection = (1e-2, 1e1), 1, proj
    model = est.estimate(x, y, None, RNG.new_with_seed(123), want_kinv=True)
    omodel = oest.estimate(x, y, None, RNG.new_with_seed(123), lib_minimizer())
    assert abs(model.lml - omodel.lml) <= 1e-7 * abs(omodel.lml)
    np.testing.assert_allclose(model.length_scales(), omodel.length_scales(), rtol=1e-4)
    xs = np.random.default_rng(9).random((40, d))
    np.testing.assert_allclose(model.predict_mean_a(xs), omodel.predict_mean_a(xs), rtol=1e-6)
    fmin = float(y.min())
    mean, ei = model.predict_mean_ei_a(xs, fmin)
    omean, oei = omodel.predict_mean_ei_a(xs, fmin)
    np.testing.assert_allclose(mean, omean, rtol=1e-6)
    np.testing.assert_allclose(ei, oei, rtol=1e-4, atol=1e-9)
    assert (ei >= 0).all()
    m1, e1 = model.predict_mean_ei(xs[3], fmin)
    assert m1 == mean[3] and e1 == ei[3]
    assert model.predict_mean(xs[5]) == pytest.approx(omodel.predict_mean_a(xs[5:6])[0], rel=1e-6)
    assert model.predict_confidence_bound(xs[0], 1.5) == pytest.approx(omodel.predict_confidence_bound(xs[0], 1.5), rel=1e-5)
    st = model.predict_statistics(xs[0])
    assert st.q1 < st.q2 < st.q3 and st.std > 0 and st.iqr() == st.q3 - st.q1
    if proj == "linear":
        assert st.q2 == pytest.approx(st.mean, rel=1e-12)  # symmetric in natural units
        assert st.cv == pytest.approx(st.std / st.mean, rel=1e-12)
    # extend: same theta, more data, no optimisation (gpr.rs:293-337, fit.rs:33-68)
    x2, y2 = _data(n + 10, d, 77)
    ext = est.extend(x2, y2, model)
    oext = oest.extend(x2, y2, omodel)
    assert ext.length_scales() == model.length_scales()
    assert ext.lml == pytest.approx(oext.lml, rel=1e-6)
    np.testing.assert_allclose(ext.predict_mean_a(xs), oext.predict_mean_a(xs), rtol=1e-6)
    # prior reuse: estimate() starting from the prior's kernel and noise (gpr.rs:402-409)
    again = est.estimate(x2, y2, ext, RNG.new_with_seed(5))
    assert again.lml >= ext.lml - 1e-9


def test_bounds_errors_surface_like_the_reference():
    import hbetune_rs_b200 as h
    x, y = _data(20, 2, 1)
    with pytest.raises(h.estimator.NoiseBounds):
        h.EstimatorGPR(2).with_noise_bounds(2.0, 5.0).estimate(x, y, None, RNG.new_with_seed(1))  # start 1.0 not in bounds
    with pytest.raises(h.estimator.LengthScaleBounds):
        h.EstimatorGPR(2).with_length_scale_bounds([(2.0, 1.0)] * 2).estimate(x, y, None, RNG.new_with_seed(1))


def test_fit_f32_close_to_f64():
    """--use-32 (main.rs:240-244): f32 data and linear algebra, f64 hyper-parameters.  The two fits walk
    different trajectories, so this is a behavioural check (like tests/gpr_tests.rs), not parity."""
    import hbetune_rs_b200 as h
    x, y = _data(100, 2, 9372)
    xs = np.random.default_rng(0).random((30, 2))
    preds = {}
    for A in (np.float64, np.float32):
        est = h.EstimatorGPR(2, dtype=A).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(1)
        model = est.estimate(x.astype(A), y.astype(A), None, RNG.new_with_seed(123))
        preds[A] = model.predict_mean_a(xs.astype(A))
        assert preds[A].dtype == A
    assert np.abs(preds[np.float32] - preds[np.float64]).max() < 0.05 * np.abs(preds[np.float64]).max()


def test_gpr_behaviour_1d_like_reference_suite():
    """tests/gpr_tests.rs:72-225 style: a 1-D fit reproduces its data and is more certain near data."""
    import hbetune_rs_b200 as h
    xs = np.array([0.1, 0.5, 0.5, 0.9])[:, None]
    ys = np.array([1.0, 1.8, 2.2, 3.0])
    est = h.EstimatorGPR(1).with_noise_bounds(1e-3, 1e1).n_restarts_optimizer(2)
    model = est.estimate(xs, ys, None, RNG.new_with_seed(123))
    got = model.predict_mean_a(np.array([[0.1], [0.5], [0.9]]))
    np.testing.assert_allclose(got, [1.0, 2.0, 3.0], atol=0.15)
    near = model.predict_statistics(np.array([0.5])).std
    far = model.predict_statistics(np.array([0.0])).std
    assert near < far


@pytest.mark.parametrize("proj", ["linear", "logarithmic"])
def test_device_acquisition_epilogues_match_host_path(proj):
    """Rows f1/f2: EI, de-normalisation and the arg-best picks on the device agree with the host-side path
    (which is itself checked against the oracle above), including the reference's tie-breaking."""
    import hbetune_rs_b200 as h
    n, d = 60, 2
    x, y = _data(n, d, 11)
    est = h.EstimatorGPR(d).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(1)
    est.y_projection(h.LINEAR if proj == "linear" else h.LOGARITHMIC)
    model = est.estimate(x, y, None, RNG.new_with_seed(3))
    xs = np.random.default_rng(1).random((2000, d))
    xs[1500] = xs[10]  # exact duplicates: identical EI / bound -> tie-breaking is exercised
    xs[1999] = xs[10]
    fmin = float(np.sort(y)[3])
    mean_h, ei_h = model.predict_mean_ei_a(xs, fmin)
    mean_d, ei_d, best = model.predict_mean_ei_device(xs, fmin)
    np.testing.assert_allclose(mean_d, mean_h, rtol=4e-16)  # exp() differs by an ulp between libm and CUDA
    # far-tail EI values cancel (-(mu - fmin) Phi(z) + sigma phi(z)): CUDA's and libm's erfc/exp differ by an ulp there
    np.testing.assert_allclose(ei_d, ei_h, rtol=1e-9, atol=1e-14 * ei_h.max())
    last_max = len(ei_d) - 1 - int(np.argmax(ei_d[::-1]))  # Iterator::max_by returns the last maximum
    assert best == last_max
    cb_d, best_cb = model.predict_confidence_bound_device(xs, 1.0)
    for i in (0, 10, 1500, 1999):
        assert cb_d[i] == pytest.approx(model.predict_confidence_bound(xs[i], 1.0), rel=1e-12)
    assert best_cb == int(np.argmin(cb_d))  # first minimum
    assert cb_d[1500] == cb_d[10] == cb_d[1999]
    # force a tie for the arg-best: evaluate only the duplicated rows
    dup = xs[[10, 1500, 1999]]
    _, ei3, b3 = model.predict_mean_ei_device(dup, fmin)
    _, bcb3 = model.predict_confidence_bound_device(dup, 1.0)
    assert ei3[0] == ei3[1] == ei3[2] and b3 == 2 and bcb3 == 0


@pytest.mark.parametrize("objective,d,gens,log_y", [("sphere", 2, 10, False), ("rosenbrock", 8, 6, True)])
def test_generation_sequence_matches_oracle(objective, d, gens, log_y):
    """BASELINE configs 1-2 shaped (sphere 2-D popsize 10; rosenbrock 8-D with --transform-objective log): the
    minimizer refits the GP every generation starting from the previous model's kernel and noise
    (src/core/minimize.rs:465-499 -> gpr.rs:402-409).  The EA itself is out of scope; a fixed seeded sampler
    stands in for it.  Per generation the chosen hyper-parameters, the LML and the predictions of the GPU
    estimator and of the oracle estimator (same optimiser, same RNG stream) must agree."""
    import hbetune_rs_b200 as h
    rng_np = np.random.default_rng(1)

    def f(x):
        z = x * 10 - 5
        if objective == "sphere":
            return (z ** 2).sum(axis=1)
        return (100 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1 - z[:, :-1]) ** 2).sum(axis=1)

    est = h.EstimatorGPR(d).with_noise_bounds(1e-3, 1e1)
    oest = oad.EstimatorGPR(d)
    oest.noise_bounds = (1e-3, 1e1)
    if log_y:
        est.y_projection(h.LOGARITHMIC)
        oest.y_projection = "logarithmic"
    rng_g, rng_o = RNG.new_with_seed(1), RNG.new_with_seed(1)
    x = rng_np.random((10, d))
    model = omodel = None
    probe = rng_np.random((25, d))
    for gen in range(gens):
        y = f(x)
        model = est.estimate(x, y, model, rng_g)
        omodel = oest.estimate(x, y, omodel, rng_o, lib_minimizer())
        assert abs(model.lml - omodel.lml) <= 1e-6 * max(1.0, abs(omodel.lml)), (gen, model.lml, omodel.lml)
        np.testing.assert_allclose(np.log(model.length_scales()), np.log(omodel.length_scales()), atol=2e-3)
        assert math.log(model.noise.value) == pytest.approx(math.log(omodel.noise.value), abs=2e-3)
        np.testing.assert_allclose(model.predict_mean_a(probe), omodel.predict_mean_a(probe), rtol=1e-4, atol=1e-6)
        # "the same selected EA individuals": the acquisition's pick among a pool of candidates (max EI, the LAST maximum
        # wins, acquisition.rs:177-202) and the suggestion's pick (lowest confidence bound, the FIRST minimum,
        # minimize.rs:680-714) agree between the GPU model and the oracle model
        pool = rng_np.random((300, d))
        fmin = float(y.min())
        best_gpu, _, _ = h.find_best_candidate_by_ei(pool, model, fmin)
        _, ei_o = omodel.predict_mean_ei_a(pool, fmin)
        best_o = len(ei_o) - 1 - int(np.argmax(np.asarray(ei_o)[::-1]))
        assert best_gpu == best_o, (gen, best_gpu, best_o)
        cb_gpu, _ = h.find_best_individual_by_confidence_bound(pool, model, 1.0)
        cb_o = np.array([omodel.predict_confidence_bound(c, 1.0) for c in pool])
        assert cb_gpu == int(np.argmin(cb_o)), (gen, cb_gpu, int(np.argmin(cb_o)))
        x = np.concatenate([x, rng_np.random((10, d))])


def test_callers_batched_selection_matches_the_one_point_per_call_loops():
    """find_best_candidate_by_ei (acquisition.rs:177-202, LAST maximum) and
    find_best_individual_by_confidence_bound (minimize.rs:680-714, FIRST minimum) as one device pass each, against
    the reference's pattern -- a Python loop of single-point predictions with the reference's comparison rules."""
    import hbetune_rs_b200 as h
    n, d = 90, 3
    x, y = _data(n, d, 11)
    est = h.EstimatorGPR(d).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(1)
    model = est.estimate(x, y, None, RNG.new_with_seed(3))
    rng = np.random.default_rng(4)
    cand = rng.random((200, d))
    cand[150] = cand[20]  # exact ties: duplicates of the eventual winners are planted below
    fmin = float(y.min())
    # -- EI: loop of predict_mean_ei, max_by semantics (last maximum wins)
    means, eis = zip(*(model.predict_mean_ei(c, fmin) for c in cand))
    best_loop = 0
    for i in range(1, len(eis)):
        if not eis[i] < eis[best_loop]:  # max_by returns the last element among equals
            best_loop = i
    cand2 = np.concatenate([cand, cand[best_loop:best_loop + 1]])  # a later duplicate of the winner must win
    i, mean_i, ei_i = h.find_best_candidate_by_ei(cand2, model, fmin)
    assert i == len(cand2) - 1
    assert mean_i == pytest.approx(means[best_loop], rel=1e-9) and ei_i == pytest.approx(eis[best_loop], rel=1e-7, abs=1e-12)
    i, _, _ = h.find_best_candidate_by_ei(cand, model, fmin)
    assert i == best_loop
    # -- confidence bound: loop of predict_confidence_bound, strict < (first minimum stays)
    cb = 1.3
    ucb = [model.predict_confidence_bound(c, cb) for c in cand]
    first = 0
    for j in range(1, len(ucb)):
        if ucb[j] < ucb[first]:
            first = j
    cand3 = np.concatenate([cand, cand[first:first + 1]])  # a later duplicate of the winner must NOT win
    j, y_j = h.find_best_individual_by_confidence_bound(cand3, model, cb)
    assert j == first
    assert y_j == pytest.approx(model.predict_mean(cand[first]), rel=1e-12)
    # -- FitnessVia::Prediction for a whole population
    np.testing.assert_allclose(h.predicted_fitness(cand[:17], model), [model.predict_mean(c) for c in cand[:17]], rtol=1e-9)
    with pytest.raises(RuntimeError):
        h.find_best_candidate_by_ei(np.empty((0, d)), model, fmin)
    with pytest.raises(RuntimeError):
        h.find_best_individual_by_confidence_bound(np.empty((0, d)), model, cb)
