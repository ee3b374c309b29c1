"""Restart loop on the GPU (hbegp_fit_runs) against the oracle's fit_kernel driven by the SAME bounded
L-BFGS (the host library's; NLopt's is absent).  Fitted thetas can only be compared
between two runs of this optimiser; the trajectory is sensitive to 1e-13 differences in LML (SURVEY.md
H4), so the tight check is on the LML reached and the looser one on theta."""
import math

import numpy as np
import pytest

from oracle import gpr as ogpr
from oracle.rng import RNG
from tests.util import lib_minimizer, oracle_lml, synth

pytestmark = pytest.mark.gpu


def _kernels(mod, d, c0=1.0):
    bv = mod.BoundedValue
    kernel = mod.Product(mod.ConstantKernel(bv(c0, 1e-2, 1e2)), mod.Matern(2.5, [bv(1.0, 1e-2, 1e2)] * d))
    noise = bv(1.0, 1e-2, 1e1)
    return kernel, noise


@pytest.mark.parametrize("n,d,restarts", [(40, 2, 2), (120, 3, 4)])
def test_fit_matches_oracle_fit(n, d, restarts):
    import hbetune_rs_b200 as h
    x, y = synth(n, d)
    ok, onoise = _kernels(ogpr, d)
    ref = ogpr.fit_kernel(ok, x, y, RNG.new_with_seed(938_274), restarts, onoise, lib_minimizer(), keep_trace=True)
    gk, gnoise = _kernels(h, d)
    with h.Context() as ctx:
        fk = h.FittedKernel.new(ctx, gk, x, y, RNG.new_with_seed(938_274), restarts, gnoise, want_kinv=True)
        # every theta the oracle's runs evaluated: LML and gradient agree to 1e-9
        thetas = np.array([t for t, _, _ in ref.trace])
        lml, grad, status = ctx.lml_grad_batch(thetas, lo=None, hi=None)
        xs = np.random.default_rng(0).random((50, d))
        var = np.zeros(50)
        mean = h.predict(fk, xs, var)
    for (t, l_ref, g_ref), l, g in zip(ref.trace, lml, grad):
        assert abs(l - l_ref) <= 1e-9 * abs(l_ref)
        np.testing.assert_allclose(g, g_ref, rtol=1e-8, atol=1e-9 * np.abs(g_ref).max())
    assert abs(fk.lml - ref.lml) <= 1e-7 * abs(ref.lml), (fk.lml, ref.lml)
    th_gpu = np.array([math.log(fk.noise.value)] + fk.kernel.theta())
    th_ref = np.array([math.log(ref.noise.value)] + ref.kernel.theta())
    np.testing.assert_allclose(th_gpu, th_ref, rtol=0, atol=1e-4)
    var_ref = np.zeros(50)
    mean_ref = ogpr.predict(ref.kernel, ref.alpha, xs, x, ref.k_inv, var_ref)
    np.testing.assert_allclose(mean, mean_ref, atol=1e-5)
    np.testing.assert_allclose(var, var_ref, atol=1e-5)


def test_reference_simple_case():
    """src/gpr/predict.rs:54-99 (seed 938_274, 4 restarts): mean ~ [0, .5, 1, 1.5, 2] +- 0.1, var ~ 0.03 +- 0.03."""
    import hbetune_rs_b200 as h
    bv = h.BoundedValue
    xs = np.array([[0.0], [0.5], [0.5], [1.0]])
    ys = np.array([0.0, 0.8, 1.2, 2.0])
    kernel = h.Product(h.ConstantKernel(bv(3.0, 0.1, 4.0)), h.Matern(2.5, [bv(1.5, 0.1, 2.0)]))
    with h.Context() as ctx:
        fk = h.FittedKernel.new(ctx, kernel, xs, ys, RNG.new_with_seed(938_274), 4, bv(1.0, 0.001, 1.0))
        px = np.array([[0.0], [0.25], [0.5], [0.75], [1.0]])
        var = np.zeros(5)
        mean = h.predict(fk, px, var)
    np.testing.assert_allclose(mean, [0.0, 0.5, 1.0, 1.5, 2.0], atol=0.1)
    np.testing.assert_allclose(var, np.full(5, 0.03), atol=0.03)


def test_capture_is_best_over_all_evaluations():
    """fit.rs:115-125: the model is the best LML seen at ANY evaluation, first one wins ties."""
    import hbetune_rs_b200 as h
    x, y = synth(60, 2)
    gk, gnoise = _kernels(h, 2)
    lo, hi = h.FittedKernel._theta_bounds(gk, gnoise)
    starts = np.array([[0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [math.log(0.5), 0.3, -0.5, 0.2]])
    with h.Context() as ctx:
        ctx.set_data(x, y)
        res, thetas = ctx.fit_runs(starts, lo, hi)
        best = h.lib.hbegp_pick_best_run(len(res), res)
        lml, _, _ = ctx.lml_grad_batch(thetas)
    for r in range(3):
        assert res[r].status == 0 and 0 <= res[r].best_eval < res[r].n_evals <= 150
        assert lml[r] == res[r].best_lml  # the captured theta reproduces the captured LML bit for bit
    assert res[0].best_lml == res[1].best_lml and res[0].n_evals == res[1].n_evals  # identical runs
    assert best == int(np.argmax([r.best_lml for r in res]))  # first of the tied maxima
    assert not (best == 1)


@pytest.mark.parametrize("n,d,restarts", [(40, 2, 2), (120, 3, 3)])
def test_fit_f32_matches_f32_oracle_fit(n, d, restarts):
    """--use-32: f32 data and linear algebra on both sides, same optimiser, same start points.  Measured
    (probes/f32_fit_check.py): LML relative 5e-8..2e-7, theta 1e-4..3e-4 (ln units), mean 5e-5..1.3e-4; the bounds
    below leave an order of magnitude for trajectory branching."""
    import hbetune_rs_b200 as h
    A = np.float32
    x, y = synth(n, d, A=A)

    def kernels(mod):
        bv = mod.BoundedValue
        return (mod.Product(mod.ConstantKernel(bv(1.0, 1e-2, 1e2)), mod.Matern(2.5, [bv(1.0, 1e-2, 1e2)] * d)),
                bv(1.0, 1e-1, 1e1))

    ok, onoise = kernels(ogpr)
    ref = ogpr.fit_kernel(ok, x, y, RNG.new_with_seed(7), restarts, onoise, lib_minimizer(), A=A)
    gk, gnoise = kernels(h)
    with h.Context(0, h.F32) as ctx:
        fk = h.FittedKernel.new(ctx, gk, x, y, h.RNG.new_with_seed(7), restarts, gnoise)
        xs = np.random.default_rng(0).random((30, d)).astype(A)
        var = np.zeros(30, dtype=A)
        mean = h.predict(fk, xs, var)
    assert mean.dtype == A and fk.alpha.dtype == A
    assert abs(fk.lml - ref.lml) <= 1e-5 * abs(ref.lml)
    th_gpu = np.array([math.log(fk.noise.value)] + fk.kernel.theta())
    th_ref = np.array([math.log(ref.noise.value)] + ref.kernel.theta())
    np.testing.assert_allclose(th_gpu, th_ref, rtol=0, atol=3e-3)
    var_ref = np.zeros(30, dtype=A)
    mean_ref = ogpr.predict(ref.kernel, ref.alpha, xs, x, ref.k_inv, var_ref, A)
    np.testing.assert_allclose(mean, mean_ref, rtol=0, atol=1e-3 * max(1.0, np.abs(mean_ref).max()))
    np.testing.assert_allclose(var, var_ref, rtol=0, atol=3e-4)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_restart_loop_matches_single_process(world):
    """hbegp_fit_runs_sharded with `world` ranks simulated by threads (one context each on the same GPU, an in-process
    sum all-reduce): every rank must return the single-process records bit for bit."""
    import threading
    import hbetune_rs_b200 as h
    x, y = synth(150, 3, seed=5)
    gk, gnoise = _kernels(h, 3)
    lo, hi = h.FittedKernel._theta_bounds(gk, gnoise)
    rng = np.random.default_rng(17)
    starts = np.stack([rng.uniform(np.log(lo), np.log(hi)) for _ in range(7)])
    with h.Context() as ctx:
        ctx.set_data(x, y)
        ref, ref_thetas = ctx.fit_runs(starts, lo, hi)
    barrier = threading.Barrier(world)
    pending, out = {}, {}

    def make_allreduce(rank):
        def ar(a):
            pending[rank] = a.copy()
            barrier.wait()
            total = sum(pending[r] for r in range(world))
            barrier.wait()
            a[:] = total
        return ar

    def worker(rank):
        with h.Context() as c:
            c.set_data(x, y)
            out[rank] = c.fit_runs(starts, lo, hi, rank=rank, world=world, allreduce=make_allreduce(rank))

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(120) for t in ts]
    assert sorted(out) == list(range(world))
    for rank in range(world):
        res, thetas = out[rank]
        for r in range(7):
            assert (res[r].best_lml, res[r].best_eval, res[r].n_evals, res[r].status, res[r].final_f) == \
                   (ref[r].best_lml, ref[r].best_eval, ref[r].n_evals, ref[r].status, ref[r].final_f)
        np.testing.assert_array_equal(thetas, ref_thetas)


def test_fit_spread_against_the_committed_oracle_fits():
    """22 seeds of the BASELINE config 3 problem family (n = 1024, d = 8, the reference's default 2 restarts): the GPU
    fit against the oracle fit (tests/golden/fit_spread_oracle.json, made on the CPU by tests/probes/fit_spread.py with
    the same bounded L-BFGS).  Trajectories branch on 1e-13 differences (evaluation counts differ by up to
    20 %), the optimum does not: measured worst case over the seeds (profiles/r02_fit_spread.json) is 9e-11 relative on
    the fitted LML and 4.8e-5 on ln theta (median 8e-6); the bounds below leave a factor of ten."""
    from tests.probes.fit_spread import run_gpu, problem
    res = run_gpu()
    assert res["seeds"] >= 20
    assert res["worst"]["d_lml_same_theta_rel"] <= 1e-9   # same theta: the north_star f64 tolerance (all seeds)
    assert res["worst"]["d_lml_rel"] <= 1e-9               # each fit's own optimum: same LML to the same tolerance
    assert res["worst"]["max_d_ln_theta"] <= 5e-4
    assert res["median"]["max_d_ln_theta"] <= 1e-4
    # Seeds whose fit ended in another basin (round 2: seed 8 flips between LML 203.32 and 235.84 when the bottom node of
    # the factorisation rounds differently in the 13th digit; the reference's own result would flip the same way between
    # BLAS builds): at most 2 of 22, and each must be a genuine optimum of the SAME objective — the oracle evaluated at
    # the GPU's theta returns the GPU's LML to 1e-9 and a gradient that is flat in every coordinate not on a bound.
    assert res["same_optimum"] >= res["seeds"] - 2
    for r in res["other_optimum"]:
        x, y = problem(r["seed"])
        ref = oracle_lml(np.array(r["theta_gpu"]), x, y)
        assert abs(ref.lml - r["lml_gpu"]) <= 1e-9 * abs(ref.lml), (r["seed"], ref.lml, r["lml_gpu"])
        g = np.abs(np.asarray(ref.lml_gradient))
        assert np.sort(g)[: len(g) - 2].max() <= 1e-2 * max(1.0, abs(ref.lml)), (r["seed"], g)
