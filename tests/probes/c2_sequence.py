"""Probe: BASELINE config 2 shaped run — rosenbrock 8-D, max-nevals 500, --transform-objective log, population 10,
the GP refitted every generation from the previous model (src/core/minimize.rs:465-499).  The EA is out of scope;
a seeded sampler supplies the 10 new points per generation.  Reports the GP time of the whole run on the GPU and
the oracle's CPU time for a bounded sample of generations."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hbetune_rs_b200 as h
from oracle import adapter as oad
from oracle.rng import RNG
from tests.util import lib_minimizer

d, pop, nevals = 8, 10, 500
rng_np = np.random.default_rng(1)


def f(x):
    z = x * 10 - 5
    return (100 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1 - z[:, :-1]) ** 2).sum(axis=1)


xs_all = rng_np.random((nevals, d))
ys_all = f(xs_all)
est = h.EstimatorGPR(d).y_projection(h.LOGARITHMIC)
rng = RNG.new_with_seed(1)
model = None
gpu_times, evals, predict_times = [], [], []
probe = rng_np.random((1610, d))  # ~ the 1610 single-point predictions per generation of the reference
for gen in range(nevals // pop):
    n = (gen + 1) * pop
    t0 = time.perf_counter()
    model = est.estimate(xs_all[:n], ys_all[:n], model, rng)
    gpu_times.append(time.perf_counter() - t0)
    evals.append(model.fitted.n_evals)
    t0 = time.perf_counter()
    for i in range(0, 200):
        model.predict_mean_ei(probe[i], float(ys_all[:n].min()))
    predict_times.append((time.perf_counter() - t0) / 200 * 1610)
out = {"generations": len(gpu_times), "gpu_fit_s_total": sum(gpu_times), "gpu_fit_s_last": gpu_times[-1],
       "evals_total": int(sum(evals)), "gpu_predict_s_total_1610_calls_per_gen": sum(predict_times),
       "final_lml": model.lml, "final_length_scales": model.length_scales()}
# bounded CPU sample: the oracle's estimate() for a few generations (same optimiser, fresh start, no prior)
oest = oad.EstimatorGPR(d)
oest.y_projection = "logarithmic"
cpu = {}
for n in (50, 200, 500):
    t0 = time.perf_counter()
    oest.estimate(xs_all[:n], ys_all[:n], None, RNG.new_with_seed(1), lib_minimizer())
    cpu[n] = time.perf_counter() - t0
    t0 = time.perf_counter()
    est.estimate(xs_all[:n], ys_all[:n], None, RNG.new_with_seed(1))
    cpu[f"gpu_{n}"] = time.perf_counter() - t0
out["fresh_fit_seconds_oracle_vs_gpu"] = cpu
print(json.dumps(out))
