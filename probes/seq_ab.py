"""Generation sequence at small n (the reference's loop: ten more rows per generation, refit each time): total time of the
fits with cached graphs patched in place (HBEGP_GRAPH_UPDATE=1) against rebuilt (0)."""
import os, sys, time, json, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hbetune_rs_b200 as h

d = 8
_r = np.random.default_rng(3)
x = _r.random((500, d))
y = np.sin(2 * np.pi * x).sum(axis=1) + 0.05 * _r.standard_normal(500)
y = (y - y.min()) / (y - y.min()).mean() + 0.05
lo = np.array([1e-2, 1e-2] + [1e-3] * d); hi = np.array([1e1, 1e2] + [1e3] * d)
rng = np.random.default_rng(2)
out = {}
with h.Context(0, h.F64) as ctx:
    for rep in range(3):
        t0 = time.perf_counter(); evals = 0
        for n in range(10, 501, 10):
            ctx.set_data(x[:n], y[:n])
            starts = np.log(rng.uniform(lo, hi, size=(3, d + 2)))
            res, _ = ctx.fit_runs(starts, lo, hi)
            evals += sum(r.n_evals for r in res)
        out[f"pass{rep}_s"] = round(time.perf_counter() - t0, 4); out[f"pass{rep}_evals"] = int(evals)
print(json.dumps(out))
