"""HBEGP_PAD A/B: (a) steady loop of lml_grad_batch at the C3 shape (per call, host buffers), (b) a whole fit with 32 restarts
(batch sizes shrink as runs finish: every distinct size is a graph capture unless padded)."""
import os, sys, time, json, argparse
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import hbetune_rs_b200 as h

out = {}
for n, d, r in [(1024, 8, 32), (512, 8, 32), (500, 8, 2)]:
    a = argparse.Namespace(n=n, d=d, restarts=r, m=8, dtype="f64")
    _, x, y, lo, hi, th, _ = bench.workload(a)
    with h.Context(0, h.F64) as ctx:
        ctx.set_data(x, y)
        for _ in range(12): ctx.lml_grad_batch(th)
        t0 = time.perf_counter()
        for _ in range(50): ctx.lml_grad_batch(th)
        out[f"call_n{n}_B{len(th)}_ms"] = round((time.perf_counter() - t0) / 50 * 1e3, 4)
        starts = np.log(np.random.default_rng(5).uniform(lo, hi, size=(len(th), len(lo))))
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            res, _ = ctx.fit_runs(starts, lo, hi)
            ts.append(time.perf_counter() - t0)
        out[f"fit_n{n}_R{len(th)}_s"] = [round(t, 4) for t in ts]
        out[f"fit_n{n}_evals"] = int(sum(r_.n_evals for r_ in res))
print(json.dumps(out))
