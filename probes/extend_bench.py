"""Timing of hbegp_model_extend: block append vs full evaluation."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402


def synth(n, d, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1) + 0.1 * rng.standard_normal(n)
    y = y - y.min()
    return x, y / y.mean() + 0.05


for n_old, k, d in ((1024, 16, 8), (4096, 16, 16), (4096, 64, 16), (4096, 512, 16), (8192, 64, 16)):
    x, y = synth(n_old + k, d)
    th = np.array([math.log(0.05), 0.0] + [math.log(1.5)] * d)
    ctx = h.Context(0, h.F64)
    ctx.set_data(x[:n_old], y[:n_old])
    prior = ctx.model(th)
    ctx.set_data(x, y)
    res = {}
    for name, mk in (("append", lambda: h.Model(ctx, prior=prior, want_alpha=False)),
                     ("full", lambda: ctx.model(th, want_alpha=False))):
        mk()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            m = mk()
        res[name] = (time.perf_counter() - t0) / reps * 1e3
    xs = np.random.default_rng(3).random((64, d))
    a, f = h.Model(ctx, prior=prior), ctx.model(th)
    dm = np.abs(a.predict(xs)[0] - f.predict(xs)[0]).max()
    dv = np.abs(a.predict(xs)[1] - f.predict(xs)[1]).max()
    print(f"n_old={n_old} +{k} d={d}: append {res['append']:.2f} ms  full {res['full']:.2f} ms  "
          f"(x{res['full'] / res['append']:.1f})  appended={a.appended}  |dmean|={dm:.2e} |dvar|={dv:.2e}", flush=True)
