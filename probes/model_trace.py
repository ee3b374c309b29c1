"""Where the time of model creation / extend goes (HBEGP_TRACE_MODEL=1 prints the phases on stderr)."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402

d = 16
for n_old, k in ((4096, 64), (1024, 16)):
    rng = np.random.default_rng(1)
    x = rng.random((n_old + k, d))
    y = np.sin(2 * np.pi * x).sum(axis=1)
    y = (y - y.min()) / (y - y.min()).mean() + 0.05
    th = np.array([math.log(0.05), 0.0] + [math.log(1.5)] * d)
    ctx = h.Context(0, h.F64)
    ctx.set_data(x[:n_old], y[:n_old])
    prior = ctx.model(th)
    ctx.set_data(x, y)
    keep = []
    for i in range(4):
        t0 = time.perf_counter()
        m = ctx.model(th, want_alpha=False)
        print(f"n={n_old + k} full #{i}: {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)
    for i in range(4):
        t0 = time.perf_counter()
        m = h.Model(ctx, prior=prior, want_alpha=False)
        print(f"n={n_old + k} extend #{i}: {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)
