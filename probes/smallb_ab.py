"""Whole-evaluation time at n = 4096 / 8192 for small batches (environment switches are printed with the result)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402

tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("HBEGP_"))
for n, d in ((4096, 16), (8192, 16)):
    rng = np.random.default_rng(1)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1)
    y = (y - y.min()) / (y - y.min()).mean() + 0.05
    out = []
    for B in (1, 2, 4, 9, 17):
        th = np.repeat(np.array([[math.log(0.05), 0.0] + [math.log(1.5)] * d]), B, axis=0)
        ctx = h.Context(0, h.F64)
        ctx.set_data(x, y)
        ctx.bench_phase(th, 3, 2)
        out.append(f"B={B}: {ctx.bench_phase(th, 3, 4):.2f}")
        del ctx
    print(f"[{tag}] n={n} ms: " + " | ".join(out), flush=True)
