"""A/B of an environment switch on the factor+inverse phase (hbegp_bench_phase), e.g.
   HBEGP_NODE128=0 python probes/phase_ab.py   vs   HBEGP_NODE128=1 python probes/phase_ab.py"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402

tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("HBEGP_"))
for n, d, B in ((1024, 8, 33), (1024, 8, 5), (2048, 16, 17), (4096, 16, 9), (4096, 16, 65)):
    rng = np.random.default_rng(1)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1)
    y = (y - y.min()) / (y - y.min()).mean() + 0.05
    th = np.repeat(np.array([[math.log(0.05), 0.0] + [math.log(1.5)] * d]), B, axis=0)
    ctx = h.Context(0, h.F64)
    ctx.set_data(x, y)
    ctx.bench_phase(th, 3, 2)
    t = [ctx.bench_phase(th, ph, 10 if n <= 2048 else 3) for ph in (0, 1, 2, 3)]
    print(f"[{tag}] n={n} d={d} B={B}: assemble {t[0]:.3f} factor+inverse {t[1] - t[0]:.3f} alpha+kinv {t[2] - t[1]:.3f} "
          f"grad+finish {t[3] - t[2]:.3f} whole evaluation {t[3]:.3f} ms", flush=True)
    del ctx
