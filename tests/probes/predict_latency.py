"""Probe: latency of batch-of-1 predictions (the reference's callers predict one point at a time)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hbetune_rs_b200 as h
from tests.util import synth

for n, d in [(100, 2), (500, 8), (1024, 8), (4096, 16)]:
    x, y = synth(n, d)
    with h.Context() as ctx:
        ctx.set_data(x, y)
        model = ctx.model(np.array([math.log(0.1), 0.0] + [math.log(0.5)] * d))
        xs = np.random.default_rng(0).random((1, d))
        for want_var in (False, True):
            for _ in range(20):
                model.predict(xs, want_var)
            t0 = time.perf_counter()
            reps = 300
            for _ in range(reps):
                model.predict(xs, want_var)
            dt = (time.perf_counter() - t0) / reps
            print(f"n={n} d={d} m=1 variance={want_var}: {dt * 1e6:.1f} us per call")
        xs = np.random.default_rng(0).random((1000, d))
        t0 = time.perf_counter()
        for _ in range(20):
            model.predict(xs, True)
        print(f"n={n} d={d} m=1000 variance=True: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us per call")
