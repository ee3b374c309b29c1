"""Full-evaluation time against n (B = 1 and B = 8): looks for cliffs off the tuned sizes."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402

d = 16
for n in (4096, 4160, 4352, 4608, 5120, 6144):
    rng = np.random.default_rng(1)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1)
    y = (y - y.min()) / (y - y.min()).mean() + 0.05
    th = np.array([[math.log(0.05), 0.0] + [math.log(1.5)] * d])
    ctx = h.Context(0, h.F64)
    ctx.set_data(x, y)
    out = []
    for B in (1, 8):
        ths = np.repeat(th, B, axis=0)
        ctx.bench_phase(ths, 3, 1)
        t = [ctx.bench_phase(ths, ph, 3) for ph in (0, 1, 2, 3)]
        out.append(f"B={B}: asm {t[0]:.2f} fac {t[1]-t[0]:.2f} kinv {t[2]-t[1]:.2f} grad {t[3]-t[2]:.2f} total {t[3]:.2f} ms "
                   f"({(n/1e3)**3 * B / t[3]:.1f} TF n^3)")
    t0 = time.perf_counter()
    for _ in range(3):
        m = ctx.model(th[0], want_alpha=False)
    tm = (time.perf_counter() - t0) / 3 * 1e3
    print(f"n={n}: " + " | ".join(out) + f" | model_create {tm:.2f} ms", flush=True)
    del m, ctx
