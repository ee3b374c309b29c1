"""One batched evaluation (after one warm-up) for an ncu launch list: python probes/one_eval.py N D B"""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402

n, d, B = (int(a) for a in sys.argv[1:4])
rng = np.random.default_rng(1)
x = rng.random((n, d))
y = np.sin(2 * np.pi * x).sum(axis=1)
y = (y - y.min()) / (y - y.min()).mean() + 0.05
th = np.repeat(np.array([[math.log(0.05), 0.0] + [math.log(1.5)] * d]), B, axis=0)
ctx = h.Context(0, h.F64)
ctx.set_data(x, y)
for _ in range(2):
    lml, grad, st = ctx.lml_grad_batch(th)
print(lml[0], st[0])
