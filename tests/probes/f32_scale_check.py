"""f32 (--use-32) accuracy at scale: LML / gradient of the f32 CUDA path and of the f32 oracle (LAPACK spotrf/spotri),
both against the f64 oracle, at growing n (is the inverse-based recursion worse than LAPACK in single precision?)."""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import hbetune_rs_b200 as h  # noqa: E402
from tests.util import oracle_lml, synth  # noqa: E402

d = 16
for n in (512, 1024, 2048, 4096):
    x, y = synth(n, d, seed=1)
    for noise in (0.05, 0.01):
        th = np.array([math.log(noise), 0.0] + [math.log(1.5)] * d)
        ref = oracle_lml(th, x, y, A=np.float64)
        o32 = oracle_lml(th, x.astype(np.float32), y.astype(np.float32), A=np.float32)
        with h.Context(0, h.F32) as ctx:
            ctx.set_data(x.astype(np.float32), y.astype(np.float32))
            lml, grad, st = ctx.lml_grad_batch(th[None, :])
        gref = np.array(ref.lml_gradient)
        e_o = abs(o32.lml - ref.lml) / abs(ref.lml) if o32 is not None else float("nan")
        e_g = abs(lml[0] - ref.lml) / abs(ref.lml)
        ge_o = np.abs(np.array(o32.lml_gradient) - gref).max() / np.abs(gref).max() if o32 is not None else float("nan")
        ge_g = np.abs(grad[0] - gref).max() / np.abs(gref).max()
        print(f"n={n} noise={noise}: lml f64 {ref.lml:.6f} | rel err oracle-f32 {e_o:.2e} cuda-f32 {e_g:.2e} (status {st[0]}) | "
              f"grad rel err oracle-f32 {ge_o:.2e} cuda-f32 {ge_g:.2e}", flush=True)
