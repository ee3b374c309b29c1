"""Phase times of one batched evaluation at the C3 shape (n = 1024, d = 8, B = 33) and at small batches of the
north-star shape; used to decide where the small-n regime loses its time."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, bench
import hbetune_rs_b200 as h

def phases(n, d, restarts, dtype="f64", reps=20):
    a = argparse.Namespace(n=n, d=d, restarts=restarts, m=8, dtype=dtype)
    _, x, y, lo, hi, th, _ = bench.workload(a)
    with h.Context(0, h.F64 if dtype == "f64" else h.F32) as ctx:
        ctx.set_data(x, y)
        t = [ctx.bench_phase(th, ph, reps) for ph in (0, 1, 2, 3)]
        t5 = ctx.bench_phase(th, 5, reps)
    return {"n": n, "B": len(th), "dtype": dtype, "assemble": t[0], "factor_inverse": t[1] - t[0], "alpha_kinv": t[2] - t[1],
            "grad_finish": t[3] - t[2], "eval": t[3], "kinv_gemm": t5}

for cfg in [(1024, 8, 32, "f64"), (1024, 8, 32, "f32"), (512, 8, 32, "f64"), (4096, 16, 8, "f64"), (4096, 16, 0, "f64"), (500, 8, 2, "f64")]:
    print(json.dumps(phases(*cfg, reps=20 if cfg[0] <= 1024 else 3)), flush=True)
