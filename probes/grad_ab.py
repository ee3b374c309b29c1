"""A/B of k_grad_contract builds (HBEGP_LIB): gradient phase time at the north-star and C3 shapes."""
import os, sys, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import hbetune_rs_b200 as h
for n, d, r, reps in ((4096, 16, 64, 3), (1024, 8, 32, 20)):
    a = argparse.Namespace(n=n, d=d, restarts=r, m=8, dtype="f64")
    _, x, y, lo, hi, th, _ = bench.workload(a)
    with h.Context(0, h.F64) as ctx:
        ctx.set_data(x, y)
        t2 = ctx.bench_phase(th, 2, reps); t3 = ctx.bench_phase(th, 3, reps)
        lml, grad, st = ctx.lml_grad_batch(th[:2])
    print(os.environ.get("HBEGP_LIB", "default").split("/")[-1], n, "grad+finish ms %.4f" % (t3 - t2), "eval %.3f" % t3, "grad[0][:3]", grad[0][:3], flush=True)
