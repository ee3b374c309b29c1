"""A/B of per-launch priorities (csrc/launch.h): whole-evaluation time per shape, one process per setting because the
priority policy is read once per process.  Usage: python probes/prio_ab.py [SETTING ...] where a SETTING is a
comma-separated list of environment assignments, e.g. HBEGP_PRIO_CTAS=0,HBEGP_STREAMS=8 (PRIO_SHAPES=0,4 picks shapes;
inside a SETTING write the indices with '+': PRIO_SHAPES=0+4)."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [(1024, 8, 32, "f64"), (1024, 8, 32, "f32"), (512, 8, 32, "f64"), (2048, 16, 32, "f64"), (4096, 16, 8, "f64"),
          (4096, 16, 64, "f64"), (4096, 16, 64, "f32"), (4096, 16, 8, "f32"), (1024, 8, 8, "f64"), (500, 8, 2, "f64"), (2048, 16, 8, "f64")]

def child():
    import argparse, bench
    import hbetune_rs_b200 as h
    out = {}
    pick = os.environ.get("PRIO_SHAPES")
    shapes = [SHAPES[int(i)] for i in pick.split("+")] if pick else SHAPES
    for n, d, r, dt in shapes:
        a = argparse.Namespace(n=n, d=d, restarts=r, m=8, dtype=dt)
        _, x, y, lo, hi, th, _ = bench.workload(a)
        with h.Context(0, h.F64 if dt == "f64" else h.F32) as ctx:
            ctx.set_data(x, y)
            reps = 30 if n <= 1024 else (10 if n <= 2048 else 4)
            ctx.bench_phase(th, 3, 2)
            out[f"n{n}_B{len(th)}_{dt}"] = round(min(ctx.bench_phase(th, 3, reps) for _ in range(3)), 4)
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    if os.environ.get("PRIO_CHILD"):
        child()
    else:
        for setting in (sys.argv[1:] or ["HBEGP_PRIO_CTAS=0", "HBEGP_PRIO_CTAS=444"]):
            env = dict(os.environ, PRIO_CHILD="1")
            env.update(kv.split("=", 1) for kv in setting.split(","))
            r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
            print(f"{setting}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]}", flush=True)
