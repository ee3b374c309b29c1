"""Spread of the fitted hyper-parameters between the GPU fit and the oracle fit over many seeds (VERDICT r01 item 5,
SURVEY H4): the same problem family as BASELINE config 3 (n = 1024, d = 8, Matern-2.5 + noise), the reference's default
2 restarts (src/core/gpr.rs:219-236), one data seed and one RNG seed per row, both fits driven by the same bounded
L-BFGS (the library's; NLopt's is not available).

  python tests/probes/fit_spread.py oracle [first_seed] [n_seeds]   CPU only: writes tests/golden/fit_spread_oracle.json
  python tests/probes/fit_spread.py gpu                             GPU: refits every seed, prints / writes the table

The oracle side needs ~1-2 minutes per seed on 8 cores and no GPU, so its results are committed as a fixture and only
the GPU side runs on the B200 box (tests/test_gpu_fit.py::test_fit_spread_against_the_committed_oracle_fits).
"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N, D, RESTARTS = 1024, 8, 2
FIXTURE = os.path.join(ROOT, "tests", "golden", "fit_spread_oracle.json")
OTHER_OPTIMUM = 1e-2  # |d ln theta| beyond which two fits are different local optima, not the same one


def problem(seed):
    from tests.util import synth
    return synth(N, D, seed=100 + seed)


def kernels(mod, y):
    """EstimatorGPR::new defaults (gpr.rs:219-236) with the noise floor of SURVEY 8 d2 and estimate_amplitude (gpr.rs:429-450)."""
    bv = mod.BoundedValue
    y64 = np.asarray(y, dtype=np.float64)
    srt = np.sort(y64)
    lo = max(srt[int(math.floor((len(srt) - 1) * 0.1))] ** 2 * len(y64), 2e-5) / 2
    hi = 2 * float((y64 * y64).sum())
    c0 = math.exp((math.log(lo) + math.log(hi)) / 2)
    kernel = mod.Product(mod.ConstantKernel(bv(c0, lo, hi)), mod.Matern(2.5, [bv(1.0, 1e-3, 1e3)] * D))
    return kernel, bv(1.0, 1e-2, 1e1)


def run_oracle(first, count):
    from oracle import gpr as ogpr
    from oracle.rng import RNG
    from tests.util import lib_minimizer
    rows = {}
    if os.path.exists(FIXTURE):
        rows = json.load(open(FIXTURE))["rows"]
    for seed in range(first, first + count):
        x, y = problem(seed)
        kernel, noise = kernels(ogpr, y)
        t0 = time.time()
        fk = ogpr.fit_kernel(kernel, x, y, RNG.new_with_seed(seed), RESTARTS, noise, lib_minimizer())
        theta = [math.log(fk.noise.value)] + fk.kernel.theta()
        rows[str(seed)] = {"theta": theta, "lml": fk.lml, "n_evals": fk.n_evals}
        print(f"seed {seed}: lml {fk.lml:.12g}, {fk.n_evals} evaluations, {time.time() - t0:.0f} s", flush=True)
        json.dump({"_what": "oracle fits for tests/probes/fit_spread.py (n=1024, d=8, 2 restarts, library L-BFGS on the oracle objective)",
                   "n": N, "d": D, "restarts": RESTARTS, "rows": rows}, open(FIXTURE, "w"), indent=1)


def run_gpu(out_path=None):
    import hbetune_rs_b200 as h
    fixture = json.load(open(FIXTURE))
    table = []
    with h.Context() as ctx:
        for seed_s, ref in sorted(fixture["rows"].items(), key=lambda kv: int(kv[0])):
            seed = int(seed_s)
            x, y = problem(seed)
            kernel, noise = kernels(h, y)
            fk = h.FittedKernel.new(ctx, kernel, x, y, h.RNG.new_with_seed(seed), RESTARTS, noise)
            theta = np.array([math.log(fk.noise.value)] + fk.kernel.theta())
            tref = np.array(ref["theta"])
            # the LML the GPU assigns to the ORACLE's optimum: separates "different point" from "different value"
            lml_at_ref, _, _ = ctx.lml_grad_batch(tref[None], want_grad=False)
            table.append({"seed": seed, "lml_gpu": fk.lml, "lml_oracle": ref["lml"],
                          "d_lml_rel": abs(fk.lml - ref["lml"]) / abs(ref["lml"]),
                          "d_lml_same_theta_rel": abs(float(lml_at_ref[0]) - ref["lml"]) / abs(ref["lml"]),
                          "max_d_ln_theta": float(np.abs(theta - tref).max()), "theta_gpu": [float(v) for v in theta],
                          "evals_gpu": int(fk.n_evals), "evals_oracle": ref["n_evals"]})
            fk.model.close()
    # A fit whose theta is far from the oracle's ended in a DIFFERENT local optimum: one of its L-BFGS runs branched into
    # another basin.  That is a property of the optimiser on this objective, not of the evaluation (the same seed flips
    # when the rounding of the factorisation changes in the 13th digit); such seeds are listed
    # apart and the caller checks them against the oracle at the GPU's own theta.
    same = [r for r in table if r["max_d_ln_theta"] <= OTHER_OPTIMUM]
    other = [r for r in table if r["max_d_ln_theta"] > OTHER_OPTIMUM]
    worst = {k: max(r[k] for r in same) for k in ("d_lml_rel", "max_d_ln_theta")}
    worst["d_lml_same_theta_rel"] = max(r["d_lml_same_theta_rel"] for r in table)
    med = {k: float(np.median([r[k] for r in table])) for k in ("d_lml_rel", "d_lml_same_theta_rel", "max_d_ln_theta")}
    out = {"n": N, "d": D, "restarts": RESTARTS, "seeds": len(table), "same_optimum": len(same), "worst": worst, "median": med,
           "other_optimum": other, "rows": table}
    if out_path:
        json.dump(out, open(out_path, "w"), indent=1)
    return out


if __name__ == "__main__":
    if sys.argv[1] == "oracle":
        run_oracle(int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 20)
    else:
        res = run_gpu(os.path.join(ROOT, "gpurun_out", "r02_fit_spread.json"))
        print(json.dumps({k: res[k] for k in ("seeds", "same_optimum", "worst", "median", "other_optimum")}))
