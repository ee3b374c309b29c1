"""CPU probe: evaluations used by the library's bounded L-BFGS (hbegp_minimize_by_gradient) against scipy's L-BFGS-B on the
oracle's negative LML from the same random starts (the reference uses NLopt's L-BFGS, which is not available here)."""
import math, sys, time
import numpy as np
sys.path.insert(0, ".")
from scipy.optimize import minimize
from tests.util import synth, oracle_lml, lib_minimizer

def make_obj(x, y, lo, hi):
    cnt = [0]
    def obj(th):
        cnt[0] += 1
        thc = np.clip(th, lo, hi)
        r = oracle_lml(thc, x, y)
        if r is None:
            return float("inf"), np.zeros_like(th)
        return -r.lml, -np.array(r.lml_gradient)
    return obj, cnt

rng = np.random.default_rng(3)
tot_ours = tot_sp = 0
for (n, d) in ((60, 2), (120, 4), (200, 8)):
    x, y = synth(n, d, seed=n)
    lo = np.log(np.array([1e-2, 1e-2] + [1e-3] * d)); hi = np.log(np.array([1e1, 1e3] + [1e3] * d))
    for trial in range(6):
        th0 = rng.uniform(lo, hi)
        obj, cnt = make_obj(x, y, lo, hi)
        mz = lib_minimizer(150)
        xo, fo = mz(obj, th0.copy(), list(zip(lo, hi)))
        c_ours = cnt[0]
        obj2, cnt2 = make_obj(x, y, lo, hi)
        res = minimize(lambda t: obj2(t), th0.copy(), jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)), options={"maxfun": 150, "ftol": 1e-12, "gtol": 1e-8})
        print(f"n={n} d={d} trial {trial}: ours {c_ours:3d} evals f={fo:.6f} | scipy {cnt2[0]:3d} evals f={res.fun:.6f}")
        tot_ours += c_ours; tot_sp += cnt2[0]
print("total evals: ours", tot_ours, "scipy", tot_sp)
