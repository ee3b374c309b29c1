"""Probe: distribution of lockstep batch sizes and wall time per round over a whole north-star fit."""
import collections, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hbetune_rs_b200 as h
from bench import synth

n, d, restarts = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 16, int(sys.argv[3]) if len(sys.argv) > 3 else 64
x, y = synth(n, d)
lo_c = max(np.quantile(y, 0.1) ** 2 * n, 2e-5) / 2
hi_c = 2 * float((y ** 2).sum())
bv = h.BoundedValue
kernel = h.Product(h.ConstantKernel(bv(math.sqrt(lo_c * hi_c), lo_c, hi_c)), h.Matern(2.5, [bv(1.0, 1e-3, 1e3)] * d))
ctx = h.Context()
t0 = time.perf_counter()
fk = h.FittedKernel.new(ctx, kernel, x, y, h.RNG.new_with_seed(1), restarts, bv(1.0, 1e-2, 1e1))
print("fit seconds", time.perf_counter() - t0, "evals", fk.n_evals, "lml", fk.lml)
