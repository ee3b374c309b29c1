import sys, math
sys.path.insert(0, '.')
import numpy as np
import hbetune_rs_b200 as h
from oracle import gpr as ogpr
from oracle.rng import RNG
from tests.util import synth, lib_minimizer
A = np.float32
for n, d, restarts in [(40, 2, 2), (120, 3, 3)]:
    x, y = synth(n, d, A=A)
    def kernels(mod):
        bv = mod.BoundedValue
        return mod.Product(mod.ConstantKernel(bv(1.0, 1e-2, 1e2)), mod.Matern(2.5, [bv(1.0, 1e-2, 1e2)] * d)), bv(1.0, 1e-1, 1e1)
    ok, onoise = kernels(ogpr)
    ref = ogpr.fit_kernel(ok, x, y, RNG.new_with_seed(7), restarts, onoise, lib_minimizer(), A=A)
    gk, gnoise = kernels(h)
    with h.Context(0, h.F32) as ctx:
        fk = h.FittedKernel.new(ctx, gk, x, y, h.RNG.new_with_seed(7), restarts, gnoise)
        xs = np.random.default_rng(0).random((30, d)).astype(A)
        var = np.zeros(30, dtype=A); mean = h.predict(fk, xs, var)
    var_ref = np.zeros(30, dtype=A)
    mean_ref = ogpr.predict(ref.kernel, ref.alpha, xs, x, ref.k_inv, var_ref, A)
    th_g = np.array([math.log(fk.noise.value)] + fk.kernel.theta()); th_r = np.array([math.log(ref.noise.value)] + ref.kernel.theta())
    print(n, d, "lml", fk.lml, ref.lml, "rel", abs(fk.lml-ref.lml)/abs(ref.lml), "dtheta", np.abs(th_g-th_r).max(), "dmean", np.abs(mean-mean_ref).max(), "dvar", np.abs(var-var_ref).max(), "evals", fk.n_evals, ref.n_evals)
