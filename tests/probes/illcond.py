import sys, math
sys.path.insert(0, '.')
import numpy as np
import hbetune_rs_b200 as h
from tests.util import synth, oracle_lml, oracle_kernel
from oracle import gpr as ogpr
for n, d, noise, ls in [(300, 3, 1e-2, 0.5), (300, 3, 1e-4, 0.5), (300, 3, 1e-5, 1.0), (300, 3, 1e-5, 3.0), (1000, 2, 1e-5, 1.0)]:
    x, y = synth(n, d)
    theta = np.array([math.log(noise), 0.0] + [math.log(ls)] * d)
    ref = oracle_lml(theta, x, y)
    K = oracle_kernel(theta).kernel(x, x) + noise * np.eye(n)
    cond = np.linalg.cond(K)
    with h.Context() as ctx:
        ctx.set_data(x, y)
        lml, grad, st = ctx.lml_grad_batch(theta[None])
        model = ctx.model(theta)
        xs = np.random.default_rng(3).random((50, d))
        mean, var = model.predict(xs)
    vref = np.zeros(50)
    mref = ogpr.predict(oracle_kernel(theta), ref.alpha, xs, x, ref.factorization.invc(), vref)
    g = np.array(ref.lml_gradient)
    print(f"n={n} noise={noise:g} ls={ls}: cond={cond:.2e} rel dLML={abs(lml[0]-ref.lml)/abs(ref.lml):.2e} "
          f"rel dgrad={np.abs(grad[0]-g).max()/np.abs(g).max():.2e} dmean={np.abs(mean-mref).max()/max(1,np.abs(mref).max()):.2e} "
          f"dvar={np.abs(var-vref).max():.2e} status={st[0]}")
