"""Shared helpers for the parity tests: synthetic problems and oracle shims."""
import ctypes as C
import math

import numpy as np

from oracle import gpr as ogpr
from oracle.gpr import BoundedValue, ConstantKernel, Matern, Product


def synth(n, d, seed=1, A=np.float64):
    """X ~ U[0,1)^{n x d}; y = sum_k sin(2 pi x_k) + 0.1 N(0,1), normalised like the linear YNormalize
    (ynormalize.rs:168-173) so that mean(y) = 1.05."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1) + 0.1 * rng.standard_normal(n)
    y = y - y.min()
    y = y / (y.mean() if y.mean() > 0 else 1.0) + 0.05
    return x.astype(A), y.astype(A)


def oracle_kernel(theta, nu=2.5, lo=1e-9, hi=1e9):
    """Product<ConstantKernel, Matern> at theta[1:] (theta[0] is ln noise)."""
    c = math.exp(theta[1])
    ls = [math.exp(t) for t in theta[2:]]
    return Product(ConstantKernel(BoundedValue(c, min(lo, c), max(hi, c))),
                   Matern(nu, [BoundedValue(l, min(lo, l), max(hi, l)) for l in ls]))


def oracle_lml(theta, x, y, nu=2.5, A=np.float64):
    return ogpr.lml_with_gradient(oracle_kernel(theta, nu), A(math.exp(theta[0])), x, y, A)


def random_thetas(B, d, seed=3, noise=(1e-2, 1.0), c=(0.3, 3.0), ls=(0.2, 3.0)):
    rng = np.random.default_rng(seed)
    th = np.empty((B, d + 2))
    th[:, 0] = rng.uniform(math.log(noise[0]), math.log(noise[1]), B)
    th[:, 1] = rng.uniform(math.log(c[0]), math.log(c[1]), B)
    th[:, 2:] = rng.uniform(math.log(ls[0]), math.log(ls[1]), (B, d))
    return th


def lib_minimizer(maxeval=150):
    """minimize_by_gradient(objective, x0, bounds) backed by the host library's bounded L-BFGS
    (the same optimiser hbegp_fit_runs drives), for running the ORACLE objective through it."""
    from hbetune_rs_b200 import _lib

    def minimize(objective, x0, bounds):
        n = len(x0)
        lo = np.array([b[0] for b in bounds], dtype=np.float64)
        hi = np.array([b[1] for b in bounds], dtype=np.float64)
        x = np.array(x0, dtype=np.float64)

        def cb(xp, gp, _user):
            xv = np.array([xp[i] for i in range(n)])
            f, g = objective(xv)
            for i in range(n):
                gp[i] = g[i]
            return f

        fn = _lib.OBJECTIVE_FN(cb)
        fout = C.c_double()
        rc = _lib.lib.hbegp_minimize_by_gradient(fn, None, n, x.ctypes.data_as(C.c_void_p),
                                                 lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                                 maxeval, C.byref(fout))
        assert rc >= 0
        return x, fout.value

    return minimize
