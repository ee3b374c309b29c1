"""Debug aid: predictive variance through the tcgen05 path vs the FFMA path for ragged sizes."""
import os, sys, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import synth, random_thetas
import hbetune_rs_b200 as h

def mkctx(tf32):
    os.environ["HBEGP_TF32"] = "1" if tf32 else "0"
    return h.Context(0, h.F32)

n, d, m = 320, 6, 1000
A = np.float32
x, y = synth(n, d, A=A)
xs = np.random.default_rng(2).random((m, d)).astype(A)
theta = random_thetas(1, d, seed=5, noise=(0.1, 0.5))[0]
c0 = mkctx(False); c0.set_data(x, y); m0 = c0.model(theta); mean0, var0 = m0.predict(xs, warn=False)
def rep(tag, var):
    dv = np.abs(var - var0); print(tag, "max dvar %.3e bad %d" % (dv.max(), (dv > 1e-3).sum()), flush=True)
ctx = mkctx(True); ctx.set_data(x, y)
mk = ctx.model(theta, want_kinv=True)
rep("1 tf32 ctx, kinv model, tf32 predict      ", mk.predict(xs, warn=False)[1])
c2 = mkctx(False)   # flips the process-wide policy: FFMA from here on
rep("2 same model, FFMA predict                 ", mk.predict(xs, warn=False)[1])
c3 = mkctx(True)
rep("3 same model, tf32 predict again           ", mk.predict(xs, warn=False)[1])
mn = ctx.model(theta, want_kinv=False)
rep("4 same ctx, new model without kinv, tf32   ", mn.predict(xs, warn=False)[1])
mk2 = ctx.model(theta, want_kinv=True)
rep("5 same ctx, second kinv model, tf32        ", mk2.predict(xs, warn=False)[1])
lml, grad, st = ctx.lml_grad_batch(theta[None])   # runs lauum (tf32) for the gradient
mn2 = ctx.model(theta, want_kinv=False)
rep("6 after a gradient evaluation, no-kinv model", mn2.predict(xs, warn=False)[1])
os.environ["HBEGP_GRAPHS"] = "0"
c4 = mkctx(True); c4.set_data(x, y)
mk4 = c4.model(theta, want_kinv=True)
rep("7 fresh ctx without graphs, kinv model     ", mk4.predict(xs, warn=False)[1])
