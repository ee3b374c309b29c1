"""Which GEMM class of the recursion costs the tcgen05 path its positive-definiteness margin?  (HBEGP_TF32_MASK bits:
0 panel solve, 1 T = L21 W11, 2 trailing update, 3 W21 = -W22 T, 4 K^-1 = W^T W)"""
import json, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import synth
import hbetune_rs_b200 as h

th = np.array(json.load(open(os.path.join(ROOT, "profiles", "r02_f32_optimum.json")))["fit_f32"]["theta"])
n, d = 4096, 16
noises = (0.048, 0.08, 0.12, 0.2, 0.27)
pts = []
for nz in noises:
    t = th.copy(); t[0] = math.log(nz); pts.append(t)
pts = np.array(pts)
x64, y64 = synth(n, d)
with h.Context(0, h.F64) as ctx:
    ctx.set_data(x64, y64)
    ref, _, _ = ctx.lml_grad_batch(pts)
x, y = synth(n, d, A=np.float32)
def run(tag, env):
    for k in ("HBEGP_TF32", "HBEGP_TF32_MASK", "HBEGP_TF32_MIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with h.Context(0, h.F32) as ctx:
        ctx.set_data(x, y)
        lml, grad, st = ctx.lml_grad_batch(pts)
    print("%-34s" % tag, " ".join(("%9.2f" % (lml[i] - ref[i])) if st[i] == 0 else "   NOT_PD" for i in range(len(noises))), flush=True)
print("%-34s" % "lml error vs f64 at noise =", " ".join("%9g" % z for z in noises))
run("ffma", {"HBEGP_TF32": "0"})
run("tf32 all", {})
for bit, name in enumerate(("panel L21=A21 W11^T", "T = L21 W11", "trailing A22 -= L21 L21^T", "W21 = -W22 T", "K^-1 = W^T W")):
    run("tf32 only " + name, {"HBEGP_TF32_MASK": str(1 << bit)})
run("tf32 all but trailing", {"HBEGP_TF32_MASK": str(0x3f & ~4)})
run("tf32 all, min extent 1024", {"HBEGP_TF32_MIN": "1024"})
