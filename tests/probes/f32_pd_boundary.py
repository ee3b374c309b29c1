"""Where does single precision stop factoring K on the north-star problem, and how wrong is its LML before that?
Evaluates the f32 CUDA path with the tcgen05 3xTF32 GEMMs, with the FFMA GEMMs (HBEGP_TF32=0) and the f64 path on a noise
sweep through the two f32-fitted optima (profiles/r02_f32_optimum*.json)."""
import json, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import synth
import hbetune_rs_b200 as h

th_ffma = np.array(json.load(open(os.path.join(ROOT, "profiles", "r02_f32_optimum.json")))["fit_f32"]["theta"])
th_tf32 = np.array(json.load(open(os.path.join(ROOT, "profiles", "r02_f32_optimum_tf32.json")))["fit_f32"]["theta"])
n, d = 4096, 16
pts, names = [], []
for name, base in (("ffma_opt", th_ffma), ("tf32_opt", th_tf32)):
    for nz in (0.01, 0.02, 0.03, 0.048, 0.08, 0.12, 0.2, 0.27, 0.5):
        t = base.copy(); t[0] = math.log(nz); pts.append(t); names.append(f"{name} noise={nz}")
pts = np.array(pts)
res = {}
for tag, dtype, tf in (("f64", h.F64, "1"), ("f32_tf32", h.F32, "1"), ("f32_ffma", h.F32, "0")):
    os.environ["HBEGP_TF32"] = tf
    A = np.float64 if dtype == h.F64 else np.float32
    x, y = synth(n, d, A=A)
    with h.Context(0, dtype) as ctx:
        ctx.set_data(x, y)
        lml, grad, st = ctx.lml_grad_batch(pts)
    res[tag] = (lml, st, grad)
print("%-28s %14s %14s %14s" % ("point", "f64", "f32 tf32", "f32 ffma"))
for i, nm in enumerate(names):
    f = lambda t: ("%14.3f" % res[t][0][i]) if res[t][1][i] == 0 else "        NOT_PD"
    print("%-28s %s %s %s" % (nm, f("f64"), f("f32_tf32"), f("f32_ffma")))
g64, gt, gf = res["f64"][2], res["f32_tf32"][2], res["f32_ffma"][2]
for i, nm in enumerate(names):
    if res["f32_tf32"][1][i] == 0 and res["f32_ffma"][1][i] == 0:
        s = np.abs(g64[i]).max()
        print("%-28s grad err/max: tf32 %.2e ffma %.2e" % (nm, np.abs(gt[i] - g64[i]).max() / s, np.abs(gf[i] - g64[i]).max() / s))
