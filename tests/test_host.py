"""CPU-side tests: the C-ABI library loads and exports every declared symbol, and the host logic
(bounded L-BFGS, RNG restatement, kernel parameter objects) behaves like the reference's."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from oracle.rng import RNG
from tests.util import lib_minimizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from hbetune_rs_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hbegp.h")).read()
    declared = set(re.findall(r"\b(hbegp_[a-z_0-9]+)\s*\(", header))
    declared -= {"hbegp_objective_fn"}
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert getattr(_lib.lib, name) is not None
    assert _lib.lib.hbegp_version().startswith(b"hbegp")


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hbetune_rs_b200 as h
    with pytest.raises(h.HbegpError) as e:
        h.Context()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_minimize_by_gradient_slanted_plane():
    # src/util/gradmin.rs:62-102
    minimize = lib_minimizer()
    x, f = minimize(lambda x: (float(x.sum()), np.ones(2)), [0.0, 0.0], [(-2.0, 2.0)] * 2)
    assert list(x) == [-2.0, -2.0] and f == -4.0


def test_minimize_by_gradient_rosenbrock_in_box():
    minimize = lib_minimizer(maxeval=400)

    def rosen(x):
        f = 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2
        g = np.array([-400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200 * (x[1] - x[0] ** 2)])
        return float(f), g

    x, f = minimize(rosen, [-1.2, 1.0], [(-2.0, 2.0)] * 2)
    assert f < 1e-10 and np.allclose(x, [1.0, 1.0], atol=1e-4)
    x, f = minimize(rosen, [-1.2, 1.0], [(-2.0, 0.5)] * 2)  # optimum on the boundary
    assert abs(x[0] - 0.5) < 1e-6 and abs(x[1] - 0.25) < 1e-5


def test_minimize_handles_infinite_objective():
    minimize = lib_minimizer()

    def f(x):
        if x[0] > 1.0:
            return math.inf, np.zeros(1)
        return float((x[0] - 3.0) ** 2), np.array([2 * (x[0] - 3.0)])

    x, fx = minimize(f, [0.0], [(-5.0, 5.0)])
    assert 0.9 < x[0] <= 1.0


def test_rng_matches_oracle_restatement():
    from hbetune_rs_b200 import _lib
    st = (C.c_ulonglong * 4)()
    _lib.lib.hbegp_rng_seed(17176, st)
    ref = RNG.new_with_seed(17176)
    assert list(st) == ref.s
    child = (C.c_ulonglong * 4)()
    _lib.lib.hbegp_rng_fork(st, child)
    rchild = ref.fork_random_state()
    assert list(child) == rchild.s and list(st) == ref.s
    for lo, hi in [(-2.0, 2.0), (math.log(1e-5), math.log(1e5)), (0.0, 0.0)]:
        assert _lib.lib.hbegp_rng_uniform(child, lo, hi) == rchild.uniform_inclusive(lo, hi)


def test_xoshiro_known_answer():
    # Xoshiro256** reference output for state {1, 2, 3, 4} (public test vector of the algorithm)
    r = RNG([1, 2, 3, 4])
    assert [r.next_u64() for _ in range(3)] == [11520, 0, 1509978240]


def test_kernel_parameter_objects_mirror_the_reference():
    import hbetune_rs_b200 as h
    bv = h.BoundedValue
    k = h.Product(h.ConstantKernel(bv(2.0, 1.0, 5.0)), h.Matern(2.5, [bv(1.0, 0.05, 20.0)] * 2))
    assert k.n_params() == 3 and k.theta() == [math.log(2.0), 0.0, 0.0]
    assert k.bounds()[0] == (0.0, math.log(5.0))
    k2 = k.with_clamped_theta([math.log(9.0), math.log(0.01), 0.5])
    assert k2.k1.constant.value == 5.0 and k2.k2.length_scale[0].value == 0.05
    with pytest.raises(h.BoundsError):
        k.with_theta([math.log(9.0), 0.0, 0.0])


def test_python_rng_mirror_matches_oracle_stream():
    import hbetune_rs_b200 as h
    a, b = h.RNG.new_with_seed(123), RNG.new_with_seed(123)
    assert a.state == b.s
    fa, fb = a.fork_random_state(), b.fork_random_state()
    assert fa.state == fb.s and a.state == b.s
    assert [fa.uniform_inclusive(-11.5, 11.5) for _ in range(8)] == [fb.uniform_inclusive(-11.5, 11.5) for _ in range(8)]


def test_gpu_arm_of_bench_does_not_import_the_oracle():
    """Only the cpu_baseline / --impl reference legs may touch oracle/ (via tests.util, lazily)."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    top_level = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    names = [getattr(n, "module", None) or n.names[0].name for n in top_level]
    assert not any(str(m).startswith(("oracle", "tests")) for m in names), names
    for sub in ("hbetune_rs_b200",):
        for fn in os.listdir(os.path.join(ROOT, sub)):
            if fn.endswith(".py"):
                text = open(os.path.join(ROOT, sub, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn


def test_c_abi_rejects_bad_arguments_without_crashing():
    """Every entry point returns a status (< 0 with a message) instead of aborting."""
    from hbetune_rs_b200 import _lib
    L = _lib.lib
    null = None
    assert L.hbegp_set_data(null, 10, 2, null, null) == _lib.ERR_INVALID
    assert b"null context" in L.hbegp_last_error()
    assert L.hbegp_lml_grad_batch(null, 2.5, 1, null, null, null, null, null, null) == _lib.ERR_INVALID
    assert L.hbegp_fit_runs(null, 2.5, 1, null, null, null, 150, None, null) == _lib.ERR_INVALID
    assert L.hbegp_predict(null, 1, null, null, null, None) == _lib.ERR_INVALID
    assert L.hbegp_predict_device(null, 1, null, null, null, null) == _lib.ERR_INVALID
    assert L.hbegp_predict_mean_ei(null, None, 1, null, 0.0, null, null, None, None) == _lib.ERR_INVALID
    assert L.hbegp_ctx_destroy(null) == 0 and L.hbegp_model_destroy(null) == 0  # destroying nothing is fine
    assert L.hbegp_ctx_launch_count(null) == 0
    h = C.c_void_p()
    assert L.hbegp_ctx_create(0, 7, None, C.byref(h)) == _lib.ERR_INVALID  # bad dtype, checked before CUDA
    assert L.hbegp_ctx_create(0, 0, None, None) == _lib.ERR_INVALID
    x = (C.c_double * 2)(0.0, 0.0)
    assert L.hbegp_minimize_by_gradient(_lib.OBJECTIVE_FN(lambda a, b, c: 0.0), None, 0, x, x, x, 10, None) == _lib.ERR_INVALID
    yn = _lib.YNorm()
    y = np.ones(3)
    assert L.hbegp_ynorm_fit(0, 5, 3, y.ctypes.data_as(C.c_void_p), None, y.ctypes.data_as(C.c_void_p), C.byref(yn)) == _lib.ERR_INVALID
    assert L.hbegp_estimate_amplitude(0, 0, None, None, None) == _lib.ERR_INVALID
    res = (_lib.RunResult * 1)()
    res[0].status = 1
    assert L.hbegp_pick_best_run(1, res) == -1 and L.hbegp_pick_best_run(0, res) == -1


def test_minimizer_respects_maxeval_and_bounds():
    from hbetune_rs_b200 import _lib
    calls = []

    def f(xp, gp, _):
        calls.append((xp[0], xp[1]))
        gp[0], gp[1] = 2 * (xp[0] - 10.0), 2 * (xp[1] + 10.0)
        return (xp[0] - 10.0) ** 2 + (xp[1] + 10.0) ** 2

    x = np.array([0.0, 0.0])
    lo, hi = np.array([-1.0, -2.0]), np.array([3.0, 2.0])
    fout = C.c_double()
    n = _lib.lib.hbegp_minimize_by_gradient(_lib.OBJECTIVE_FN(f), None, 2, x.ctypes.data_as(C.c_void_p),
                                            lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), 7, C.byref(fout))
    assert 1 <= n <= 7 and len(calls) == n
    assert all(lo[0] <= a <= hi[0] and lo[1] <= b <= hi[1] for a, b in calls)  # never evaluated outside the box
    n = _lib.lib.hbegp_minimize_by_gradient(_lib.OBJECTIVE_FN(f), None, 2, x.ctypes.data_as(C.c_void_p),
                                            lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), 150, C.byref(fout))
    assert list(x) == [3.0, -2.0] and fout.value == 49.0 + 64.0  # the constrained optimum is the corner
