"""Multi-GPU paths on a box with at least two GPUs; skipped on one.  The exchange is inside libhbegp.so
(NCCL); no torch.distributed is involved here."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def test_cpp_multi_gpu_through_the_c_abi_only():
    """tests/cpp/multi_gpu_test.cpp: single-process handle and communicator mode, fit and predict `memcmp`-equal to 1 GPU."""
    exe = os.path.join(ROOT, "tests", "cpp", "multi_gpu_test")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/multi_gpu_test is not built (make -C hbetune_rs_b200/csrc)")
    res = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=600)
    if res.returncode == 77:
        pytest.skip(res.stdout.strip().splitlines()[-1])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "all ok" in res.stdout


def test_estimator_over_two_gpus_equals_one_gpu():
    """EstimatorGPR (src/core/gpr.rs:215-400) on a MultiContext: same fitted model and predictions as on one GPU, bit for bit."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import hbetune_rs_b200 as h
    n, d = 700, 4
    x, y = synth(n, d, seed=5)
    y = y * 30.0 + 7.0
    xs = np.random.default_rng(3).random((5000, d))
    single = h.EstimatorGPR(d).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(5)
    m1 = single.estimate(x, y, None, h.RNG.new_with_seed(11))
    with h.MultiContext(2) as mctx:
        multi = h.EstimatorGPR(d, ctx=mctx).with_noise_bounds(1e-2, 1e1).n_restarts_optimizer(5)
        m2 = multi.estimate(x, y, None, h.RNG.new_with_seed(11))
        assert m2.lml == m1.lml and m2.kernel.theta() == m1.kernel.theta() and m2.noise.value == m1.noise.value
        np.testing.assert_array_equal(m2.alpha, m1.alpha)
        np.testing.assert_array_equal(m2.predict_mean_a(xs), m1.predict_mean_a(xs))
        mean1, ei1 = m1.predict_mean_ei_a(xs, float(y.min()))
        mean2, ei2 = m2.predict_mean_ei_a(xs, float(y.min()))
        np.testing.assert_array_equal(mean2, mean1)
        np.testing.assert_array_equal(ei2, ei1)
        # the batched acquisition epilogues run on GPU 0's replica
        assert h.find_best_candidate_by_ei(xs, m2, float(y.min())) == h.find_best_candidate_by_ei(xs, m1, float(y.min()))
        ext2 = multi.extend(np.vstack([x, xs[:7]]), np.concatenate([y, y[:7]]), m2)
        ext1 = single.extend(np.vstack([x, xs[:7]]), np.concatenate([y, y[:7]]), m1)
        assert abs(ext2.lml - ext1.lml) <= 1e-9 * abs(ext1.lml)  # one GPU appends to the prior factor, the handle refactorises
        m2.fitted.model.close()
        ext2.fitted.model.close()
