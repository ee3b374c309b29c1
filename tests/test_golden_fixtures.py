"""Committed fixtures (tests/golden/): the reference's own known-answer vectors against the oracle (CPU), and the
oracle's committed outputs against both the oracle (drift guard, CPU) and the CUDA path (GPU)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import gpr as ogpr
from oracle.gpr import BoundedValue, ConstantKernel, Matern, Product
from tests.util import oracle_kernel, oracle_lml, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = json.load(open(os.path.join(GOLDEN, "reference_kernel_goldens.json")))
ORACLE = json.load(open(os.path.join(GOLDEN, "oracle_small_cases.json")))["cases"]


def _bv(v):
    return BoundedValue(v, 0.05, 20.0)


@pytest.mark.parametrize("name,nu", [("matern_nu_1_5", 1.5), ("matern_nu_2_5", 2.5)])
def test_reference_matern_goldens(name, nu):
    g = REF[name]
    k, grad = Matern(nu, [_bv(l) for l in g["length_scale"]]).theta_grad(np.array(g["x"]))
    np.testing.assert_allclose(k, np.array(g["kernel"]), atol=6e-9)
    np.testing.assert_allclose(grad, np.array(g["gradient"]), atol=6e-9)


def test_reference_product_golden():
    g = REF["product_constant2_matern_2_5"]
    kern = Product(ConstantKernel(BoundedValue(g["constant"], 1.0, 5.0)), Matern(2.5, [_bv(l) for l in g["length_scale"]]))
    k, grad = kern.theta_grad(np.array(g["x"]))
    np.testing.assert_allclose(k, np.array(g["kernel"]), rtol=6e-9, atol=1e-12)
    np.testing.assert_allclose(grad, np.array(g["gradient"]), rtol=6e-9, atol=1e-12)


@pytest.mark.parametrize("case", ORACLE, ids=lambda c: f"n{c['n']}_d{c['d']}_nu{c['nu']}")
def test_oracle_reproduces_its_committed_fixtures(case):
    x, y = synth(case["n"], case["d"], seed=case["seed"])
    theta = np.array(case["theta"])
    res = oracle_lml(theta, x, y, nu=case["nu"])
    assert res.lml == pytest.approx(case["lml"], rel=1e-12)
    np.testing.assert_allclose(res.lml_gradient, case["lml_gradient"], rtol=1e-9, atol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ORACLE, ids=lambda c: f"n{c['n']}_d{c['d']}_nu{c['nu']}")
def test_cuda_path_matches_committed_oracle_fixtures(case):
    import hbetune_rs_b200 as h
    x, y = synth(case["n"], case["d"], seed=case["seed"])
    theta = np.array(case["theta"])
    with h.Context() as ctx:
        ctx.set_data(x, y)
        lml, grad, status = ctx.lml_grad_batch(theta[None], nu=case["nu"])
        model = ctx.model(theta, nu=case["nu"])
        mean, var = model.predict(np.array(case["xs"]))
    assert status[0] == 0
    assert abs(lml[0] - case["lml"]) <= 1e-9 * abs(case["lml"])
    g = np.array(case["lml_gradient"])
    np.testing.assert_allclose(grad[0], g, rtol=1e-8, atol=1e-9 * np.abs(g).max())
    np.testing.assert_allclose(model.alpha[:5], case["alpha_head"], rtol=0, atol=1e-8 * np.abs(model.alpha).max())
    np.testing.assert_allclose(mean, case["mean"], rtol=0, atol=1e-9 * max(1.0, np.abs(case["mean"]).max()))
    np.testing.assert_allclose(var, case["var"], rtol=0, atol=1e-9 * (math.exp(theta[1]) + 1e-5))
