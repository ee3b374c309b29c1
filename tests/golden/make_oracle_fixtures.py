#!/usr/bin/env python
"""Generates tests/golden/oracle_small_cases.json: LML, gradient, alpha and predictive mean / variance of the
CPU oracle (oracle/gpr.py) on small seeded problems.  The reference itself is Rust and cannot be built or
imported here, so these are ORACLE outputs (pinned by the reference's own golden vectors, see
reference_kernel_goldens.json), committed so that the GPU parity tests also compare against fixed numbers.

Run from the repository root:  python tests/golden/make_oracle_fixtures.py
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gpr as ogpr  # noqa: E402
from tests.util import oracle_kernel, oracle_lml, random_thetas, synth  # noqa: E402

cases = []
for (n, d, m, nu, seed) in [(12, 1, 5, 2.5, 1), (40, 2, 7, 2.5, 2), (97, 3, 9, 2.5, 3), (130, 5, 6, 1.5, 4), (70, 2, 6, 0.5, 5)]:
    x, y = synth(n, d, seed=seed)
    theta = random_thetas(1, d, seed=100 + seed, noise=(3e-2, 0.3))[0]
    res = oracle_lml(theta, x, y, nu=nu)
    xs = np.random.default_rng(50 + seed).random((m, d))
    var = np.zeros(m)
    mean = ogpr.predict(oracle_kernel(theta, nu), res.alpha, xs, x, res.factorization.invc(), var)
    cases.append({
        "n": n, "d": d, "nu": nu, "seed": seed, "theta": theta.tolist(), "xs": xs.tolist(),
        "lml": res.lml, "lml_gradient": list(res.lml_gradient), "alpha_head": res.alpha[:5].tolist(),
        "mean": mean.tolist(), "var": var.tolist(),
    })
out = os.path.join(ROOT, "tests", "golden", "oracle_small_cases.json")
json.dump({"generator": "tests/golden/make_oracle_fixtures.py", "data": "tests.util.synth(n, d, seed)", "cases": cases},
          open(out, "w"), indent=1)
print("wrote", out, len(cases), "cases")
