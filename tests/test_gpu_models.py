"""Retained-model policy: the reference keeps the model of every generation
(src/core/minimize.rs:331, :407); only the newest few keep their n x n factor on the device, older ones refactorise on
their next variance prediction and answer with the same bits."""
import math

import numpy as np
import pytest

from tests.util import random_thetas, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("A", [np.float64, np.float32])
def test_old_models_are_evicted_and_rebuilt_on_demand(A):
    import hbetune_rs_b200 as h
    d = 3
    x, y = synth(700, d, A=A)
    xs = np.random.default_rng(4).random((300, d)).astype(A)
    thetas = random_thetas(6, d, seed=2, noise=(0.05, 0.5))
    with h.Context(0, h.F64 if A == np.float64 else h.F32) as ctx:
        ctx.set_resident_models(2)
        models, first = [], []
        for g in range(6):  # the history grows like the tuner's: 10 more rows per generation
            n = 640 + 10 * g
            ctx.set_data(x[:n], y[:n])
            models.append(ctx.model(thetas[g]))
            first.append(models[-1].predict(xs))
        st = ctx.model_stats()
        assert st["live"] == 6 and st["resident"] == 2 and st["evictions"] == 4 and st["rebuilds"] == 0
        # mean only: no factor needed, nothing is rebuilt
        mean_only, _ = models[0].predict(xs, want_variance=False)
        np.testing.assert_array_equal(mean_only, first[0][0])
        assert ctx.model_stats()["rebuilds"] == 0
        # an evicted model answers a variance request with the bits it gave before (the context now holds OTHER data)
        for g in (0, 3):
            mean, var = models[g].predict(xs)
            np.testing.assert_array_equal(mean, first[g][0])
            np.testing.assert_array_equal(var, first[g][1])
        st = ctx.model_stats()
        assert st["rebuilds"] == 2 and st["resident"] == 2
        # the two most recently used are resident now; the newest-created ones were evicted in turn and still work
        mean, var = models[5].predict(xs)
        np.testing.assert_array_equal(var, first[5][1])
        ctx.set_resident_models(0)  # unlimited: nothing is evicted any more
        extra = ctx.model(thetas[0])
        assert ctx.model_stats()["resident"] >= 3
        for m in models + [extra]:
            m.close()
