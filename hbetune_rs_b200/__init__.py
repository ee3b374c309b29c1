"""hbetune_rs_b200 — B200-native Gaussian-process surrogate hot path of hbetune (latk/hbetune.rs).

The product is ``libhbegp.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/hbegp.h``).
This package is the thin host-side mirror of the reference interface used by the tests and the
benchmark: ``gpr`` (``src/gpr``: kernel objects, ``FittedKernel``, ``predict``), ``estimator``
(``src/core/gpr.rs``: ``EstimatorGPR`` / ``SurrogateModelGPR``) and ``dist`` (restart / candidate
sharding over ``torch.distributed``).
"""
from ._lib import F32, F64, NOT_PD, OK, HbegpError, lib  # noqa: F401
from .gpr import (BoundedValue, BoundsError, ConstantKernel, Context, FittedKernel, Matern, Model,  # noqa: F401
                  MultiContext, MultiModel, Product, predict)

from .estimator import (LINEAR, LOGARITHMIC, EstimatorGPR, SummaryStatistics, SurrogateModelGPR, YNormalize,  # noqa: F401
                        estimate_amplitude, expected_improvement, find_best_candidate_by_ei,
                        find_best_individual_by_confidence_bound, predicted_fitness)

from .random import RNG  # noqa: F401

__version__ = lib.hbegp_version().decode()
