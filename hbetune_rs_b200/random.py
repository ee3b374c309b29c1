"""Host-side mirror of ``src/core/random.rs``: Xoshiro256** with the reference's seeding, forking and inclusive
uniform sampling, backed by the C++ implementation in ``libhbegp.so`` (``hbegp_rng_*``)."""
from __future__ import annotations

import ctypes as C

from ._lib import lib


class RNG:
    """``RNG::new_with_seed`` / ``fork_random_state`` / ``uniform(lo..=hi)`` (``src/core/random.rs:11-37``)."""

    def __init__(self, state):
        self._s = (C.c_ulonglong * 4)(*state)

    @classmethod
    def new_with_seed(cls, seed: int) -> "RNG":
        s = (C.c_ulonglong * 4)()
        lib.hbegp_rng_seed(seed, s)
        return cls(list(s))

    @property
    def state(self):
        return list(self._s)

    def fork_random_state(self) -> "RNG":
        child = (C.c_ulonglong * 4)()
        lib.hbegp_rng_fork(self._s, child)
        return RNG(list(child))

    def uniform_inclusive(self, lo: float, hi: float) -> float:
        return lib.hbegp_rng_uniform(self._s, lo, hi)
