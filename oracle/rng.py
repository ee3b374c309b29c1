"""Restatement of ``src/core/random.rs`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

The crates behind it are absent from /root/reference and restated from their published
algorithms: rand_xoshiro 0.4.0 ``Xoshiro256StarStar`` (``seed_from_u64`` = SplitMix64 expansion,
``from_rng`` = 32 bytes of little-endian ``next_u64`` output), rand 0.7.2 ``Uniform<f64>``
(``new_inclusive`` / ``sample``).  The stream itself is unpinned by the reference's tests
(SURVEY.md section 8c-4).
"""
from __future__ import annotations

import struct

MASK = (1 << 64) - 1


def _rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & MASK


def _f64_from_bits(bits: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def _bits_from_f64(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


class RNG:
    """``src/core/random.rs:11-52``."""

    def __init__(self, state):
        self.s = list(state)

    @classmethod
    def new_with_seed(cls, seed: int) -> "RNG":
        # rand_xoshiro: seed_from_u64 -> SplitMix64 fills the four state words
        state = []
        x = seed & MASK
        for _ in range(4):
            x = (x + 0x9E3779B97F4A7C15) & MASK
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
            state.append(z ^ (z >> 31))
        return cls(state)

    def next_u64(self) -> int:
        s = self.s
        result = (_rotl((s[1] * 5) & MASK, 7) * 9) & MASK
        t = (s[1] << 17) & MASK
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = _rotl(s[3], 45)
        return result

    def fork_random_state(self) -> "RNG":
        # random.rs:22-25: SeedableRng::from_rng fills the 32-byte seed with fill_bytes (LE u64 words)
        words = [self.next_u64() for _ in range(4)]
        if all(w == 0 for w in words):
            return RNG.new_with_seed(0)
        return RNG(words)

    def uniform_inclusive(self, low: float, high: float) -> float:
        """``rng.uniform(lo..=hi)`` (``gradmin.rs:23``): rand 0.7.2 UniformFloat::new_inclusive + sample."""
        assert low <= high
        max_rand = _f64_from_bits(((MASK >> 12)) | (1023 << 52)) - 1.0  # 1 - 2^-52
        scale = (high - low) / max_rand
        while scale * max_rand + low > high:
            scale = _f64_from_bits(_bits_from_f64(scale) - 1)
        value1_2 = _f64_from_bits((self.next_u64() >> 12) | (1023 << 52))
        return (value1_2 - 1.0) * scale + low
