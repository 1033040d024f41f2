"""Synthetic plaintexts / ciphertexts for benchmarks and tests (numpy, host side).

Key generation and encryption are NOT on the accelerated path (they stay with the reference CPU
build); these few lines exist so that ``bench.py`` can produce inputs of the right shape and check
that what comes back decrypts, without touching ``oracle/``.  Conventions follow SURVEY.md 9:
``b = m + <a, s> + e`` (tlwe.c:106-115), ``phase = b - <a, s>`` (tlwe.c:135-141).
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64


def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """count uint64 words of splitmix64(seed) (the seeded host PRNG of SURVEY.md 8(d))."""
    idx = (np.arange(1, count + 1, dtype=U64) * U64(0x9E3779B97F4A7C15)) + U64(seed)
    z = idx
    z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
    return z ^ (z >> U64(31))


def binary_key(n: int, seed: int) -> np.ndarray:
    return (splitmix64_stream(seed, n) >> U64(63)).astype(U64)


def gaussian_torus(sigma: float, count: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.rint(rng.normal(0.0, sigma, count) * 2.0 ** 64).astype(np.int64).view(U64)


def encode(m, torus_base: int) -> np.ndarray:
    """int2torus(m, log2(2*torus_base)) (misc.c:25-28): message m in [0, torus_base) -> m / (2*torus_base)."""
    log = int(np.log2(2 * torus_base))
    return np.asarray(m, U64) << U64(64 - log)


def tlwe_encrypt(mu: np.ndarray, key: np.ndarray, sigma: float, seed: int) -> np.ndarray:
    """[count, n+1] TLWE samples of the torus values mu under the binary key."""
    count, n = mu.shape[0], key.shape[0]
    a = splitmix64_stream(seed, count * n).reshape(count, n)
    with np.errstate(over="ignore"):
        b = (a * key[None, :]).sum(axis=1, dtype=U64) + np.asarray(mu, U64) + gaussian_torus(sigma, count, seed + 1)
    return np.concatenate([a, b[:, None]], axis=1)


def tlwe_phase(ct: np.ndarray, key: np.ndarray) -> np.ndarray:
    n = key.shape[0]
    with np.errstate(over="ignore"):
        return ct[..., n] - (ct[..., :n] * key).sum(axis=-1, dtype=U64)


def test_vector(lut_vals: np.ndarray, N: int, k: int) -> np.ndarray:
    """trlwe_torus_packing (trlwe.c:662-667): trivial TRLWE whose b holds each LUT value N/size times."""
    size = lut_vals.shape[0]
    tv = np.zeros((k + 1, N), U64)
    tv[k] = np.repeat(np.asarray(lut_vals, U64), N // size)
    return tv


def torus_distance(a, b) -> np.ndarray:
    return np.abs((np.asarray(a, U64) - np.asarray(b, U64)).view(np.int64))
