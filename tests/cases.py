"""Seeded input generators shared by the golden-vector script and the parity tests.

Every generator is a pure function of its arguments (numpy PCG64 streams are stable across
numpy versions), so tests/golden/ only has to store shapes, seeds and expected outputs.
Points are AoS float32 arrays, S = queries [m][k], R = references [n][k], as the reference's
harness passes them (/root/reference/sources/src/generator.h:32-50)."""
from __future__ import annotations

import zlib

import numpy as np


def uniform(seed: int, k: int, m: int, n: int):
    """Uniform [0,1) floats -- the distribution the reference is run on (generator.h:17-19)."""
    rng = np.random.default_rng(seed)
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    return S, R


def quantized(seed: int, k: int, m: int, n: int, bits: int = 4):
    """Coordinates on a 2^-bits grid: every distance is exact, exact ties are everywhere and
    the lowest-index rule (strict `>`, core.cu:50) decides the answer."""
    rng = np.random.default_rng(seed)
    q = float(1 << bits)
    S = (rng.integers(0, 1 << bits, (m, k)) / q).astype(np.float32)
    R = (rng.integers(0, 1 << bits, (n, k)) / q).astype(np.float32)
    return S, R


def duplicated(seed: int, k: int, m: int, n: int, period: int = 97):
    """Uniform references where reference j (j >= period) repeats reference j - period with
    probability 1/2, so the true nearest neighbour usually exists at several indices that lie
    in different tiles / CTAs / shards; the lowest one must win."""
    rng = np.random.default_rng(seed)
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    dup = rng.random(n) < 0.5
    for j in range(period, n):
        if dup[j]:
            R[j] = R[j - period]
    return S, R


def twins(seed: int, k: int, m: int, n: int, per_query: int = 6):
    """Rounding-level near-ties.  Each query q gets `per_query` "twin" references
    q -+ perm(d) at random indices, all built from ONE difference vector d (multiples of 2^-23,
    ~17 significant bits, so every d_j^2 is inexact in float32).  In exact arithmetic the twins
    are equidistant; in float32 the winner is decided by the last-bit rounding of the
    sequential, non-fused sum of core.cu:44-49.  An FMA contraction, a pairwise/re-ordered sum
    or a flush-to-zero changes which twin wins, so this family pins the arithmetic, not just
    the algorithm.  The remaining references are uniform noise (farther away)."""
    rng = np.random.default_rng(seed)
    S = (rng.integers(1 << 22, 1 << 23, (m, k)) * 2.0 ** -23).astype(np.float32)
    R = rng.random((n, k), dtype=np.float32)
    slots = rng.permutation(n)
    t = 0
    for i in range(m):
        d = (rng.integers(1 << 13, 1 << 17, k) * 2.0 ** -23).astype(np.float32) * np.float32(0.125)
        for _ in range(per_query):
            if t >= n:
                break
            sign = rng.integers(0, 2, k).astype(np.float32) * 2 - 1
            R[slots[t]] = (S[i] - (d[rng.permutation(k)] * sign)).astype(np.float32)
            t += 1
    return S, R


def specials(seed: int, k: int, m: int, n: int):
    """Uniform data with NaN / +-Inf / huge coordinates sprinkled in.  v0 never updates on a NaN
    distance (`min > NaN` is false) nor on +Inf (`Inf > Inf` is false), so an all-bad query
    answers index 0 (core.cu:39-40, 50)."""
    rng = np.random.default_rng(seed)
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    bad = np.array([np.nan, np.inf, -np.inf, 3.0e38, -3.0e38, 1.0e-40, -0.0], dtype=np.float32)
    for j in rng.integers(0, n, max(1, n // 7)):
        R[j, rng.integers(0, k)] = bad[rng.integers(0, len(bad))]
    for i in rng.integers(0, m, max(1, m // 5)):
        S[i, rng.integers(0, k)] = bad[rng.integers(0, len(bad))]
    if m > 1:
        S[m - 1, :] = np.nan  # a query whose every distance is NaN -> index 0
    return S, R


GENERATORS = {
    "uniform": uniform,
    "quantized": quantized,
    "duplicated": duplicated,
    "twins": twins,
    "specials": specials,
}


def make(kind: str, seed: int, k: int, m: int, n: int):
    S, R = GENERATORS[kind](seed, k, m, n)
    return np.ascontiguousarray(S, np.float32), np.ascontiguousarray(R, np.float32)


def checksum(*arrays) -> int:
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).view(np.uint8).tobytes(), c)
    return c


# (kind, seed, k, m, n): small enough that the CPU oracle needs well under a second each.
# Shapes cover: m or n = 1, n smaller than any tile, n not a multiple of 4 (16-byte staging
# granularity for k = 3), every k in 3..16, k*4 bytes not 16-byte aligned.
GOLDEN_CASES = (
    [("uniform", 1, 3, 1, 1), ("uniform", 2, 3, 1, 2), ("uniform", 3, 3, 2, 8),
     ("uniform", 4, 16, 1, 1), ("uniform", 5, 16, 5, 3), ("uniform", 6, 8, 8, 4099),
     ("uniform", 7, 3, 1000, 1031), ("uniform", 8, 16, 257, 5000), ("uniform", 9, 5, 33, 777)]
    + [("uniform", 100 + k, k, 67, 1501 + k) for k in range(3, 17)]
    + [("quantized", 200 + k, k, 129, 3001 + 3 * k) for k in (3, 4, 7, 8, 13, 16)]
    + [("duplicated", 300 + k, k, 200, 4000 + k) for k in (3, 8, 16)]
    + [("twins", 400 + k, k, 300, 2500 + k) for k in (3, 4, 5, 8, 11, 16)]
    + [("specials", 500 + k, k, 64, 900 + k) for k in (3, 8, 16)]
)
