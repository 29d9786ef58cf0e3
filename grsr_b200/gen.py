"""Seeded synthetic genome generators for the workloads BASELINE.json names.

Nothing here is on the compute path; tests and bench.py use it to build inputs.
Genomes are signed permutations of 1..n stored as int32 numpy rows, the layout GRIMM keeps in
``struct genome_struct.genes`` (reference G/mcstructs.h:66-71).
"""
from __future__ import annotations

import random

import numpy as np


def random_signed_perm(n: int, rng: random.Random) -> list[int]:
    """Uniform permutation with i.i.d. signs: shuffle first, then one sign draw per element
    (the generator SURVEY.md section 8d fixes for configs 2, 3 and 5)."""
    p = list(range(1, n + 1))
    rng.shuffle(p)
    return [x if rng.random() < 0.5 else -x for x in p]


def random_genomes(G: int, n: int, seed: int) -> np.ndarray:
    """Genome 1 = identity, genomes 2..G random signed permutations (configs 2 and 3)."""
    rng = random.Random(seed)
    rows = [list(range(1, n + 1))]
    for _ in range(G - 1):
        rows.append(random_signed_perm(n, rng))
    return np.asarray(rows, dtype=np.int32)


def k_reversal_perm(n: int, k: int, rng: random.Random) -> list[int]:
    """Identity after k random signed reversals (distance <= k; many adjacencies)."""
    p = list(range(1, n + 1))
    for _ in range(k):
        i = rng.randrange(n)
        j = rng.randrange(i, n)
        p[i:j + 1] = [-x for x in reversed(p[i:j + 1])]
    return p


# ---- hurdle / fortress grammar -------------------------------------------
# A "shape" is a signed permutation of 1..len (a contiguous value range).
# Concatenating shapes with running offsets keeps their breakpoint-graph components apart.

FORTRESS_UNIT = [0, 2, 4, 3, 5, 1, 6]  # x + these offsets; three units in a row give f = 1


def random_shape(rng: random.Random, length: int, depth: int = 0) -> list[int]:
    """Random structure of exactly ``length`` elements, as 1-based signed values 1..length:
    hurdle units, fortress-unit groups, nested wraps and random-signed filler in random order."""
    out: list[int] = []
    while len(out) < length:
        rest = length - len(out)
        off = len(out)
        r = rng.random()
        if r < 0.30 and rest >= 3:
            piece = [1, 3, 2]
        elif r < 0.40:
            piece = [1]
        elif r < 0.50:
            piece = [-1]
        elif r < 0.65 and rest >= 7:
            k = min(rng.choice([1, 3, 3, 5]), rest // 7)
            piece = []
            for u in range(k):
                piece += [x + 1 + 7 * u for x in FORTRESS_UNIT]
        elif r < 0.85 and depth < 4 and rest >= 6:
            m = rng.randrange(1, min(rest - 5, max(2, length // 3)) + 1)
            inner = random_shape(rng, m, depth + 1)
            piece = [1, 3] + [(x + 3 if x > 0 else x - 3) for x in inner] + [4 + m, 2, 5 + m]
        else:
            m = min(rng.randrange(2, 9), rest)
            piece = random_signed_perm(m, rng)
        out += [(x + off if x > 0 else x - off) for x in piece]
    return out


def hurdle_rich_perm(n: int, seed: int) -> list[int]:
    """Config-4 style source permutation (destination = identity): a seeded mix of hurdle units,
    fortress units in odd groups, nested wraps and random-signed filler, exactly n elements."""
    p = random_shape(random.Random(seed), n)
    assert sorted(abs(x) for x in p) == list(range(1, n + 1))
    return p


def hurdle_chain(k: int) -> list[int]:
    """k hurdle units vs identity: b = 3k, c = k, h = k, f = 0, d = 3k."""
    p: list[int] = []
    for u in range(k):
        p += [3 * u + 1, 3 * u + 3, 3 * u + 2]
    return p


def fortress_chain(units: int) -> list[int]:
    """``units`` fortress units vs identity; 3 -> (b,c,h,f,d) = (18,6,3,1,16)."""
    p: list[int] = []
    for u in range(units):
        p += [x + 1 + 7 * u for x in FORTRESS_UNIT]
    return p


def all_pairs(G: int) -> np.ndarray:
    """Upper-triangular pair list in setinvmatrix order (reference G/uniinvdist.c:80-90)."""
    i, j = np.triu_indices(G, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)
