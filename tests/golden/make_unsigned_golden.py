#!/usr/bin/env python
"""tests/golden/unsigned_cases.json: inputs and the UNMODIFIED reference's unsigned_mcdist_nomem results
(distance, the two genomes as the reference leaves them, approximation flag) for seeded unsigned genome
pairs, linear and circular.  Needs oracle/_ref/libgrimmref.so (build container only).

Circular pairs whose first three genes of g2 form a strip that runs BACKWARDS in g1 are left out: there
the reference's find_3strip reads past the end of an array (G/unsigned.c:541-545, *g1pos still offset by
num_genes), so its result depends on the heap contents (see grsr_unsigned_host.hpp)."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util as U  # noqa: E402
from grsr_b200 import gen  # noqa: E402


def reads_out_of_bounds(a, b):
    n = len(a)
    if n < 4:
        return False
    pos = {abs(x): i for i, x in enumerate(a)}
    p = pos[abs(b[0])]
    fwd = all(abs(a[(p + k) % n]) == abs(b[k % n]) for k in range(3))
    bwd = all(abs(a[(p - k) % n]) == abs(b[k % n]) for k in range(3))
    return (not fwd) and bwd


def cases():
    for n in (1, 2, 3, 4, 5, 6, 7, 8):
        for s in range(30):
            rng = random.Random(n * 1000 + s)
            kind = s % 3
            if kind == 0:
                a = list(range(1, n + 1)); b = list(range(1, n + 1)); rng.shuffle(b)
            elif kind == 1:
                a = list(range(1, n + 1)); rng.shuffle(a)
                b = [a[abs(x) - 1] for x in gen.k_reversal_perm(n, rng.randrange(1, 4), rng)]
            else:
                a = [abs(x) for x in gen.hurdle_rich_perm(n, n * 7 + s)]; b = list(range(1, n + 1))
            if rng.random() < 0.3:
                a = [x if rng.random() < .5 else -x for x in a]
            yield a, b
    for n in (20, 40, 60, 100):
        for s in range(10):
            rng = random.Random(n * 77 + s)
            a = list(range(1, n + 1))
            if s % 2:
                b = [abs(x) for x in gen.k_reversal_perm(n, rng.randrange(2, 7), rng)]
            else:
                b = [abs(x) for x in gen.hurdle_rich_perm(n, s)]
                lim = 12 if n <= 40 else 30          # keep the singleton count (2^k evaluations) bounded
                b = b[:lim] + sorted(b[lim:])
            yield a, b


if __name__ == "__main__":
    out = []
    for a, b in cases():
        for circ in (0, 1):
            if circ and reads_out_of_bounds(a, b):
                continue
            d, g1, g2, ap = U.ref_unsigned(a, b, bool(circ))
            out.append({"g1": a, "g2": b, "circular": circ, "dist": int(d), "out_g1": g1.tolist(),
                        "out_g2": g2.tolist(), "approx": int(ap)})
    json.dump(out, open(os.path.join(HERE, "unsigned_cases.json"), "w"))
    print(len(out), "cases")
