"""Config 5: one pair, n = 100 000 random signed (seed 5) vs identity, full -s scenario on one GPU."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from grsr_b200 import capi, gen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = random.Random(5)
src = np.array(gen.random_signed_perm(n, rng), dtype=np.int32)
dst = np.arange(1, n + 1, dtype=np.int32)
t0 = time.perf_counter()
d = capi.dist_pairs(np.stack([src, dst]), [[0, 1]], stats=True)[0]
t1 = time.perf_counter()
print("n", n, "stats", d.tolist(), "distance call %.3f s" % (t1 - t0), flush=True)
t0 = time.perf_counter()
steps = capi.scenario_batch(src, dst)[0]
dt = time.perf_counter() - t0
cur = src.astype(np.int64).copy()
for s, e in steps:
    cur[s:e + 1] = -cur[s:e + 1][::-1]
ok = len(steps) == d[0] and np.array_equal(cur, dst)
print("scenario: %d steps in %.1f s = %.0f reversals/s; length == d and sorts: %s" % (len(steps), dt, len(steps) / dt, ok), flush=True)
