"""Cost of the first tier on a batch it cannot settle: hurdle-rich genomes against the identity
(every pair is deferred), two-tier path vs the block-per-pair kernel alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
n, k = 2000, 20000
g = np.asarray([list(range(1, n + 1))] + [gen.hurdle_rich_perm(n, 100 + s) for s in range(400)], dtype=np.int32)
pairs = np.array([[1 + (i % 400), 0] for i in range(k)], dtype=np.int32)
for mode in ("0", "1", "0", "1"):
    os.environ["GRSR_WALK"] = mode
    capi.dist_pairs(g, pairs[:2000], stats=True)
    t0 = time.perf_counter()
    out = capi.dist_pairs(g, pairs, stats=True)
    dt = time.perf_counter() - t0
    print("GRSR_WALK=%s  %d pairs  %.1f ms  hurdles mean %.1f" % (mode, k, dt * 1e3, out[:, 3].mean()), flush=True)
