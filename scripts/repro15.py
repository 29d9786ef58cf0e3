import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, _util as U
from grsr_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
g = U.mixed_genomes(n, 24, seed=n)
pairs = np.array([[i, j] for i in range(len(g)) for j in range(len(g))], dtype=np.int32)
got = capi.dist_pairs(g, pairs, stats=True)
print("ok", np.array_equal(got, U.orc_stats(g, pairs)))
