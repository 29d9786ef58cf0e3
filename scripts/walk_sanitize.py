"""Small two-tier distance job (sized for a run under compute-sanitizer where that is available): both walk kernels, mixed
genomes (random, hurdle-rich, near-identity), results checked against the block-per-pair kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _util as U
from grsr_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = U.mixed_genomes(n, 20, seed=5)
G = len(g)
sq = np.array([[i, j] for i in range(G) for j in range(G)], dtype=np.int32)
tri = np.array([[i, j] for j in range(G) for i in range(j)], dtype=np.int32)
os.environ["GRSR_WALK"] = "0"
want_sq, want_tri = capi.dist_pairs(g, sq, stats=True), capi.dist_pairs(g, tri, stats=True)
os.environ["GRSR_WALK"] = "1"
os.environ["GRSR_WALK_SHARED"] = "0"
assert np.array_equal(capi.dist_pairs(g, sq, stats=True), want_sq)
os.environ["GRSR_WALK_SHARED"] = "1"
assert np.array_equal(capi.dist_pairs(g, tri, stats=True), want_tri)
print("ok", len(sq), len(tri))
