"""In-process multi-GPU check: the sharded matrix / scenario batch equals the single-GPU result."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
n_dev = capi.device_count()
g = gen.random_genomes(120, 700, 3)
one = capi.dist_matrix(g)
for k in range(2, n_dev + 1):
    assert np.array_equal(capi.dist_matrix(g, ngpus=k), one), k
    assert np.array_equal(capi.stats_matrix(g[:20], ngpus=k), capi.stats_matrix(g[:20])), k
rng = random.Random(2)
src = np.array([gen.random_signed_perm(200, rng) for _ in range(64)], dtype=np.int32)
dst = np.array([gen.hurdle_rich_perm(200, s) for s in range(64)], dtype=np.int32)
a = capi.scenario_batch(src, dst)
b = capi.scenario_batch(src, dst, ngpus=n_dev)
assert all(np.array_equal(x, y) for x, y in zip(a, b))
print("multi-GPU in-process check ok on", n_dev, "devices")
