"""One launch of trial_batch_kernel at config-4 size (for ncu source-level captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
n = 5000
src = np.array(gen.hurdle_rich_perm(n, 4), dtype=np.int32); dst = np.arange(1, n + 1, dtype=np.int32)
rng = np.random.default_rng(4)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ci = rng.integers(0, n, size=k); cj = rng.integers(0, n, size=k)
cands = np.stack([np.minimum(ci, cj), np.maximum(ci, cj)], axis=1).astype(np.int32)
capi.trial_distances(src, dst, cands[:64])
d = capi.trial_distances(src, dst, cands)
print("ok", int(d.sum()))
