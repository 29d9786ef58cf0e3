"""Small scenario workload for ncu: 148 hurdle-rich pairs, n = 300."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
n, k = 300, 148
src = np.array([gen.hurdle_rich_perm(n, 9000 + s) for s in range(k)], dtype=np.int32)
dst = np.tile(np.arange(1, n + 1, dtype=np.int32), (k, 1))
steps = capi.scenario_batch(src, dst)
print("reversals", sum(len(x) for x in steps))
