"""One launch of the persistent scenario kernel at config-4 shape (hurdle/fortress-rich pairs of 5000
blocks, thousands of pairs in flight) for ncu captures: python scripts/prof_scenario.py [pairs] [step cap]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 5000
src = np.array([gen.hurdle_rich_perm(n, s) for s in range(k)], dtype=np.int32)
dst = np.tile(np.arange(1, n + 1, dtype=np.int32), (k, 1))
capi.scenario_prefix(src[:64], dst[:64], 1)
steps, done, ev = capi.scenario_prefix(src, dst, cap)
print("ok", sum(len(x) for x in steps), int(ev.sum()))
