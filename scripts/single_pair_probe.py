"""One hurdle-rich pair alone (the command-line case): candidate-parallel steps.  Seconds and steps/s for
n = 300, 1000, 5000 (first 400 steps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
for n, cap in ((300, 0), (1000, 0), (5000, 400)):
    src = np.array(gen.hurdle_rich_perm(n, 1), dtype=np.int32); dst = np.arange(1, n + 1, dtype=np.int32)
    capi.scenario_prefix(src, dst, 2)
    t0 = time.perf_counter()
    steps, done, ev = capi.scenario_prefix(src, dst, cap or n + 1)
    dt = time.perf_counter() - t0
    print("n=%d: %d steps in %.3f s = %.0f steps/s (%.0f us per step), %d evaluations" % (n, len(steps[0]), dt, len(steps[0]) / dt, dt / max(1, len(steps[0])) * 1e6, int(ev[0])))
