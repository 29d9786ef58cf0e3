"""Latency of ONE hurdle-rich pair (the CLI case): candidate-parallel steps vs persistent kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen
for n in (300, 1000):
    src = np.array(gen.hurdle_rich_perm(n, 1), dtype=np.int32)
    dst = np.arange(1, n + 1, dtype=np.int32)
    capi.scenario_batch(src[:50], np.arange(1, 51, dtype=np.int32)) if False else None
    for mode in ("stepwise", "persistent"):
        if mode == "persistent":
            os.environ["GRSR_STEPWISE_MAX"] = "0"
        else:
            os.environ.pop("GRSR_STEPWISE_MAX", None)
        capi.scenario_batch(src, dst) if n == 300 else None
        t0 = time.perf_counter()
        steps = capi.scenario_batch(src, dst)[0]
        dt = time.perf_counter() - t0
        print("n=%d %-10s %d reversals in %.3f s" % (n, mode, len(steps), dt), flush=True)
os.environ.pop("GRSR_STEPWISE_MAX", None)
