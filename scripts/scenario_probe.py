"""Ad-hoc timing of the scenario kernel on a few workloads (used to size bench.py's scenario leg)."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import capi, gen

def run(name, srcs, dsts):
    srcs = np.asarray(srcs, dtype=np.int32); dsts = np.asarray(dsts, dtype=np.int32)
    capi.scenario_batch(srcs[:1, :], dsts[:1, :])
    t0 = time.perf_counter()
    out = capi.scenario_batch(srcs, dsts)
    dt = time.perf_counter() - t0
    steps = sum(len(x) for x in out)
    print("%-40s pairs %5d n %5d steps %8d  %.3f s  %.0f steps/s" % (name, len(srcs), srcs.shape[1], steps, dt, steps / dt), flush=True)

which = sys.argv[1:] or ["r1000", "h1000", "r2000", "r5000", "h5000"]
rng = random.Random(1)
if "r1000" in which:
    run("random n=1000 x 592", [gen.random_signed_perm(1000, rng) for _ in range(592)], [list(range(1, 1001))] * 592)
if "h1000" in which:
    run("hurdle-rich n=1000 x 592", [gen.hurdle_rich_perm(1000, s) for s in range(592)], [list(range(1, 1001))] * 592)
if "r2000" in which:
    run("random n=2000 x 592", [gen.random_signed_perm(2000, rng) for _ in range(592)], [list(range(1, 2001))] * 592)
if "r5000" in which:
    run("random n=5000 x 148", [gen.random_signed_perm(5000, rng) for _ in range(148)], [list(range(1, 5001))] * 148)
if "h5000" in which:
    run("hurdle-rich n=5000 x 148", [gen.hurdle_rich_perm(5000, s) for s in range(148)], [list(range(1, 5001))] * 148)
