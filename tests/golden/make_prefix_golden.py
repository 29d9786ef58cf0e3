#!/usr/bin/env python
"""Reference step PREFIXES for scenarios too long to run whole on a CPU (configs 4 and 5 of
BASELINE.json and the sizes where the global-memory / cluster build takes over).

Run in the build container (needs oracle/_ref/libgrimmref.so = the unmodified reference plus the
thin caller oracle/ref_harness.c).  ref_scenario_prefix runs the reference's print_scenario_2 in a
forked child and stops it after the wanted number of "Step k:" lines.  Writes
tests/golden/scenario_prefixes.npz (committed; the GPU box only reads it):

  cfg5_steps            config 5: n = 100000, gen.random_signed_perm(n, Random(5)) -> identity
  rand12000_steps / rand20000_steps   random pairs at the sizes served by the global-memory build
  hr5000_<seed>_steps   config 4: gen.hurdle_rich_perm(5000, seed) -> identity, seed = pair index;
                        seed 0 is run to the end, seeds 1..7 for their first 400 steps

Inputs are regenerated from the seeds by the tests; only the step lists are stored.
"""
import ctypes
import os
import random
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from grsr_b200 import gen  # noqa: E402

I32P = ctypes.POINTER(ctypes.c_int)
CASES = {
    "cfg5": ("random", 100000, 5, 2000),
    "rand12000": ("random", 12000, 12, 1500),
    "rand20000": ("random", 20000, 20, 1500),
}
CASES["hr5000_0"] = ("hurdle", 5000, 0, 6000)          # one WHOLE config-4 scenario (3552 steps, an hour of one core)
for s in range(1, 8):
    CASES["hr5000_%d" % s] = ("hurdle", 5000, s, 400)


def case_input(kind, n, seed):
    if kind == "random":
        src = gen.random_signed_perm(n, random.Random(seed))
    else:
        src = gen.hurdle_rich_perm(n, seed)
    return np.array(src, dtype=np.int32), np.arange(1, n + 1, dtype=np.int32)


def run(name):
    kind, n, seed, cap = CASES[name]
    src, dst = case_input(kind, n, seed)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so"))
    steps = np.zeros(2 * cap, dtype=np.int32)
    t0 = time.perf_counter()
    k = lib.ref_scenario_prefix(src.ctypes.data_as(I32P), dst.ctypes.data_as(I32P), n,
                                steps.ctypes.data_as(I32P), cap)
    return name, steps[:2 * k].reshape(-1, 2).copy(), time.perf_counter() - t0


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    out_path = os.path.join(HERE, "scenario_prefixes.npz")
    have = dict(np.load(out_path)) if os.path.exists(out_path) else {}
    with ProcessPoolExecutor(max_workers=min(len(names), os.cpu_count() or 1)) as ex:
        for name, steps, dt in ex.map(run, names):
            print("%s: %d steps in %.1f s" % (name, len(steps), dt), flush=True)
            have[name + "_steps"] = steps
    np.savez_compressed(out_path, **have)
