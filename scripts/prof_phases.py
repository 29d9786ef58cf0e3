"""Per-phase cycle breakdown of the scenario kernels (diagnostic build with -DGRSR_PHASE_PROFILE:
thread 0 of every block charges clock64() deltas to the phase that was running).

    python scripts/prof_phases.py [hurdle|random|trial] [pairs] [cap] > profiles/<name>.txt

Build the twin library first (here, it travels with the snapshot): python -c "from grsr_b200 import build; build.build_cuda_profile()"
"""
import ctypes, os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GRSR_LIB"] = os.path.join(ROOT, "grsr_b200", "lib", "libgrsr_cuda_prof%s.so" % os.environ.get("GRSR_PROF_TAG", ""))
import numpy as np
from grsr_b200 import capi, gen

NAMES = {0: "other (staging, waits for the ticket, result)", 1: "rho0 / inverse of the current genome",
         2: "G0: phase A grey edges", 3: "G0: phase B pointer jumping", 4: "G0: phase C counts + cover test",
         5: "G0: cycle ids for the component analysis", 6: "breakpoint list (ordered compaction)",
         7: "candidate search (arg-min over breakpoint pairs)",
         8: "trial: phase A grey edges", 9: "trial: phase B pointer jumping", 10: "trial: phase C counts + cover test",
         11: "trial: cycle ids for the component analysis", 13: "apply reversal, emit step",
         20: "slow path S1 forest init", 21: "slow path S2 reach per cycle", 22: "slow path S3 block tables",
         23: "slow path S4 extreme pokers per cycle", 24: "slow path S5 union-find", 25: "slow path S6 component orientation",
         26: "slow path S7 compaction of unoriented vertices", 27: "slow path S8 runs", 28: "slow path S9 hurdles / fortress"}

def read(reset=True):
    cyc = np.zeros(64, dtype=np.uint64); cnt = np.zeros(64, dtype=np.uint64)
    fn = capi.lib().grsr_prof_read
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert fn(cyc.ctypes.data, cnt.ctypes.data, 1 if reset else 0) == 0
    return cyc.astype(np.float64), cnt.astype(np.float64)

def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "hurdle"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
    os.environ["GRSR_STEPWISE_MAX"] = "0"
    if kind == "trial":
        src = np.array(gen.hurdle_rich_perm(n, 4), dtype=np.int32); dst = np.arange(1, n + 1, dtype=np.int32)
        rng = np.random.default_rng(4)
        ci = rng.integers(0, n, size=k); cj = rng.integers(0, n, size=k)
        cands = np.stack([np.minimum(ci, cj), np.maximum(ci, cj)], axis=1).astype(np.int32)
        capi.trial_distances(src, dst, cands[:64]); read()
        t0 = time.perf_counter(); capi.trial_distances(src, dst, cands); dt = time.perf_counter() - t0
        print("trial_batch_kernel: %d candidates of one hurdle-rich pair n=%d in %.3f s = %.3g evaluations/s" % (k, n, dt, k / dt))
    else:
        if kind == "hurdle":
            src = np.array([gen.hurdle_rich_perm(n, s) for s in range(k)], dtype=np.int32)
        else:
            rng = random.Random(4)
            src = np.array([gen.random_signed_perm(n, rng) for _ in range(k)], dtype=np.int32)
        dst = np.tile(np.arange(1, n + 1, dtype=np.int32), (k, 1))
        capi.scenario_prefix(src[:2], dst[:2], 1); read()
        t0 = time.perf_counter(); steps, done, ev = capi.scenario_prefix(src, dst, cap); dt = time.perf_counter() - t0
        tot = sum(len(x) for x in steps)
        print("scenario_kernel: %d %s pairs n=%d, first %d steps: %d reversals, %d full evaluations in %.3f s"
              " = %.3g reversals/s, %.3g evaluations/s" % (k, kind, n, cap, tot, ev.sum(), dt, tot / dt, ev.sum() / dt))
    cyc, cnt = read()
    tot = cyc.sum()
    print("%-55s %8s %12s %12s" % ("phase", "share", "entries", "cycles/entry"))
    for i in range(64):
        if cyc[i] == 0 and cnt[i] == 0: continue
        print("%2d %-52s %7.2f%% %12d %12.0f" % (i, NAMES.get(i, "?"), 100 * cyc[i] / tot, cnt[i], cyc[i] / max(cnt[i], 1)))

main()
