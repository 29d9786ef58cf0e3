"""Diagnostic: how often the block-level walks settle an evaluation inside the scenario kernel
(needs scripts/_dbg/libgrsr_dbg.so = the library built with -DGRSR_BW_STATS:
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared
-DGRSR_BW_STATS -o scripts/_dbg/libgrsr_dbg.so grsr_b200/csrc/grsr_capi.cu)."""
import ctypes, os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from grsr_b200 import build as _b
_b.LIB = os.path.join(ROOT, "scripts", "_dbg", "libgrsr_dbg.so")
from grsr_b200 import capi, gen
L = capi.lib()
L.grsr_debug_bw_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
rng = random.Random(1)
for n, k in ((1000, 296), (2000, 148), (5000, 37)):
    srcs = np.asarray([gen.random_signed_perm(n, rng) for _ in range(k)], dtype=np.int32)
    dsts = np.asarray([list(range(1, n + 1))] * k, dtype=np.int32)
    out = (ctypes.c_ulonglong * 8)()
    L.grsr_debug_bw_stats(out, 1)
    t0 = time.perf_counter()
    res = capi.scenario_batch(srcs, dsts)
    dt = time.perf_counter() - t0
    L.grsr_debug_bw_stats(out, 1)
    steps = sum(len(x) for x in res)
    print("n=%d pairs=%d steps=%d %.3fs  no-unoriented settled=%d budget-unsettled=%d covered=%d uncovered=%d" %
          (n, k, steps, dt, out[0], out[1], out[2], out[3]), flush=True)
