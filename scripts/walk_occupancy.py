"""Lane occupancy of the walk kernel's tau steps, measured in the SIMT emulator (a special CPU build of
the kernel source with -DGRSR_WALK_STATS): histogram of lanes that take each step instruction, share of
steps after the node pool ran dry, start sections and walks per pair.
Usage: python scripts/walk_occupancy.py [n] [pairs]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from grsr_b200 import gen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 40
extra = sys.argv[3:]
emu = os.path.join(ROOT, "tests", "emu")
so = "/tmp/libgrsr_emu_stats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DGRSR_WALK_STATS", *extra, "-x", "c++", "-o", so,
                       os.path.join(emu, "emu_lib.cpp")], cwd=emu)
lib = ctypes.CDLL(so)
stats = (ctypes.c_ulonglong * 64).in_dll(lib, "g_walk_stats")
G = 9
g = np.ascontiguousarray(gen.random_genomes(G, n, 7), dtype=np.int32)
pairs = np.array([[i, j] for j in range(G) for i in range(j)], dtype=np.int32)[:npairs]
out = np.zeros((len(pairs), 5), dtype=np.int32)
d = ctypes.c_int(0)
P = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
lib.emu_dist_pairs_walk(P(g), G, n, P(pairs), ctypes.c_longlong(len(pairs)), P(out), 5, 2, 2, ctypes.byref(d), 1)
s = [int(x) for x in stats]
np_ = s[45]
instr = sum(s[:33]); lanes = sum(i * s[i] for i in range(33))
print("pairs %d  deferred %d  step instr/pair %.1f  lane-steps/pair %.1f  mean active %.2f" % (np_, d.value, instr / np_, lanes / np_, lanes / instr))
print("pool phase: %.1f instr/pair, mean active %.2f; dry phase: %.1f instr/pair, mean active %.2f" % (
    s[40] / np_, s[42] / max(1, s[40]), s[41] / np_, s[43] / max(1, s[41])))
print("start sections/pair %.1f  walks/pair %.1f" % (s[44] / np_, s[46] / np_))
# expected wavefronts of a random 16-bit gather by k lanes (balls in 32 bins, two entries per word ignored)
rng = np.random.default_rng(1)
ew = [0.0] + [float(np.mean([np.bincount(rng.integers(0, 32, k), minlength=32).max() for _ in range(4000)])) for k in range(1, 33)]
print("expected LDS wavefronts/pair on the steps: %.0f" % (sum(ew[i] * s[i] for i in range(33)) / np_))
print("histogram (active lanes: share of step instr):", " ".join("%d:%.1f%%" % (i, 100 * s[i] / instr) for i in range(33) if s[i]))
