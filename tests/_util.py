"""Test helpers: ctypes bindings of the oracle (oracle/liboracle.so), of the reference harness
(oracle/_ref/libgrimmref.so, when it has been built) and of the emulated kernels."""
from __future__ import annotations

import ctypes
import itertools
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I32P = ctypes.POINTER(ctypes.c_int)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "grimm")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so")
EXAMPLE = os.path.join(GOLDEN, "example_mgr_macro.txt")


def _p(a):
    return a.ctypes.data_as(I32P)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_orc = None


def oracle():
    global _orc
    if _orc is None:
        _orc = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    return _orc


def have_reference() -> bool:
    return os.path.exists(REF_LIB) and os.path.exists(REF_BIN)


_ref = None


def reference():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_LIB)
    return _ref


def orc_stats(genomes, pairs) -> np.ndarray:
    g = _i32(genomes)
    pr = _i32(pairs).reshape(-1, 2)
    out = np.zeros((len(pr), 5), dtype=np.int32)
    oracle().orc_dist_pairs(_p(g), g.shape[1], _p(pr), ctypes.c_long(len(pr)), _p(out), 5)
    return out


def ref_stats(genomes, pairs, threads: int = 4) -> np.ndarray:
    g = _i32(genomes)
    pr = _i32(pairs).reshape(-1, 2)
    out = np.zeros((len(pr), 5), dtype=np.int32)
    reference().ref_dist_pairs(_p(g), g.shape[1], _p(pr), ctypes.c_long(len(pr)), _p(out), 5, threads)
    return out


def orc_scenario(src, dst):
    s = _i32(src)
    d = _i32(dst)
    n = len(s)
    steps = np.zeros(2 * (n + 2), dtype=np.int32)
    ev = np.zeros(n + 2, dtype=np.int64)
    k = oracle().orc_scenario(_p(s), _p(d), n, _p(steps), n + 2,
                              ev.ctypes.data_as(ctypes.POINTER(ctypes.c_long)))
    assert k >= 0, k
    return steps[:2 * k].reshape(-1, 2).copy(), int(ev[:k].sum())


def ref_scenario(src, dst):
    s = _i32(src)
    d = _i32(dst)
    n = len(s)
    steps = np.zeros(2 * (n + 2), dtype=np.int32)
    k = reference().ref_scenario(_p(s), _p(d), n, _p(steps), n + 2)
    assert k >= 0, k
    return steps[:2 * k].reshape(-1, 2).copy()


def orc_trial_distances(src, dst, cands) -> np.ndarray:
    s = _i32(src)
    d = _i32(dst)
    c = _i32(cands).reshape(-1, 2)
    out = np.zeros(len(c), dtype=np.int32)
    oracle().orc_trial_distances(_p(s), _p(d), len(s), _p(c), ctypes.c_long(len(c)), _p(out))
    return out


def breakpoint_candidates(src, dst):
    """(i, j) pairs print_rev_stats_t1 tries (G/testrev.c:156-184): i = 0 or a breakpoint start,
    j = n-1 or a breakpoint end, i <= j, in i-major order."""
    n = len(src)
    succ = {}
    for a, b in zip(dst[:-1], dst[1:]):
        succ[int(a)] = int(b)
        succ[-int(b)] = -int(a)
    starts = [i for i in range(n) if i == 0 or succ.get(int(src[i - 1])) != int(src[i])]
    ends = [j for j in range(n) if j == n - 1 or succ.get(int(src[j])) != int(src[j + 1])]
    return np.array([[i, j] for i in starts for j in ends if j >= i], dtype=np.int32).reshape(-1, 2)


def emu_trial_distances(src, dst, cands, threads: int = 0, grid: int = 3) -> np.ndarray:
    s = _i32(src)
    d = _i32(dst)
    c = _i32(cands).reshape(-1, 2)
    out = np.zeros(len(c), dtype=np.int32)
    emulator().emu_trial_distances(_p(s), _p(d), len(s), _p(c), ctypes.c_longlong(len(c)), _p(out), threads, grid)
    return out


def orc_circular_align(genomes) -> np.ndarray:
    g = _i32(genomes).copy()
    oracle().orc_circular_align(_p(g), g.shape[0], g.shape[1])
    return g


def apply_steps(src, steps) -> np.ndarray:
    cur = np.array(src, dtype=np.int64)
    for s, e in steps:
        cur[s:e + 1] = -cur[s:e + 1][::-1]
    return cur


def all_signed_perms(n: int) -> np.ndarray:
    rows = []
    for p in itertools.permutations(range(1, n + 1)):
        for s in range(1 << n):
            rows.append([(-x if (s >> k) & 1 else x) for k, x in enumerate(p)])
    return np.asarray(rows, dtype=np.int32)


_emu = None


def emulator():
    global _emu
    if _emu is None:
        from grsr_b200 import build
        build.build_emu()
        _emu = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libgrsr_emu.so"))
        _emu.emu_last_stepwise_evals.restype = ctypes.c_longlong
    return _emu


def emu_stats(genomes, pairs, threads: int = 64, grid: int = 2, large: bool = False,
              cluster: int = 0) -> np.ndarray:
    g = _i32(genomes)
    pr = _i32(pairs).reshape(-1, 2)
    out = np.zeros((len(pr), 5), dtype=np.int32)
    if cluster:
        emulator().emu_dist_pairs_cluster(_p(g), g.shape[0], g.shape[1], _p(pr), ctypes.c_longlong(len(pr)),
                                          _p(out), 5, threads, grid, cluster)
        return out
    fn = emulator().emu_dist_pairs_large if large else emulator().emu_dist_pairs
    fn(_p(g), g.shape[0], g.shape[1], _p(pr), ctypes.c_longlong(len(pr)), _p(out),
                              5, threads, grid)
    return out


def emu_stats_walk(genomes, pairs, wpb: int = 2, grid: int = 2, shared: int = 0):
    """Two-tier distance path (walk kernel + deferred pairs) in the emulator: (stats, #deferred).
    shared: 0 = every warp stages its own pi row, 1 = the warps of a block share one, -1 = decided
    by walk_order_kernel as in the library."""
    g = _i32(genomes)
    pr = _i32(pairs).reshape(-1, 2)
    out = np.full((len(pr), 5), -7, dtype=np.int32)
    deferred = ctypes.c_int(0)
    emulator().emu_dist_pairs_walk(_p(g), g.shape[0], g.shape[1], _p(pr), ctypes.c_longlong(len(pr)),
                                   _p(out), 5, wpb, grid, ctypes.byref(deferred), shared)
    return out, deferred.value


def emu_scenario(srcs, dsts, threads: int = 64, grid: int = 2, large: bool = False, cluster: int = 0):
    S = _i32(srcs)
    D = _i32(dsts)
    npairs, n = S.shape
    ms = n + 1
    steps = np.zeros((npairs, ms, 2), dtype=np.int32)
    ns = np.zeros(npairs, dtype=np.int32)
    st = np.zeros(npairs, dtype=np.int32)
    ev = np.zeros(npairs, dtype=np.uint64)
    evp = ev.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong))
    if cluster:
        emulator().emu_scenario_cluster(_p(S), _p(D), npairs, n, _p(steps), _p(ns), ms, _p(st), evp,
                                        threads, grid, cluster)
    else:
        fn = emulator().emu_scenario_large if large else emulator().emu_scenario
        fn(_p(S), _p(D), npairs, n, _p(steps), _p(ns), ms, _p(st), evp, threads, grid)
    return [steps[p, :ns[p]].copy() for p in range(npairs)], st, ev


def emu_scenario_stepwise(srcs, dsts, threads: int = 0, cap: int = 0):
    S = _i32(srcs)
    D = _i32(dsts)
    npairs, n = S.shape
    ms = n + 1
    cap = cap or max(64, 2 * (n + 1))
    steps = np.zeros((npairs, ms, 2), dtype=np.int32)
    ns = np.zeros(npairs, dtype=np.int32)
    st = np.zeros(npairs, dtype=np.int32)
    sw = ctypes.c_int(0)
    emulator().emu_scenario_stepwise(_p(S), _p(D), npairs, n, _p(steps), _p(ns), ms, _p(st), threads, cap,
                                     ctypes.byref(sw))
    return [steps[p, :ns[p]].copy() for p in range(npairs)], st, sw.value


def mixed_genomes(n: int, count: int, seed: int) -> np.ndarray:
    """identity + a seeded mix of random, few-reversal and hurdle/fortress-rich genomes."""
    import random

    from grsr_b200 import gen
    rows = [list(range(1, n + 1))]
    for s in range(count - 1):
        rng = random.Random(seed * 7919 + s)
        kind = s % 4
        if kind == 0:
            rows.append(gen.hurdle_rich_perm(n, seed * 7919 + s))
        elif kind == 1:
            rows.append(gen.random_signed_perm(n, rng))
        elif kind == 2:
            rows.append(gen.k_reversal_perm(n, rng.randrange(1, 10), rng))
        else:
            p = gen.hurdle_rich_perm(n, seed * 7919 + s)
            q = gen.random_signed_perm(n, rng)
            rows.append([(q[abs(x) - 1] if x > 0 else -q[abs(x) - 1]) for x in p])
    return np.asarray(rows, dtype=np.int32)


def parse_genomes(path):
    """Minimal reader of the `>name / genes / $` format for unichromosomal fixtures
    (names, int32 table); comments and `$ ; ( ) { }` are skipped."""
    names, rows, cur = [], [], None
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            if line.startswith(">"):
                names.append(line[1:])
                cur = []
                rows.append(cur)
                continue
            for tok in line.replace("$", " ").replace(";", " ").split():
                cur.append(int(tok))
    return names, np.asarray(rows, dtype=np.int32)


def run_cli(binary, args, cwd):
    import subprocess
    r = subprocess.run([binary] + list(args), cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout.decode(), r.stderr.decode()


def normalise_report(text: str) -> str:
    """Drop the two header lines that echo argv[0] and the input path."""
    keep = [ln for ln in text.splitlines() if not ln.startswith("Running:")]
    return "\n".join(keep) + ("\n" if text.endswith("\n") else "")


def error_head(err: str) -> str:
    """The reference prints its error text, then a blank line and the usage screen: keep the text."""
    return err.split("\n\nGRIMM", 1)[0].rstrip("\n")


def ref_unsigned(g1, g2, circular=False):
    """(distance, signed g1, signed g2, approx) from the reference's unsigned_mcdist_nomem."""
    a, b = _i32(g1), _i32(g2)
    o1, o2 = np.zeros_like(a), np.zeros_like(b)
    ap = ctypes.c_int(0)
    d = reference().ref_unsigned_dist(_p(a), _p(b), len(a), 1 if circular else 0, _p(o1), _p(o2), ctypes.byref(ap))
    return d, o1, o2, ap.value


def emu_unsigned(g1, g2, circular=False, threads=0, grid=3):
    a, b = _i32(g1), _i32(g2)
    o1, o2 = np.zeros_like(a), np.zeros_like(b)
    info = np.zeros(7, dtype=np.int32)
    rc = emulator().emu_unsigned_dist(_p(a), _p(b), len(a), 1 if circular else 0, _p(o1), _p(o2), _p(info), threads, grid)
    assert rc == 0
    return int(info[0]), o1, o2, info
