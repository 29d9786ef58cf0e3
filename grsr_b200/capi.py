"""ctypes binding of the C ABI in include/grsr.h (libgrsr_cuda.so).

Used by tests/ and bench.py; it adds nothing to the path -- every function forwards to the library
and raises GrsrError with grsr_last_error() on a non-zero return.  If the library has not been
built this module raises at first use: there is no fallback implementation.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

I32P = ctypes.POINTER(ctypes.c_int32)

OK, ERR_INVALID, ERR_CUDA, ERR_NODEVICE, ERR_TOOLARGE, ERR_NOMEM, ERR_SEARCH = 0, -1, -2, -3, -4, -5, -6


class GrsrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"grsr error {code}: {msg}")
        self.code = code


_lib = None

# every exported symbol of include/grsr.h with its ctypes signature
SIGNATURES = {
    "grsr_abi_version": (ctypes.c_int, []),
    "grsr_last_error": (ctypes.c_char_p, []),
    "grsr_device_count": (ctypes.c_int, []),
    "grsr_max_genes": (ctypes.c_int, []),
    "grsr_max_genes_on_chip": (ctypes.c_int, [ctypes.c_int]),
    "grsr_release": (None, []),
    "grsr_dist_pairs": (ctypes.c_int, [I32P, ctypes.c_int, ctypes.c_int, I32P, ctypes.c_int64, I32P,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "grsr_dist_matrix": (ctypes.c_int, [I32P, ctypes.c_int, ctypes.c_int, I32P, ctypes.c_int, ctypes.c_int]),
    "grsr_dist_matrix_shard": (ctypes.c_int, [I32P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              I32P, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "grsr_stats_matrix": (ctypes.c_int, [I32P, ctypes.c_int, ctypes.c_int, I32P, ctypes.c_int, ctypes.c_int]),
    "grsr_scenario_batch": (ctypes.c_int, [I32P, I32P, ctypes.c_int, ctypes.c_int, I32P, I32P, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int]),
    "grsr_scenario_prefix": (ctypes.c_int, [I32P, I32P, ctypes.c_int, ctypes.c_int, I32P, I32P, ctypes.c_int,
                                            I32P, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int]),
    "grsr_unsigned_dist": (ctypes.c_int, [I32P, I32P, ctypes.c_int, ctypes.c_int, I32P, I32P, ctypes.c_void_p, ctypes.c_int]),
    "grsr_trial_distances": (ctypes.c_int, [I32P, I32P, ctypes.c_int, I32P, ctypes.c_int64, I32P, ctypes.c_int]),
    "grsr_signed_inverse_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "grsr_dist_pairs_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "grsr_dist_pairs_sorted_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "grsr_triangle_pairs_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                               ctypes.c_void_p]),
}


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("GRSR_LIB", _build.LIB)      # diagnostic builds (scripts/prof_phases.py)
        if not os.path.exists(path):
            raise GrsrError(ERR_NODEVICE, f"{path} is not built (run python -m grsr_b200.build); "
                                          "there is no CPU implementation to fall back to")
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise GrsrError(rc, lib().grsr_last_error().decode())


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(I32P)


def device_count() -> int:
    return lib().grsr_device_count()


def max_genes() -> int:
    return lib().grsr_max_genes()


def max_genes_on_chip(scenario: bool = False) -> int:
    return lib().grsr_max_genes_on_chip(1 if scenario else 0)


def release() -> None:
    lib().grsr_release()


def triangle_pairs(G: int) -> np.ndarray:
    """Pair list of the upper triangle in the library's flat order: column by column,
    (0,1), (0,2), (1,2), (0,3), ... (pair (i,j), i<j, has index j(j-1)/2 + i)."""
    j, i = np.tril_indices(G, k=-1)
    return np.stack([i, j], axis=1).astype(np.int32)


def dist_pairs(genomes, pairs, stats: bool = False, first_device: int = 0, ngpus: int = 1) -> np.ndarray:
    """Distances (or {d,b,c,h,f} rows) for pairs[p] = (gamma row, pi row)."""
    g = _i32(genomes)
    pr = _i32(pairs).reshape(-1, 2)
    stride = 5 if stats else 1
    out = np.zeros((len(pr), stride) if stats else (len(pr),), dtype=np.int32)
    _check(lib().grsr_dist_pairs(_p(g), g.shape[0], g.shape[1], _p(pr), len(pr), _p(out), stride,
                                 first_device, ngpus))
    return out


def dist_matrix(genomes, first_device: int = 0, ngpus: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    g = _i32(genomes)
    G = g.shape[0]
    if out is None:
        out = np.zeros((G, G), dtype=np.int32)
    _check(lib().grsr_dist_matrix(_p(g), G, g.shape[1], _p(out), first_device, ngpus))
    return out


def dist_matrix_shard(genomes, shard: int, num_shards: int, device: int = 0,
                      out: np.ndarray | None = None) -> np.ndarray:
    g = _i32(genomes)
    G = g.shape[0]
    P = G * (G - 1) // 2
    lo, hi = P * shard // num_shards, P * (shard + 1) // num_shards
    if out is None:
        out = np.zeros(max(hi - lo, 1), dtype=np.int32)
    cnt = ctypes.c_int64(0)
    _check(lib().grsr_dist_matrix_shard(_p(g), G, g.shape[1], shard, num_shards, _p(out),
                                        ctypes.byref(cnt), device))
    return out[:cnt.value]


def stats_matrix(genomes, first_device: int = 0, ngpus: int = 1) -> np.ndarray:
    g = _i32(genomes)
    G = g.shape[0]
    out = np.zeros((G, G, 5), dtype=np.int32)
    _check(lib().grsr_stats_matrix(_p(g), G, g.shape[1], _p(out), first_device, ngpus))
    return out


def scenario_batch(src, dst, first_device: int = 0, ngpus: int = 1) -> list[np.ndarray]:
    """For each (src[p], dst[p]) the list of reversals (start, end) GRIMM's -s prints."""
    s = _i32(src)
    d = _i32(dst)
    if s.ndim == 1:
        s = s[None, :]
        d = d[None, :]
    npairs, n = s.shape
    max_steps = n + 1
    steps = np.zeros((npairs, max_steps, 2), dtype=np.int32)
    ns = np.zeros(npairs, dtype=np.int32)
    _check(lib().grsr_scenario_batch(_p(s), _p(d), npairs, n, _p(steps), _p(ns), max_steps,
                                     first_device, ngpus))
    return [steps[p, :ns[p]].copy() for p in range(npairs)]


def scenario_prefix(src, dst, max_steps: int, first_device: int = 0, ngpus: int = 1):
    """The first max_steps reversals of every scenario: (step lists, complete flags, evaluations)."""
    s = _i32(src)
    d = _i32(dst)
    if s.ndim == 1:
        s = s[None, :]
        d = d[None, :]
    npairs, n = s.shape
    steps = np.zeros((npairs, max(max_steps, 1), 2), dtype=np.int32)
    ns = np.zeros(npairs, dtype=np.int32)
    done = np.zeros(npairs, dtype=np.int32)
    ev = np.zeros(npairs, dtype=np.int64)
    _check(lib().grsr_scenario_prefix(_p(s), _p(d), npairs, n, _p(steps), _p(ns), max_steps, _p(done),
                                      ev.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), first_device, ngpus))
    return [steps[p, :ns[p]].copy() for p in range(npairs)], done.astype(bool), ev


class UnsignedInfo(ctypes.Structure):
    _fields_ = [("dist", ctypes.c_int32), ("approx", ctypes.c_int32), ("passes", ctypes.c_int32),
                ("num_sing", ctypes.c_int32 * 2), ("num_doub", ctypes.c_int32 * 2)]


def unsigned_dist(g1, g2, circular: bool = False, device: int = 0):
    """(distance, optimally signed g1, optimally signed g2, info) of two unsigned genomes (grimm -u)."""
    a = _i32(g1)
    b = _i32(g2)
    o1 = np.zeros_like(a)
    o2 = np.zeros_like(b)
    info = UnsignedInfo()
    _check(lib().grsr_unsigned_dist(_p(a), _p(b), len(a), 1 if circular else 0, _p(o1), _p(o2),
                                    ctypes.byref(info), device))
    return info.dist, o1, o2, info


def trial_distances(src, dst, cands, device: int = 0) -> np.ndarray:
    """d(trial_c, dst) for every candidate reversal (i, j) of one pair (grimm -t 1's inner loop)."""
    s = _i32(src)
    d = _i32(dst)
    c = _i32(cands).reshape(-1, 2)
    out = np.zeros(len(c), dtype=np.int32)
    _check(lib().grsr_trial_distances(_p(s), _p(d), len(s), _p(c), len(c), _p(out), device))
    return out


# ---- device-resident entry points (raw CUDA device pointers, e.g. torch.Tensor.data_ptr()) --------

def signed_inverse_dev(d_genomes: int, G: int, n: int, d_sinv: int, stream: int) -> None:
    _check(lib().grsr_signed_inverse_dev(d_genomes, G, n, d_sinv, stream))


def dist_pairs_dev(d_genomes: int, d_sinv: int, n: int, d_pairs: int, npairs: int, d_out: int,
                   stride: int, stream: int) -> None:
    _check(lib().grsr_dist_pairs_dev(d_genomes, d_sinv, n, d_pairs, npairs, d_out, stride, stream))


def dist_pairs_sorted_dev(d_genomes: int, d_sinv: int, n: int, d_pairs: int, npairs: int, d_out: int,
                          stride: int, stream: int) -> None:
    _check(lib().grsr_dist_pairs_sorted_dev(d_genomes, d_sinv, n, d_pairs, npairs, d_out, stride, stream))


def triangle_pairs_dev(G: int, first: int, count: int, d_pairs: int, stream: int) -> None:
    _check(lib().grsr_triangle_pairs_dev(G, first, count, d_pairs, stream))
