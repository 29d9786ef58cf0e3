"""The kernel SOURCE (grsr_core.cuh / grsr_scenario.cuh) executed on the CPU by the SIMT emulator
(tests/emu) and compared with the oracle.  This does not replace the GPU parity tests -- it lets the
kernel logic (indexing, barriers, warp collectives, every block shape) be checked where no GPU
exists.  Sizes are small because each emulated thread is a fiber."""
import random

import numpy as np
import pytest

import _util as U
from grsr_b200 import gen


def _ordered(G):
    return np.array([[i, j] for i in range(G) for j in range(G)], dtype=np.int32)


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_emulated_distance_exhaustive(n):
    perms = U.all_signed_perms(n)
    g = np.vstack([np.arange(1, n + 1, dtype=np.int32)[None, :], perms])
    k = len(g)
    pairs = np.array([[0, i] for i in range(1, k)] + [[i, 0] for i in range(1, k)], dtype=np.int32)
    assert np.array_equal(U.emu_stats(g, pairs, 0, 3), U.orc_stats(g, pairs))


# (n, forced threads): 0 = the shape pick_shape chooses; the others force other kernel instances
@pytest.mark.parametrize("n,threads", [(7, 0), (15, 0), (16, 0), (33, 0), (63, 0), (64, 0), (127, 0),
                                        (255, 0), (300, 128), (300, 256), (600, 0), (1100, 0), (1100, 512)])
def test_emulated_distance_structured(n, threads):
    g = U.mixed_genomes(n, 12 if n < 200 else 6, seed=n)
    pairs = _ordered(len(g))
    got = U.emu_stats(g, pairs, threads, 3)
    assert np.array_equal(got, U.orc_stats(g, pairs))


def test_emulated_fortress_fixture():
    z = np.load(U.GOLDEN + "/stats_mixed.npz")
    g = z["gfort"]
    assert np.array_equal(U.emu_stats(g, _ordered(len(g)), 0, 2), z["sfort"])


@pytest.mark.parametrize("n", [1, 2, 5, 12, 29, 40])
def test_emulated_scenarios(n):
    srcs, dsts = [], []
    for s in range(8):
        seed = n * 100 + s
        rng = random.Random(seed)
        kind = s % 4
        if kind == 0:
            src, dst = gen.random_signed_perm(n, rng), gen.random_signed_perm(n, rng)
        elif kind == 1:
            src, dst = gen.hurdle_rich_perm(n, seed), list(range(1, n + 1))
        elif kind == 2:
            src, dst = list(range(1, n + 1)), gen.hurdle_rich_perm(n, seed)
        else:
            src = gen.k_reversal_perm(n, rng.randrange(1, 8), rng)
            dst = gen.k_reversal_perm(n, rng.randrange(0, 5), rng)
        srcs.append(src)
        dsts.append(dst)
    steps, status, evals = U.emu_scenario(srcs, dsts, 0, 3)
    for q in range(len(srcs)):
        want, want_evals = U.orc_scenario(srcs[q], dsts[q])
        assert status[q] == 0
        assert np.array_equal(steps[q], want)
        assert evals[q] == want_evals      # same candidates tried in the same order


# ---- the global-memory build (one block per pair) and its cluster variant (several blocks per pair) --

@pytest.mark.parametrize("cluster", [0, 2, 3])
def test_emulated_large_build_distances(cluster):
    for n in (3, 17, 64, 100):
        g = U.mixed_genomes(n, 8, seed=n)
        pairs = _ordered(len(g))
        got = U.emu_stats(g, pairs, 64, 2, large=True, cluster=cluster)
        assert np.array_equal(got, U.orc_stats(g, pairs)), (n, cluster)
    z = np.load(U.GOLDEN + "/stats_mixed.npz")
    g = z["gfort"]
    assert np.array_equal(U.emu_stats(g, _ordered(len(g)), 64, 2, large=True, cluster=cluster), z["sfort"])


@pytest.mark.parametrize("cluster", [0, 2, 4])
def test_emulated_large_build_scenarios(cluster):
    for n in (2, 8, 21):
        srcs, dsts = [], []
        for s in range(8):
            seed = n * 100 + s
            rng = random.Random(seed)
            if s % 2 == 0:
                src, dst = gen.random_signed_perm(n, rng), gen.random_signed_perm(n, rng)
            else:
                src, dst = gen.hurdle_rich_perm(n, seed), list(range(1, n + 1))
            srcs.append(src)
            dsts.append(dst)
        steps, status, evals = U.emu_scenario(srcs, dsts, 32, 2, large=True, cluster=cluster)
        for q in range(len(srcs)):
            want, want_evals = U.orc_scenario(srcs[q], dsts[q])
            assert status[q] == 0 and np.array_equal(steps[q], want) and evals[q] == want_evals


def test_emulated_trial_distances():
    for n in (5, 17, 40):
        for s in range(4):
            rng = random.Random(n * 10 + s)
            src = gen.hurdle_rich_perm(n, n * 10 + s) if s % 2 else gen.random_signed_perm(n, rng)
            dst = list(range(1, n + 1)) if s < 2 else gen.random_signed_perm(n, rng)
            cands = U.breakpoint_candidates(src, dst)
            if len(cands):
                assert np.array_equal(U.emu_trial_distances(src, dst, cands), U.orc_trial_distances(src, dst, cands))


def test_emulated_stepwise_scenarios():
    """Candidate-parallel steps while hurdles exist (prepare / score / select kernels driven by the
    same host loop as the C ABI), then the persistent kernel: the reference's step lists, also when
    the candidate buffer is small enough to force several batches per step."""
    for n in (2, 5, 13, 29):
        srcs, dsts = [], []
        for s in range(8):
            seed = n * 100 + s
            rng = random.Random(seed)
            if s % 2 == 0:
                src, dst = gen.hurdle_rich_perm(n, seed), list(range(1, n + 1))
            else:
                src, dst = gen.random_signed_perm(n, rng), gen.hurdle_rich_perm(n, seed)
            srcs.append(src)
            dsts.append(dst)
        wants = [U.orc_scenario(s, d) for s, d in zip(srcs, dsts)]
        ref_evals = sum(w[1] for w in wants)
        ref_steps = sum(len(w[0]) for w in wants)
        for cap in (0, 2 * (n + 1), 8, 3):              # 8 and 3: batches end inside diagonals, many per step
            steps, status, stepwise_steps = U.emu_scenario_stepwise(srcs, dsts, 0, cap)
            for q in range(len(srcs)):
                assert status[q] == 0 and np.array_equal(steps[q], wants[q][0])
            assert n < 5 or stepwise_steps > 0
            # no candidate is scored twice within a step: the candidate-parallel phase scores at most what
            # the reference's sequential scan scores, plus less than one batch per step
            scored = U.emulator().emu_last_stepwise_evals()
            if cap in (3, 8):
                assert scored <= ref_evals + cap * ref_steps, (n, cap, scored, ref_evals, ref_steps)


# ---- the two-tier distance path: warp-per-pair walk kernel + deferred pairs -------------------------

@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_emulated_walk_exhaustive(n):
    perms = U.all_signed_perms(n)
    g = np.vstack([np.arange(1, n + 1, dtype=np.int32)[None, :], perms])
    k = len(g)
    pairs = np.array([[0, i] for i in range(1, k)] + [[i, 0] for i in range(1, k)], dtype=np.int32)
    got, _ = U.emu_stats_walk(g, pairs, 2, 2)
    assert np.array_equal(got, U.orc_stats(g, pairs))


@pytest.mark.parametrize("n,wpb", [(5, 1), (7, 2), (15, 1), (16, 4), (33, 2), (63, 1), (64, 2), (127, 3),
                                    (255, 1), (256, 2), (300, 2), (600, 1), (1100, 2), (2100, 1)])
def test_emulated_walk_structured(n, wpb):
    """every node-stride / words-per-lane shape; hurdle-rich rows are deferred to the second tier,
    random ones are settled by the walk kernel"""
    g = U.mixed_genomes(n, 12 if n < 200 else (6 if n < 2000 else 4), seed=n)
    pairs = _ordered(len(g))
    got, deferred = U.emu_stats_walk(g, pairs, wpb, 2)
    want = U.orc_stats(g, pairs)
    assert np.array_equal(got, want)
    easy = int(np.sum((want[:, 3] == 0)))
    assert deferred <= len(pairs) - easy // 2      # the first tier settles most hurdle-free pairs


def test_emulated_walk_random_pairs_not_deferred():
    rng = random.Random(11)
    g = np.asarray([gen.random_signed_perm(400, rng) for _ in range(12)], dtype=np.int32)
    pairs = _ordered(len(g))
    got, deferred = U.emu_stats_walk(g, pairs, 2, 3)
    assert np.array_equal(got, U.orc_stats(g, pairs))
    assert deferred <= len(pairs) // 20


def test_emulated_walk_fortress_fixture():
    z = np.load(U.GOLDEN + "/stats_mixed.npz")
    g = z["gfort"]
    got, _ = U.emu_stats_walk(g, _ordered(len(g)), 2, 2)
    assert np.array_equal(got, z["sfort"])


@pytest.mark.parametrize("n,wpb", [(3, 2), (9, 3), (40, 2), (129, 3), (300, 2)])
def test_emulated_walk_shared_rows(n, wpb):
    """the kernel whose warps share one staged pi row per block, on a pi-sorted list (column-by-column
    triangle, runs of every length) and -- forced -- on an unsorted one; automatic choice by
    walk_order_kernel gives the same numbers"""
    g = U.mixed_genomes(n, 14, seed=50 + n)
    tri = np.array([[i, j] for j in range(len(g)) for i in range(j)], dtype=np.int32)
    want = U.orc_stats(g, tri)
    for shared in (1, -1, 0):
        got, _ = U.emu_stats_walk(g, tri, wpb, 3, shared=shared)
        assert np.array_equal(got, want), shared
    sq = _ordered(len(g))
    got, _ = U.emu_stats_walk(g, sq, wpb, 2, shared=1)
    assert np.array_equal(got, U.orc_stats(g, sq))


@pytest.mark.parametrize("n,count,seed", [(4200, 4, 11), (6000, 3, 12), (8000, 3, 13)])
def test_emu_block_shapes_with_1024_threads(n, count, seed):
    """8192 < 2n+2 <= 16384 (config 4's size range) runs 1024-thread blocks with 10 / 12 / 16 jumping
    words per thread (pick_shape); hurdle-rich and random rows, every ordered pair, vs the oracle."""
    g = U.mixed_genomes(n, count, seed=seed)
    pairs = np.array([[i, j] for i in range(count) for j in range(count)], dtype=np.int32)
    assert np.array_equal(U.emu_stats(g, pairs, threads=0, grid=2), U.orc_stats(g, pairs))


@pytest.mark.parametrize("shift", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_walk_node_grid(shift):
    """WalkGrid: node k sits in slot k (vertex >> shift == k) on an even vertex, is_node() recognises exactly
    those vertices, and the H entries (16-bit index vertex + 1) of ANY 32 consecutive nodes lie in 32
    different 4-byte banks -- the warp examines 32 nodes per shared-memory load."""
    emu = U.emulator()
    K = 300
    v = [emu.emu_walk_grid_vertex(shift, k) for k in range(K)]
    for k, x in enumerate(v):
        assert x % 2 == 0 and x >> shift == k
        assert emu.emu_walk_grid_is_node(shift, x) == 1
    nodes = set(v)
    for x in range(0, min(v[-1], 4000), 2):
        assert emu.emu_walk_grid_is_node(shift, x) == (1 if x in nodes else 0)
    for k0 in range(K - 32):
        banks = {((x + 1) >> 1) % 32 for x in v[k0:k0 + 32]}
        assert len(banks) == 32, (shift, k0)
