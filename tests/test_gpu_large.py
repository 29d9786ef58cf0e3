"""GPU parity tests of the global-memory ("large") build that serves genomes too long for the
shared-memory kernels: forced at small sizes (GRSR_FORCE_LARGE) for full oracle parity, and at its
natural sizes for distances vs the oracle and scenario properties."""
import os
import random

import numpy as np
import pytest

import _util as U
from grsr_b200 import capi, gen

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_large():
    os.environ["GRSR_FORCE_LARGE"] = "1"
    yield
    os.environ.pop("GRSR_FORCE_LARGE", None)


def _ordered(G):
    return np.array([[i, j] for i in range(G) for j in range(G)], dtype=np.int32)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6])
def test_forced_large_exhaustive(force_large, n):
    perms = U.all_signed_perms(n)
    g = np.vstack([np.arange(1, n + 1, dtype=np.int32)[None, :], perms])
    k = len(g)
    pairs = np.array([[0, i] for i in range(1, k)] + [[i, 0] for i in range(1, k)], dtype=np.int32)
    assert np.array_equal(capi.dist_pairs(g, pairs, stats=True), U.orc_stats(g, pairs))


@pytest.mark.parametrize("n", [17, 64, 255, 1000, 2000])
def test_forced_large_structured(force_large, n):
    g = U.mixed_genomes(n, 16 if n <= 255 else 8, seed=n)
    pairs = _ordered(len(g))
    assert np.array_equal(capi.dist_pairs(g, pairs, stats=True), U.orc_stats(g, pairs))


def test_forced_large_fortress_fixture(force_large):
    z = np.load(U.GOLDEN + "/stats_mixed.npz")
    g = z["gfort"]
    assert np.array_equal(capi.dist_pairs(g, _ordered(len(g)), stats=True), z["sfort"])


@pytest.mark.parametrize("n", [1, 2, 7, 33, 100, 300])
def test_forced_large_scenarios(force_large, n):
    srcs, dsts = [], []
    for s in range(12):
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
    got = capi.scenario_batch(np.array(srcs, dtype=np.int32), np.array(dsts, dtype=np.int32))
    for s, d, steps in zip(srcs, dsts, got):
        want, _ = U.orc_scenario(s, d)
        assert np.array_equal(steps, want)


def test_natural_large_distances():
    """Beyond the on-chip limit the library switches builds by itself."""
    assert capi.max_genes_on_chip(False) < 16000 <= capi.max_genes()
    for n, count in ((16000, 6), (40000, 3), (100000, 2)):
        rng = random.Random(n)
        rows = [list(range(1, n + 1))] + [gen.random_signed_perm(n, rng) for _ in range(count - 2)]
        rows.append(gen.hurdle_rich_perm(n, 7))
        g = np.array(rows, dtype=np.int32)
        pairs = _ordered(count)
        assert np.array_equal(capi.dist_pairs(g, pairs, stats=True), U.orc_stats(g, pairs)), n


def test_natural_large_scenario_properties():
    """n = 20000 (beyond every on-chip kernel): the scenario has exactly d steps and sorts."""
    n = 20000
    rng = random.Random(3)
    src = np.array(gen.k_reversal_perm(n, 400, rng), dtype=np.int32)
    dst = np.arange(1, n + 1, dtype=np.int32)
    steps = capi.scenario_batch(src, dst)[0]
    d = capi.dist_pairs(np.stack([src, dst]), [[0, 1]])[0]
    assert len(steps) == d and d > 300
    assert np.array_equal(U.apply_steps(src, steps), dst)


def test_too_many_genes_is_refused():
    n = capi.max_genes() + 1
    g = np.zeros((1, 4), dtype=np.int32)
    with pytest.raises(capi.GrsrError):
        capi.dist_pairs(g, [[0, 0]])
    assert n > 1000000


@pytest.fixture(params=[2, 8, 16])
def force_cluster(request):
    """global-memory build with one thread-block cluster of 2 / 8 / 16 blocks per pair"""
    os.environ["GRSR_FORCE_LARGE"] = "1"
    os.environ["GRSR_CLUSTER"] = str(request.param)
    yield request.param
    os.environ.pop("GRSR_FORCE_LARGE", None)
    os.environ.pop("GRSR_CLUSTER", None)


def test_cluster_build_distances(force_cluster):
    for n in (5, 64, 300, 2000):
        g = U.mixed_genomes(n, 8, seed=n + force_cluster)
        pairs = _ordered(len(g))
        assert np.array_equal(capi.dist_pairs(g, pairs, stats=True), U.orc_stats(g, pairs)), n
    z = np.load(U.GOLDEN + "/stats_mixed.npz")
    g = z["gfort"]
    assert np.array_equal(capi.dist_pairs(g, _ordered(len(g)), stats=True), z["sfort"])


def test_cluster_build_scenarios(force_cluster):
    for n in (3, 40, 200):
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
        got = capi.scenario_batch(np.array(srcs, dtype=np.int32), np.array(dsts, dtype=np.int32))
        for s, d, steps in zip(srcs, dsts, got):
            want, _ = U.orc_scenario(s, d)
            assert np.array_equal(steps, want), (n, force_cluster)


def test_single_long_pair_uses_a_cluster_and_sorts():
    """One random pair of n = 12000 (beyond the on-chip scenario kernels): the library gives it a
    16-block cluster; the scenario has exactly d steps and sorts."""
    n = 12000
    rng = random.Random(8)
    src = np.array(gen.random_signed_perm(n, rng), dtype=np.int32)
    dst = np.arange(1, n + 1, dtype=np.int32)
    steps = capi.scenario_batch(src, dst)[0]
    d = capi.dist_pairs(np.stack([src, dst]), [[0, 1]])[0]
    assert len(steps) == d and d > n - 20
    assert np.array_equal(U.apply_steps(src, steps), dst)
