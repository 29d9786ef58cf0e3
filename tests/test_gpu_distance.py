"""GPU parity tests of the batched distance path, through the C ABI, against the oracle."""
import numpy as np
import pytest

import _util as U
from grsr_b200 import capi, gen

pytestmark = pytest.mark.gpu


def _all_ordered(G):
    return np.array([[i, j] for i in range(G) for j in range(G)], dtype=np.int32)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6])
def test_exhaustive_small_n(n):
    perms = U.all_signed_perms(n)
    g = np.vstack([np.arange(1, n + 1, dtype=np.int32)[None, :], perms])
    k = len(g)
    pairs = np.array([[0, i] for i in range(1, k)] + [[i, 0] for i in range(1, k)], dtype=np.int32)
    got = capi.dist_pairs(g, pairs, stats=True)
    assert np.array_equal(got, U.orc_stats(g, pairs))


def test_exhaustive_n7_vs_identity():
    perms = U.all_signed_perms(7)
    g = np.vstack([np.arange(1, 8, dtype=np.int32)[None, :], perms])
    pairs = np.array([[0, i] for i in range(1, len(g))], dtype=np.int32)
    got = capi.dist_pairs(g, pairs, stats=True)
    assert np.array_equal(got, U.orc_stats(g, pairs))
    # reference -F 7 -L histogram of d
    assert list(np.bincount(got[:, 0])) == [1, 28, 448, 5586, 35612, 138229, 262688, 202464, 64]


@pytest.mark.parametrize("n", [7, 8, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 129, 255, 500, 1000, 2000])
def test_structured_and_random(n):
    g = U.mixed_genomes(n, 24 if n <= 500 else 10, seed=n)
    pairs = _all_ordered(len(g))
    got = capi.dist_pairs(g, pairs, stats=True)
    want = U.orc_stats(g, pairs)
    assert np.array_equal(got, want)
    assert np.array_equal(capi.dist_pairs(g, pairs), want[:, 0])


def test_hurdle_and_fortress_vectors():
    for k, want in ((5, (15, 15, 5, 5, 0)),):
        p = gen.hurdle_chain(k)
        g = np.array([list(range(1, len(p) + 1)), p], dtype=np.int32)
        assert tuple(capi.dist_pairs(g, [[0, 1]], stats=True)[0]) == want
    for units, want in ((3, (16, 18, 6, 3, 1)), (33, (166, 198, 66, 33, 1))):
        p = gen.fortress_chain(units)
        g = np.array([list(range(1, len(p) + 1)), p], dtype=np.int32)
        assert tuple(capi.dist_pairs(g, [[0, 1]], stats=True)[0]) == want


def test_example_matrices_config1():
    g = U.orc_circular_align(U.parse_genomes(U.EXAMPLE)[1])
    assert capi.dist_matrix(g).tolist() == [[0, 20, 12], [20, 0, 15], [12, 15, 0]]
    st = capi.stats_matrix(g)
    assert st[:, :, 1].tolist() == [[0, 25, 16], [25, 0, 21], [16, 21, 0]]
    assert st[:, :, 2].tolist() == [[0, 6, 5], [6, 0, 6], [5, 6, 0]]
    assert st[:, :, 3].tolist() == [[0, 1, 1], [1, 0, 0], [1, 0, 0]]
    assert st[:, :, 4].sum() == 0


def test_config2_matrix_25x500():
    g = U.orc_circular_align(gen.random_genomes(25, 500, 1))
    m = capi.dist_matrix(g)
    pairs = gen.all_pairs(25)
    want = U.orc_stats(g, pairs)[:, 0]
    assert np.array_equal(m[pairs[:, 0], pairs[:, 1]], want)
    assert np.array_equal(m, m.T) and not m.diagonal().any()
    golden = np.loadtxt(U.GOLDEN + "/config2_matrix.txt", dtype=np.int32)
    assert np.array_equal(m, golden)


def test_matrix_shards_equal_whole():
    g = gen.random_genomes(40, 300, 3)
    whole = capi.dist_matrix(g)
    pairs = capi.triangle_pairs(40)
    assert pairs[:4].tolist() == [[0, 1], [0, 2], [1, 2], [0, 3]]
    flat = whole[pairs[:, 0], pairs[:, 1]]
    for k in (2, 3, 8):
        parts = [capi.dist_matrix_shard(g, s, k) for s in range(k)]
        assert np.array_equal(np.concatenate(parts), flat)


def test_config3_sample_and_properties():
    """Full-size config 3 table (1000 x 2000): a 3000-pair sample against the oracle, plus
    symmetry d(a,b) == d(b,a) and d(a,a) == 0 on the GPU path itself."""
    g = gen.random_genomes(1000, 2000, 7)
    rng = np.random.default_rng(0)
    pairs = rng.integers(0, 1000, size=(3000, 2)).astype(np.int32)
    got = capi.dist_pairs(g, pairs, stats=True)
    assert np.array_equal(got, U.orc_stats(g, pairs))
    rev = capi.dist_pairs(g, pairs[:, ::-1])
    assert np.array_equal(rev, got[:, 0])
    diag = capi.dist_pairs(g, np.stack([np.arange(1000), np.arange(1000)], 1))
    assert not diag.any()


def test_rejects_bad_input():
    g = np.array([[1, 2, 3], [1, 2, 2]], dtype=np.int32)
    with pytest.raises(capi.GrsrError) as e:
        capi.dist_matrix(g)
    assert e.value.code == capi.ERR_INVALID and "duplicate gene" in str(e.value)
    with pytest.raises(capi.GrsrError):
        capi.dist_pairs(np.array([[1, 2, 3]], dtype=np.int32), [[0, 1]])
    with pytest.raises(capi.GrsrError) as e:
        capi.dist_matrix(np.array([[1, 2, 3]], dtype=np.int32), first_device=0, ngpus=99)
    assert e.value.code == capi.ERR_NODEVICE


@pytest.mark.gpu
@pytest.mark.parametrize("n", [29, 200, 500, 2000, 4500])
def test_two_tier_path_equals_block_per_pair_path(n, monkeypatch):
    """The warp-per-pair walk kernel + deferred second tier (GRSR_WALK=1) and the block-per-pair kernel
    alone (GRSR_WALK=0) must give identical {d,b,c,h,f}; the oracle checks a sample."""
    from grsr_b200 import capi
    g = U.mixed_genomes(n, 24 if n <= 2000 else 12, seed=1000 + n)
    pairs = np.array([[i, j] for i in range(len(g)) for j in range(len(g))], dtype=np.int32)
    monkeypatch.setenv("GRSR_WALK", "1")
    two = capi.dist_pairs(g, pairs, stats=True)
    monkeypatch.setenv("GRSR_WALK", "0")
    one = capi.dist_pairs(g, pairs, stats=True)
    assert np.array_equal(two, one)
    sample = pairs[:: max(1, len(pairs) // 60)]
    monkeypatch.setenv("GRSR_WALK", "1")
    assert np.array_equal(capi.dist_pairs(g, sample, stats=True), U.orc_stats(g, sample))


@pytest.mark.gpu
def test_two_tier_path_exhaustive_n6(monkeypatch):
    from grsr_b200 import capi
    monkeypatch.setenv("GRSR_WALK", "1")
    n = 6
    perms = U.all_signed_perms(n)
    g = np.vstack([np.arange(1, n + 1, dtype=np.int32)[None, :], perms])
    k = len(g)
    pairs = np.array([[0, i] for i in range(1, k)] + [[i, 0] for i in range(1, k)], dtype=np.int32)
    assert np.array_equal(capi.dist_pairs(g, pairs, stats=True), U.orc_stats(g, pairs))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [700, 3000, 9000])
def test_two_tier_path_on_degenerate_inputs(n, monkeypatch):
    """inputs the walk kernel treats specially: identical genomes (every edge an adjacency), genomes a
    few reversals apart (mostly adjacencies, short cycles without nodes), unsigned-like genomes (every
    cycle unoriented: everything is deferred), in both pair orders (private / shared pi rows)"""
    import random
    from grsr_b200 import capi
    rng = random.Random(n)
    ident = list(range(1, n + 1))
    rows = [ident, ident]
    rows += [gen.k_reversal_perm(n, k, rng) for k in (1, 2, 5, 40, 300)]
    pos = ident[:]
    rng.shuffle(pos)
    rows += [pos, [-x for x in pos], gen.random_signed_perm(n, rng)]
    base = gen.k_reversal_perm(n, 7, rng)
    rows += [[(base[abs(x) - 1] if x > 0 else -base[abs(x) - 1]) for x in gen.k_reversal_perm(n, k, rng)] for k in (1, 3, 9)]
    g = np.asarray(rows, dtype=np.int32)
    G = len(g)
    sq = np.array([[i, j] for i in range(G) for j in range(G)], dtype=np.int32)
    tri = np.array([[i, j] for j in range(G) for i in range(j)] * 3, dtype=np.int32)   # pi-sorted runs
    for pairs, shared in ((sq, "0"), (tri, "1")):
        monkeypatch.setenv("GRSR_WALK", "0")
        one = capi.dist_pairs(g, pairs, stats=True)
        monkeypatch.setenv("GRSR_WALK", "1")
        monkeypatch.setenv("GRSR_WALK_SHARED", shared)      # force the private- / shared-row kernel
        two = capi.dist_pairs(g, pairs, stats=True)
        monkeypatch.delenv("GRSR_WALK_SHARED")
        assert np.array_equal(two, one)
    small = sq[:: max(1, len(sq) // 40)]
    assert np.array_equal(capi.dist_pairs(g, small, stats=True), U.orc_stats(g, small))
