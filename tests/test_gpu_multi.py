"""The in-process entry points a GRIMM binding calls, on one and on several GPUs: grsr_dist_matrix
(the kernels scatter both triangle halves themselves, on ngpus > 1 into the first device's matrix
over NVLink peer mappings), grsr_stats_matrix, and the scenario calls with their cross-device pair
queue.  Multi-device cases skip on a one-GPU box."""
import os
import random

import numpy as np
import pytest

import _util as U
from grsr_b200 import capi, gen

pytestmark = pytest.mark.gpu


def _ndev():
    return capi.device_count()


def _matrix_from_oracle(g):
    G = len(g)
    pairs = capi.triangle_pairs(G)
    d = U.orc_stats(g, pairs)[:, 0]
    m = np.zeros((G, G), dtype=np.int32)
    m[pairs[:, 0], pairs[:, 1]] = d
    return m + m.T


@pytest.mark.parametrize("G,n", [(2, 5), (7, 33), (40, 200), (300, 64)])
def test_matrix_mode_equals_oracle(G, n):
    g = U.mixed_genomes(n, G, seed=G + n)
    out = np.full((G, G), -5, dtype=np.int32)            # every entry must be written
    capi.dist_matrix(g, out=out)
    assert np.array_equal(out, _matrix_from_oracle(g))


def test_matrix_equals_concatenated_shards():
    g = gen.random_genomes(200, 700, 3)
    m = capi.dist_matrix(g)
    flat = np.concatenate([capi.dist_matrix_shard(g, s, 3) for s in range(3)])
    pairs = capi.triangle_pairs(200)
    assert np.array_equal(m[pairs[:, 0], pairs[:, 1]], flat)
    assert np.array_equal(m, m.T) and not m.diagonal().any()


def test_invalid_table_is_reported_after_the_non_blocking_validation():
    g = gen.random_genomes(40, 300, 5).copy()
    g[17, 5] = g[17, 6]
    with pytest.raises(capi.GrsrError, match="genome 18: duplicate gene"):
        capi.dist_matrix(g)
    g[3, 0] = 0
    with pytest.raises(capi.GrsrError, match="genome 4: input gene == 0"):
        capi.dist_matrix(g)
    # and the next call on a good table is unaffected
    good = gen.random_genomes(40, 300, 5)
    assert np.array_equal(capi.dist_matrix(good), _matrix_from_oracle(good))
    # empty jobs validate too (reference check_genes runs before any distance)
    with pytest.raises(capi.GrsrError):
        capi.dist_pairs(g, np.zeros((0, 2), dtype=np.int32))


def test_concurrent_device_calls_on_two_streams_do_not_share_scratch():
    """Two grsr_dist_pairs_dev calls in flight on different streams of one device: each
    takes its deferred-pair list from the stream-ordered pool."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    n, G = 300, 80
    ga = U.mixed_genomes(n, G, seed=1)          # hurdle-rich rows: many deferred pairs
    gb = U.mixed_genomes(n, G, seed=2)
    pairs = capi.triangle_pairs(G)
    outs = []
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    keep = []
    for g, st in zip((ga, gb), streams):
        with torch.cuda.stream(st):
            dg = torch.from_numpy(g).to(dev, non_blocking=False)
            ds = torch.empty_like(dg)
            dp = torch.from_numpy(pairs).to(dev)
            do = torch.full((len(pairs),), -9, dtype=torch.int32, device=dev)
            keep.append((dg, ds, dp, do))
    torch.cuda.synchronize()
    for rep in range(5):
        for (dg, ds, dp, do), st in zip(keep, streams):
            capi.signed_inverse_dev(dg.data_ptr(), G, n, ds.data_ptr(), st.cuda_stream)
            capi.dist_pairs_dev(dg.data_ptr(), ds.data_ptr(), n, dp.data_ptr(), len(pairs), do.data_ptr(), 1,
                                st.cuda_stream)
    torch.cuda.synchronize()
    for g, (dg, ds, dp, do) in zip((ga, gb), keep):
        assert np.array_equal(do.cpu().numpy(), U.orc_stats(g, pairs)[:, 0])


def test_caller_device_is_left_alone():
    torch = pytest.importorskip("torch")
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(1)
    g = gen.random_genomes(30, 100, 9)
    capi.dist_matrix(g, first_device=0, ngpus=1)
    assert torch.cuda.current_device() == 1
    torch.cuda.set_device(0)


@pytest.mark.skipif("_ndev() < 2")
def test_multi_gpu_matrix_table_gather_and_errors():
    """ngpus > 1, both table paths: every device uploads the whole table (default), or (GRSR_GATHER=1)
    every device uploads and validates 1/N of the rows and the rest arrives by NVLink peer copies.  Same
    matrix; an offence in any slice is reported with its global genome number."""
    g = U.mixed_genomes(120, 75, seed=8)
    want = _matrix_from_oracle(g)
    for gather in ("0", "1"):
        os.environ["GRSR_GATHER"] = gather
        try:
            for k in range(2, _ndev() + 1):
                assert np.array_equal(capi.dist_matrix(g, ngpus=k), want), (gather, k)
            for row in (0, 37, 74):                         # first, middle and last slice
                bad = g.copy()
                bad[row, 3] = bad[row, 4]
                with pytest.raises(capi.GrsrError, match="genome %d: duplicate gene" % (row + 1)):
                    capi.dist_matrix(bad, ngpus=_ndev())
            assert np.array_equal(capi.dist_matrix(g, ngpus=_ndev()), want)      # and the next call is fine
            small = U.mixed_genomes(40, 3, seed=2)          # fewer genomes than devices on an 8-GPU box
            assert np.array_equal(capi.dist_matrix(small, ngpus=_ndev()), _matrix_from_oracle(small))
        finally:
            os.environ.pop("GRSR_GATHER", None)


@pytest.mark.skipif("_ndev() < 2")
@pytest.mark.parametrize("no_peer", [0, 1])
def test_multi_gpu_matrix_and_stats(no_peer):
    os.environ["GRSR_NO_PEER"] = str(no_peer)
    try:
        g = gen.random_genomes(160, 700, 3)
        one = capi.dist_matrix(g)
        assert np.array_equal(one, _matrix_from_oracle(g))
        for k in range(2, _ndev() + 1):
            assert np.array_equal(capi.dist_matrix(g, ngpus=k), one), k
            assert np.array_equal(capi.stats_matrix(g[:24], ngpus=k), capi.stats_matrix(g[:24])), k
        h = U.mixed_genomes(150, 90, seed=4)
        assert np.array_equal(capi.dist_matrix(h, ngpus=_ndev()), _matrix_from_oracle(h))
    finally:
        os.environ.pop("GRSR_NO_PEER", None)


@pytest.mark.skipif("_ndev() < 2")
@pytest.mark.parametrize("no_peer", [0, 1])
def test_multi_gpu_scenario_queue(no_peer):
    """All devices draw pairs from one ticket; very uneven pairs (random vs hurdle-rich vs sorted)."""
    os.environ["GRSR_NO_PEER"] = str(no_peer)
    try:
        n, k = 160, 96
        rng = random.Random(2)
        srcs, dsts = [], []
        for s in range(k):
            if s % 3 == 0:
                srcs.append(gen.random_signed_perm(n, rng)); dsts.append(gen.hurdle_rich_perm(n, s))
            elif s % 3 == 1:
                srcs.append(gen.hurdle_rich_perm(n, 1000 + s)); dsts.append(list(range(1, n + 1)))
            else:
                srcs.append(gen.k_reversal_perm(n, 3, rng)); dsts.append(list(range(1, n + 1)))
        srcs = np.array(srcs, dtype=np.int32); dsts = np.array(dsts, dtype=np.int32)
        want = [U.orc_scenario(s, d)[0] for s, d in zip(srcs, dsts)]
        for ng in sorted({2, _ndev()}):
            got = capi.scenario_batch(srcs, dsts, ngpus=ng)      # inputs read from the mapped host arena
            assert all(np.array_equal(a, b) for a, b in zip(got, want)), ng
            os.environ["GRSR_NO_MAPPED_INPUT"] = "1"             # inputs copied to every device instead
            try:
                got = capi.scenario_batch(srcs, dsts, ngpus=ng)
            finally:
                os.environ.pop("GRSR_NO_MAPPED_INPUT", None)
            assert all(np.array_equal(a, b) for a, b in zip(got, want)), ng
            pre, done, _ = capi.scenario_prefix(srcs, dsts, 7, ngpus=ng)
            assert all(np.array_equal(a, b[:7]) for a, b in zip(pre, want)), ng
    finally:
        os.environ.pop("GRSR_NO_PEER", None)


@pytest.mark.skipif("_ndev() < 2")
def test_concurrent_multi_device_calls_from_two_host_threads():
    """Two host threads in multi-device calls at once (matrix and scenarios): the calls serialise on the
    worker pool, take their per-device locks in one order, and both return the right answers."""
    import threading
    g = U.mixed_genomes(90, 60, seed=5)
    want_m = _matrix_from_oracle(g)
    rng = random.Random(9)
    srcs = np.array([gen.random_signed_perm(90, rng) for _ in range(40)], dtype=np.int32)
    dsts = np.array([gen.hurdle_rich_perm(90, s) for s in range(40)], dtype=np.int32)
    want_s = [U.orc_scenario(s, d)[0] for s, d in zip(srcs, dsts)]
    errors = []

    def matrices():
        try:
            for _ in range(6):
                assert np.array_equal(capi.dist_matrix(g, ngpus=_ndev()), want_m)
                assert np.array_equal(capi.dist_matrix(g, ngpus=1), want_m)
        except Exception as e:                               # noqa: BLE001
            errors.append(e)

    def scenarios():
        try:
            for _ in range(3):
                got = capi.scenario_batch(srcs, dsts, ngpus=_ndev())
                assert all(np.array_equal(a, b) for a, b in zip(got, want_s))
        except Exception as e:                               # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=matrices), threading.Thread(target=scenarios), threading.Thread(target=matrices)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ths), "deadlock between concurrent multi-device calls"
    assert not errors, errors
