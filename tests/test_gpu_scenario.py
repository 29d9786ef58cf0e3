"""GPU parity tests of the scenario search, through the C ABI, against oracle and golden vectors."""
import json
import random

import numpy as np
import pytest

import _util as U
from grsr_b200 import capi, gen

pytestmark = pytest.mark.gpu


def _check_pairs(srcs, dsts):
    got = capi.scenario_batch(np.asarray(srcs, dtype=np.int32), np.asarray(dsts, dtype=np.int32))
    for s, d, steps in zip(srcs, dsts, got):
        want, _ = U.orc_scenario(s, d)
        assert np.array_equal(steps, want), (list(s), list(d))
        assert np.array_equal(U.apply_steps(s, steps), np.asarray(d))


def test_golden_reference_scenarios():
    cases = json.load(open(U.GOLDEN + "/scenarios_mixed.json"))
    by_n = {}
    for c in cases:
        by_n.setdefault(len(c["src"]), []).append(c)
    for n, cs in by_n.items():
        got = capi.scenario_batch(np.array([c["src"] for c in cs], dtype=np.int32),
                                  np.array([c["dst"] for c in cs], dtype=np.int32))
        for c, steps in zip(cs, got):
            assert steps.tolist() == c["steps"]


def test_published_docx_scenarios():
    """The three worked scenarios printed in the paper supplement (Additional_file_2.docx)."""
    for c in json.load(open(U.GOLDEN + "/docx_scenarios.json")):
        src, dst = c["source"], c["destination"]
        steps = capi.scenario_batch(src, dst)[0]
        assert len(steps) == len(c["step_values"])
        cur = np.array(src)
        for (s, e), (sv, ev), row in zip(steps, c["step_values"], c["rows"][1:]):
            assert (cur[s], cur[e]) == (sv, ev)
            cur[s:e + 1] = -cur[s:e + 1][::-1]
            assert cur.tolist() == row


def test_example_config1_scenarios():
    g = U.orc_circular_align(U.parse_genomes(U.EXAMPLE)[1])
    srcs, dsts = [g[0], g[0], g[1]], [g[1], g[2], g[2]]
    got = capi.scenario_batch(np.array(srcs), np.array(dsts))
    assert [len(x) for x in got] == [20, 12, 15]
    assert got[0][0].tolist() == [3, 3] and got[0][-1].tolist() == [5, 24]
    _check_pairs(srcs, dsts)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 33, 64, 100, 200])
def test_mixed_pairs(n):
    srcs, dsts = [], []
    for s in range(16):
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
    _check_pairs(srcs, dsts)


def test_random_n1000_and_round_trip_n5000():
    rng = random.Random(5)
    src, dst = gen.random_signed_perm(1000, rng), list(range(1, 1001))
    _check_pairs([src], [dst])
    # size-independent property at config-4 size: the scenario has exactly d steps and sorts
    srcs = np.array([gen.random_signed_perm(5000, rng) for _ in range(4)], dtype=np.int32)
    dsts = np.array([gen.random_signed_perm(5000, rng) for _ in range(4)], dtype=np.int32)
    got = capi.scenario_batch(srcs, dsts)
    for k in range(4):
        g = np.stack([srcs[k], dsts[k]])
        d = capi.dist_pairs(g, [[0, 1]])[0]
        assert len(got[k]) == d
        assert np.array_equal(U.apply_steps(srcs[k], got[k]), dsts[k])


def test_identical_genomes_have_empty_scenario():
    p = list(range(1, 41))
    assert len(capi.scenario_batch(p, p)[0]) == 0


def test_trial_distances_t1_counts_of_the_example():
    """`grimm -C -g 1,3 -t 1` on the shipped example: 18 breakpoint positions, 153 reversals tried,
    12 / 54 / 87 of them change the distance by -1 / 0 / +1."""
    g = U.orc_circular_align(U.parse_genomes(U.EXAMPLE)[1])
    cands = U.breakpoint_candidates(g[0], g[2])
    d2 = capi.trial_distances(g[0], g[2], cands)
    d = capi.dist_pairs(g, [[0, 2]])[0]
    assert len(cands) == 153
    assert [(d2 == d - 1).sum(), (d2 == d).sum(), (d2 == d + 1).sum()] == [12, 54, 87]
    assert np.array_equal(d2, U.orc_trial_distances(g[0], g[2], cands))


@pytest.mark.parametrize("n", [2, 9, 40, 150, 600])
def test_trial_distances_vs_oracle(n):
    for s in range(4):
        rng = random.Random(n * 10 + s)
        src = gen.hurdle_rich_perm(n, n * 10 + s) if s % 2 else gen.random_signed_perm(n, rng)
        dst = list(range(1, n + 1)) if s < 2 else gen.random_signed_perm(n, rng)
        cands = U.breakpoint_candidates(src, dst)
        if len(cands) > 4000:
            cands = cands[:: len(cands) // 4000 + 1]
        if len(cands) == 0:
            continue
        assert np.array_equal(capi.trial_distances(src, dst, cands), U.orc_trial_distances(src, dst, cands))


def test_stepwise_and_persistent_paths_agree():
    """With few pairs the library scores the candidates of a hurdle step across the whole GPU
    (grsr_stepwise.cuh); with GRSR_STEPWISE_MAX=0 everything runs in the persistent kernel.  Both
    must give the reference's step lists."""
    import os
    n = 120
    srcs = np.array([gen.hurdle_rich_perm(n, 700 + s) for s in range(6)], dtype=np.int32)
    dsts = np.tile(np.arange(1, n + 1, dtype=np.int32), (6, 1))
    a = capi.scenario_batch(srcs, dsts)
    os.environ["GRSR_STEPWISE_MAX"] = "0"
    try:
        b = capi.scenario_batch(srcs, dsts)
    finally:
        os.environ.pop("GRSR_STEPWISE_MAX", None)
    for q in range(6):
        want, _ = U.orc_scenario(srcs[q], dsts[q])
        assert np.array_equal(a[q], want) and np.array_equal(b[q], want)
