"""grsr_b200/bin/deleTransp, the C restatement of scr/deleTransp.java (the
step before `grimm` in scr/reptA.sh:58).  Java is absent from this image, so parity is pinned on what
the reference ships: scr/Example/3.Repreport/reportA.txt (tests/golden/example_reportA.txt), which
records, for the three genome pairs of the example, every transposition / inverted transposition the
Java program filtered out (with the ORIGINAL block ids, i.e. after the four block-id maps it writes)
and every reversal step GRIMM then printed (13 / 9 / 7 steps).  tests/_repta.py replays that pair loop."""
import os

import pytest

import _repta as R
import _util as U
from grsr_b200 import build

REPORT = os.path.join(U.GOLDEN, "example_reportA.txt")
LABEL = {"T": "Transposition", "IT": "Inverted Transposition", "B": "Block Interchange",
         "IB": "Inverted Block Interchange", "HIB": "Half Inverted Block Interchange"}


def _check(grimm, tmp_path):
    dele = build.build_dele_transp()
    rows = R.genome_rows(U.EXAMPLE)
    want = R.parse_reportA(REPORT)
    assert sorted(want) == [(1, 2), (1, 3), (2, 3)]
    assert [len(want[k]["reversals"]) for k in sorted(want)] == [13, 9, 7]
    for (i, j), w in sorted(want.items()):
        d = tmp_path / ("p%d_%d" % (i, j))
        d.mkdir()
        tbi, moved, steps = R.pair_report(dele, grimm, rows, i, j, str(d))
        assert [(LABEL[k], a, b) for k, a, b in moved] == w["moved"], (i, j)
        assert steps == w["reversals"], (i, j)
        assert int(tbi[0]) + len(steps) == w["total"], (i, j)
        assert int(tbi[0]) == sum(int(x) for x in tbi[1].split())


@pytest.mark.reference
@pytest.mark.skipif(not U.have_reference(), reason="oracle/_ref not built")
def test_reportA_reproduced_with_the_reference_grimm(tmp_path):
    _check(U.REF_BIN, tmp_path)


@pytest.mark.gpu
def test_reportA_reproduced_with_the_gpu_grimm(tmp_path):
    _check(build.build_cli(), tmp_path)


def test_hand_cases(tmp_path):
    """Small scaffolds worked by hand from the Java source: identical genomes merge into one block;
    one transposed block (1 3 2 4 -> 2 moved behind 3) is reported and removed."""
    import subprocess
    dele = build.build_dele_transp()

    def run(g0, g1):
        r = subprocess.run([dele, str(len(g0))] + [str(x) for x in g0] + ["$"] + [str(x) for x in g1] + ["$"],
                           cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        return r.stdout.splitlines(), open(tmp_path / "mgr_macro2.txt").read()

    out, macro = run([1, 2, 3, 4, 5], [1, 2, 3, 4, 5])
    assert out == ["0", "0 0 0 0 0"] and macro == ">genome1\n1 $\n>genome2\n1 $\n"
    assert open(tmp_path / "AftMerge.txt").read() == "1:5=1\n"
    out, macro = run([7, 8, 9, 10, 11, 12], [7, 9, 8, 10, -12, -11])
    assert out[0] == "1" and out[1] == "1 0 0 0 0" and out[2].startswith("T ")
    assert macro.startswith(">genome1\n") and macro.count("$") == 2
