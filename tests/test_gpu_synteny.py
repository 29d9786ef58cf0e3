"""GRIMM-Synteny as a caller of the path.  oracle/_ref/grimm_synt_grsr is
GRIMM-Synteny built from ITS OWN sources plus grsr_b200/csrc/grimm_synteny_shim.c, linked with
-Wl,--wrap=mcdist_noncircular against libgrsr_cuda.so (recipe: oracle/Makefile); every unichromosomal
mcdist_noncircular call (gscomp.c:1479) then runs on the GPU.  Its report on the shipped example must
equal the unmodified program's and the report the reference ships
(scr/Example/2.Scaffolds/anchors_500_3000_c/report.txt), time stamps aside."""
import os
import re
import subprocess

import pytest

import _util as U

pytestmark = pytest.mark.gpu

GS_REF = os.path.join(U.ROOT, "oracle", "_ref", "grimm_synt_ref")
GS_GRSR = os.path.join(U.ROOT, "oracle", "_ref", "grimm_synt_grsr")
COORDS = os.path.join(U.GOLDEN, "example_unique_coords.txt")
SHIPPED = os.path.join(U.GOLDEN, "example_synteny_report.txt")

STAMP = re.compile(r"^[A-Z][a-z]{2} [A-Z][a-z]{2} +\d+ \d\d:\d\d:\d\d \d{4}: ")


def _run(binary, outdir):
    os.makedirs(outdir, exist_ok=True)
    r = subprocess.run([binary, "-f", COORDS, "-d", str(outdir), "-c", "-Q", "-m", "500", "-g", "3000"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-500:]
    return {name: open(os.path.join(outdir, name)).read() for name in sorted(os.listdir(outdir))}


def _strip(report):
    """time stamps and the two lines that echo the command-line paths removed"""
    return "\n".join(STAMP.sub("", ln) for ln in report.splitlines()
                     if not ln.startswith(("Input file:", "Output directory:")))


@pytest.mark.skipif(not (os.path.exists(GS_REF) and os.path.exists(GS_GRSR)),
                    reason="oracle/_ref/grimm_synt_* not built (make -C oracle synteny)")
def test_grimm_synteny_with_the_shim_reproduces_the_example(tmp_path):
    ref = _run(GS_REF, tmp_path / "ref")
    got = _run(GS_GRSR, tmp_path / "grsr")
    assert set(ref) == set(got) >= {"report.txt", "mgr_macro.txt", "blocks.txt"}
    for name in ref:
        assert _strip(got[name]) == _strip(ref[name]), name
    assert _strip(got["report.txt"]) == _strip(open(SHIPPED).read())
    # the blocks it hands to GRIMM are the hot path's own example input
    assert got["mgr_macro.txt"] == open(U.EXAMPLE).read()


CHECK = os.path.join(U.ROOT, "oracle", "_ref", "synteny_matrix_check")


@pytest.mark.skipif(not os.path.exists(CHECK), reason="oracle/_ref/synteny_matrix_check not built")
def test_shim_batched_matrix_and_single_pair_equal_the_reference_function():
    """gs_grsr_micro_matrix (one GPU call per block) and the wrapped mcdist_noncircular (one per pair)
    against the reference's own mcdist_noncircular linked into the same program (oracle/synteny_matrix_check.c)."""
    r = subprocess.run([CHECK], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK 60"), r.stdout + r.stderr
