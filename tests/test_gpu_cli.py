"""The host program (grsr_b200/bin/grimm) against the stdout/stderr/exit status the unmodified
reference binary produced for the same command lines (tests/golden/cli)."""
import json
import os

import pytest

import _util as U
from grsr_b200 import build

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(U.GOLDEN, "cli.json")))
CLI_DIR = os.path.join(U.GOLDEN, "cli")


@pytest.fixture(scope="module")
def grimm():
    return build.build_cli()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cli_matches_reference(grimm, case):
    rc, out, err = U.run_cli(grimm, case["args"], CLI_DIR)
    want_out = open(os.path.join(CLI_DIR, case["name"] + ".out")).read()
    want_err = open(os.path.join(CLI_DIR, case["name"] + ".err")).read()
    assert rc == case["status"]
    assert U.normalise_report(out) == U.normalise_report(want_out)
    assert U.error_head(err) == U.error_head(want_err)


def test_output_file_option(grimm, tmp_path):
    out = tmp_path / "report.txt"
    args = ["-f", os.path.join("..", "example_mgr_macro.txt"), "-C", "-g", "1,2", "-o", str(out)]
    rc, so, se = U.run_cli(grimm, args, CLI_DIR)
    assert rc == 0 and so == ""
    want = open(os.path.join(CLI_DIR, "example_C_g12.out")).read()
    assert U.normalise_report(out.read_text()) == U.normalise_report(want)
    rc, so, se = U.run_cli(grimm, args, CLI_DIR)          # refuses to overwrite (G/mcmain.c:488-493)
    assert rc == 255 and "already exists" in se


def test_reptA_style_extraction(grimm):
    """What scr/reptA.sh:62,453-454 does with the report: grep + cut."""
    rc, out, _ = U.run_cli(grimm, ["-f", os.path.join("..", "example_mgr_macro.txt"), "-C", "-g", "1,2"], CLI_DIR)
    dist = [ln.split("\t")[1] for ln in out.splitlines() if "Reversal Distance" in ln]
    assert dist == ["20"]
    step1 = [ln for ln in out.splitlines() if ln.startswith("Step 1: ")][0]
    assert step1.split("[")[1].split("]")[0] == "4" and step1.split("[")[2].split("]")[0] == "4"
