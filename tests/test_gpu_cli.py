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


def test_resident_mode_serves_the_same_reports(grimm, tmp_path):
    """`grimm --server sock` keeps the CUDA context; GRSR_SERVER=sock grimm ... hands the invocation
    (argv, cwd, stdout, stderr) to it.  Same text, same exit status -- including the error paths that
    exit(-1) -- and a per-invocation wall time without the 0.4-1 s CUDA start-up."""
    import subprocess
    import time
    sock = str(tmp_path / "grsr.sock")
    srv = subprocess.Popen([grimm, "--server", sock], stderr=subprocess.PIPE)
    try:
        for _ in range(600):
            if os.path.exists(sock):
                break
            assert srv.poll() is None, srv.stderr.read().decode()
            time.sleep(0.05)
        env = dict(os.environ, GRSR_SERVER=sock)
        names = ["example_C_g12", "example_C_m_v", "err_dup", "example_C_g14", "nofile", "hurdles_L_g21_t1",
                 "gollan_L_u_v", "example_C_g12", "messy_C_v"]
        walls = []
        for name in names:
            case = next(c for c in CASES if c["name"] == name)
            t0 = time.perf_counter()
            r = subprocess.run([grimm] + case["args"], cwd=CLI_DIR, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            walls.append(time.perf_counter() - t0)
            want_out = open(os.path.join(CLI_DIR, name + ".out")).read()
            want_err = open(os.path.join(CLI_DIR, name + ".err")).read()
            assert r.returncode == case["status"], (name, r.stderr.decode())
            assert U.normalise_report(r.stdout.decode()) == U.normalise_report(want_out), name
            assert U.error_head(r.stderr.decode()) == U.error_head(want_err), name
        assert srv.poll() is None                      # error exits of requests did not end the server
        assert min(walls) < 0.1, walls                 # no CUDA start-up per invocation
    finally:
        srv.kill()
        srv.wait()
    # nobody listening: the invocation runs in its own process
    rc, out, _ = U.run_cli(grimm, ["-f", os.path.join("..", "example_mgr_macro.txt"), "-C", "-m"], CLI_DIR)
    assert rc == 0 and "Distance Matrix" in out
