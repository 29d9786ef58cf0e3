"""CPU-side checks of the boundary: the library loads, exports every symbol include/grsr.h
declares, refuses to compute without a device, and the host program's input/option errors match the
reference's text and exit status."""
import ctypes
import json
import os
import re

import pytest

import _util as U
from grsr_b200 import build, capi


def _declared_symbols():
    text = open(os.path.join(U.ROOT, "include", "grsr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grsr_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build.build_cuda())
    names = _declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(capi.SIGNATURES)
    assert capi.lib().grsr_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import numpy as np
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = np.array([[1, 2, 3], [3, -2, 1]], dtype=np.int32)
    for call in (lambda: capi.dist_matrix(g), lambda: capi.dist_pairs(g, [[0, 1]]),
                 lambda: capi.stats_matrix(g), lambda: capi.scenario_batch(g[0], g[1]),
                 lambda: capi.dist_matrix_shard(g, 0, 1), lambda: capi.scenario_prefix(g[0], g[1], 2),
                 lambda: capi.unsigned_dist(g[0], g[1]), lambda: capi.trial_distances(g[0], g[1], [[0, 1]])):
        with pytest.raises(capi.GrsrError) as e:
            call()
        assert e.value.code == capi.ERR_NODEVICE


def test_product_does_not_reference_the_oracle():
    """Nothing under grsr_b200/ or include/ may include, link or import oracle/ or the emulator."""
    bad = []
    for base in ("grsr_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(U.ROOT, base)):
            for fn in files:
                if not fn.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                    continue
                path = os.path.join(dirpath, fn)
                text = open(path).read()
                if fn == "build.py":
                    text = text.split("def build_oracle", 1)[0]
                for m in re.finditer(r"(#include|import|from|CDLL|dlopen)[^\n]*(oracle|cuda_emu|libgrsr_emu)", text):
                    bad.append((path, m.group(0)))
    assert not bad, bad


CASES = [c for c in json.load(open(os.path.join(U.GOLDEN, "cli.json"))) if c["status"] != 0]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cli_error_paths_match_reference(case):
    grimm = build.build_cli()
    cli_dir = os.path.join(U.GOLDEN, "cli")
    rc, out, err = U.run_cli(grimm, case["args"], cli_dir)
    want_err = open(os.path.join(cli_dir, case["name"] + ".err")).read()
    assert rc == case["status"] == 255
    assert out == ""
    assert U.error_head(err) == U.error_head(want_err)


def test_cli_refuses_out_of_scope_modes():
    grimm = build.build_cli()
    cli_dir = os.path.join(U.GOLDEN, "cli")
    f = os.path.join("..", "example_mgr_macro.txt")
    for args in (["-f", f], ["-f", f, "-C", "-U", "5"], ["-f", f, "-L", "-g", "1,2", "-t", "2"], ["-f", f, "-C", "-W", "x"]):
        rc, out, err = U.run_cli(grimm, args, cli_dir)
        assert rc == 255 and "outside the path" in err
