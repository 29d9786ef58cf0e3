"""Build recipe: nvcc for the CUDA library (sm_100a only), gcc for the host CLI.

Everything is built in-tree (grsr_b200/lib, grsr_b200/bin) so the artefacts travel with the
repository snapshot to the GPU box.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "grsr_b200")
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
BINDIR = os.path.join(PKG, "bin")
LIB = os.path.join(LIBDIR, "libgrsr_cuda.so")
CLI = os.path.join(BINDIR, "grimm")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str], cwd: str | None = None) -> None:
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def cuda_sources() -> list[str]:
    return [os.path.join(CSRC, f) for f in ("grsr_capi.cu", "grsr_core.cuh", "grsr_scenario.cuh", "grsr_stepwise.cuh", "grsr_walk.cuh",
                                               "grsr_unsigned.cuh", "grsr_unsigned_host.hpp")] + [
        os.path.join(ROOT, "include", "grsr.h")]


def build_cuda(force: bool = False) -> str:
    """libgrsr_cuda.so: all kernels + the C ABI, compiled for sm_100a."""
    os.makedirs(LIBDIR, exist_ok=True)
    if force or not _newer(LIB, cuda_sources()):
        _run([_nvcc(), *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "grsr_capi.cu")])
    return LIB


def build_cuda_profile(force: bool = False, tag: str = "", defines: tuple = ()) -> str:
    """Diagnostic twin of the library with per-phase cycle counters compiled in (scripts/prof_phases.py);
    tag / defines name experimental variants (e.g. tag="_x16", defines=("-DGRSR_SMALL_EXTENT=16",))."""
    out = os.path.join(LIBDIR, "libgrsr_cuda_prof%s.so" % tag)
    os.makedirs(LIBDIR, exist_ok=True)
    if force or not _newer(out, cuda_sources()):
        _run([_nvcc(), *NVCC_FLAGS, "-DGRSR_PHASE_PROFILE", *defines, "-o", out, os.path.join(CSRC, "grsr_capi.cu")])
    return out


def build_cuda_variant(tag: str, defines: tuple = (), force: bool = False) -> str:
    """Experimental twin of the library (lib/libgrsr_cuda<tag>.so) for same-box A/B runs: select it with
    GRSR_LIB=<path> (grsr_b200/capi.py), e.g. tag="_g8", defines=("-DGRSR_WALK_GROUP=8",)."""
    out = os.path.join(LIBDIR, "libgrsr_cuda%s.so" % tag)
    os.makedirs(LIBDIR, exist_ok=True)
    if force or not _newer(out, cuda_sources()):
        _run([_nvcc(), *NVCC_FLAGS, *defines, "-o", out, os.path.join(CSRC, "grsr_capi.cu")])
    return out


def cli_sources() -> list[str]:
    return [os.path.join(CSRC, f) for f in ("grimm_main.c",)] + [os.path.join(ROOT, "include", "grsr.h")]


def build_cli(force: bool = False) -> str:
    """bin/grimm: the C host program with GRIMM's command line, linked against libgrsr_cuda.so."""
    os.makedirs(BINDIR, exist_ok=True)
    build_cuda()
    if force or not _newer(CLI, cli_sources() + [LIB]):
        _run(["gcc", "-O2", "-Wall", "-Wextra", "-std=gnu99", "-I", os.path.join(ROOT, "include"),
              "-o", CLI, os.path.join(CSRC, "grimm_main.c"),
              "-L", LIBDIR, "-lgrsr_cuda", "-Wl,-rpath,$ORIGIN/../lib"])
    return CLI


DELE = os.path.join(BINDIR, "deleTransp")


def build_dele_transp(force: bool = False) -> str:
    """bin/deleTransp: host C restatement of scr/deleTransp.java (the step before grimm in reptA.sh)."""
    os.makedirs(BINDIR, exist_ok=True)
    src = os.path.join(CSRC, "dele_transp.c")
    if force or not _newer(DELE, [src]):
        _run(["gcc", "-O2", "-Wall", "-Wextra", "-std=gnu99", "-o", DELE, src])
    return DELE


def build_oracle() -> None:
    """Test infrastructure: the C restatement and, where /root/reference exists, the reference."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_emu() -> None:
    """Test infrastructure: CPU build of the kernel source through the SIMT emulator."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_cli(force)
    build_dele_transp(force)
    build_oracle()
    build_emu()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB, CLI)
