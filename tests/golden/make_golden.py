#!/usr/bin/env python
"""Regenerates every fixture in tests/golden/ from the reference.

Run in the build container, where /root/reference exists and oracle/_ref has been built
(`make -C oracle ref`).  The fixtures it writes are committed; the GPU box and the CPU test-suite
only read them.  Sources of truth:

  docx_scenarios.json     the three worked scenarios PUBLISHED in the paper supplement
                          /root/reference/Additional_file_2.docx (parsed, not computed)
  example_mgr_macro.txt   copy of the reference's shipped example input
                          scr/Example/2.Scaffolds/anchors_500_3000_c/mgr_macro.txt (data, not code)
  cli/*.out + cli.json    stdout/stderr/exit status of the UNMODIFIED reference binary
                          (oracle/_ref/grimm) for a list of command lines
  config2_matrix.txt      reference `-C -m` matrix for config 2 (25 x 500, seed 1)
  countperms_F6.txt       reference `-F 6 -L` and `-F 5t -L` tables (all signed permutations)
  stats_mixed.npz         reference {d,b,c,h,f} (libgrimmref.so) for seeded mixed genome sets
  scenarios_mixed.json    reference step lists for seeded (source, destination) pairs
"""
import json
import os
import random
import re
import subprocess
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _util as U  # noqa: E402
from grsr_b200 import gen  # noqa: E402

REF = "/root/reference"
GRIMM = U.REF_BIN


def write_genomes(path, genomes, names=None):
    with open(path, "w") as f:
        for k, row in enumerate(genomes):
            f.write(">%s\n" % (names[k] if names else "genome%d" % (k + 1)))
            f.write(" ".join(str(int(x)) for x in row) + " $\n")


def docx_cases():
    z = zipfile.ZipFile(os.path.join(REF, "Additional_file_2.docx"))
    x = z.read("word/document.xml").decode("utf8")
    paras = []
    for p in re.findall(r"<w:p[ >].*?</w:p>", x, flags=re.S):
        t = "".join(re.findall(r"<w:t[^>]*>(.*?)</w:t>", p, flags=re.S))
        if t.strip():
            paras.append(t)
    cases, cur = [], None
    i = 0
    while i < len(paras):
        t = paras[i]
        if t.startswith("Step 0: (Source)"):
            cur = {"rows": [[int(v) for v in paras[i + 1].split()]], "steps": []}
            cases.append(cur)
            i += 2
            continue
        m = re.match(r"Step (\d+): (-?\d+) through (-?\d+): Reversal", t)
        if m and cur is not None:
            cur["steps"].append([int(m.group(2)), int(m.group(3))])
            cur["rows"].append([int(v) for v in paras[i + 1].split()])
            i += 2
            continue
        i += 1
    out = []
    for c in cases:
        out.append({"source": c["rows"][0], "destination": c["rows"][-1],
                    "step_values": c["steps"], "rows": c["rows"]})
    return out


CLI_CASES = [
    # name, args (input path substituted for {f})
    ("example_C_m", ["-f", "{f}", "-C", "-m"]),
    ("example_C_m_v", ["-f", "{f}", "-C", "-m", "-v"]),
    ("example_C_nog", ["-f", "{f}", "-C"]),
    ("example_C_nog_s", ["-f", "{f}", "-C", "-s"]),
    ("example_L_m", ["-f", "{f}", "-L", "-m"]),
    ("example_L_m_v", ["-f", "{f}", "-L", "-m", "-v"]),
    ("example_C_g12", ["-f", "{f}", "-C", "-g", "1,2"]),
    ("example_C_g13", ["-f", "{f}", "-C", "-g", "1,3"]),
    ("example_C_g23", ["-f", "{f}", "-C", "-g", "2,3"]),
    ("example_C_g21_v", ["-f", "{f}", "-C", "-g", "2,1", "-v"]),
    ("example_C_g32_d", ["-f", "{f}", "-C", "-g", "3,2", "-d"]),
    ("example_C_g31_s", ["-f", "{f}", "-C", "-g", "3,1", "-s"]),
    ("example_L_g12_v", ["-f", "{f}", "-L", "-g", "1,2", "-v"]),
    ("example_L_g23", ["-f", "{f}", "-L", "-g", "2,3"]),
    ("example_C_g11", ["-f", "{f}", "-C", "-g", "1,1"]),
    ("example_C_g12_S4", ["-f", "{f}", "-C", "-g", "1,2", "-S", "4"]),
    ("example_C_g12_m", ["-f", "{f}", "-C", "-g", "1,2", "-m"]),
    ("example_C_g14", ["-f", "{f}", "-C", "-g", "1,4"]),
    ("example_C_gbad", ["-f", "{f}", "-C", "-g", "1"]),
    ("example_C_g13_t1", ["-f", "{f}", "-C", "-g", "1,3", "-t", "1"]),
    ("example_C_g12_t1_d_v", ["-f", "{f}", "-C", "-g", "1,2", "-t", "1", "-d", "-v"]),
    ("example_L_g32_t1_s", ["-f", "{f}", "-L", "-g", "3,2", "-t", "1", "-s"]),
    ("example_C_nog_t1", ["-f", "{f}", "-C", "-t", "1"]),
    ("example_C_t9", ["-f", "{f}", "-C", "-g", "1,2", "-t", "9"]),
    ("example_C_g13_S1", ["-f", "{f}", "-C", "-g", "1,3", "-S", "1"]),
    ("example_C_g12_S1_d_v", ["-f", "{f}", "-C", "-g", "1,2", "-S", "1", "-d", "-v"]),
    ("example_L_g32_S1", ["-f", "{f}", "-L", "-g", "3,2", "-S", "1"]),
    ("nofile", ["-C"]),
    ("missingfile", ["-f", "/nonexistent/x.txt", "-C"]),
]


def fortress_genomes(n=63):
    """identity + fortress-heavy rows (f = 1 against the identity for several of them)."""
    ident = list(range(1, n + 1))
    rows = [ident]
    for units in (3, 5, 7, 9):
        p = gen.fortress_chain(units)
        rows.append(p + list(range(len(p) + 1, n + 1)))
    seed = 0
    while len(rows) < 16:
        p = gen.hurdle_rich_perm(n, 5000 + seed)
        seed += 1
        st = U.orc_stats(np.array([ident, p], dtype=np.int32), [[0, 1]])[0]
        if st[4] == 1:
            rows.append(p)
    return np.asarray(rows, dtype=np.int32)


def run_ref(args, cwd):
    r = subprocess.run([GRIMM] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout.decode(), r.stderr.decode()


def main():
    assert os.path.exists(GRIMM), "build the reference first: make -C oracle ref"
    os.makedirs(os.path.join(HERE, "cli"), exist_ok=True)

    with open(os.path.join(HERE, "docx_scenarios.json"), "w") as f:
        json.dump(docx_cases(), f)

    # extra small inputs for CLI cases
    extra = {}
    rng = random.Random(11)
    n100 = [list(range(1, 121))] + [gen.random_signed_perm(120, rng) for _ in range(2)]
    write_genomes(os.path.join(HERE, "cli", "in_n120.txt"), n100, ["alpha", "a_rather_long_genome_name_over_twenty", "c"])
    extra["n120"] = "in_n120.txt"
    n1k = [list(range(1, 1001)), gen.k_reversal_perm(1000, 12, rng)]
    write_genomes(os.path.join(HERE, "cli", "in_n1000.txt"), n1k)
    extra["n1000"] = "in_n1000.txt"
    hur = [list(range(1, 61)), gen.hurdle_rich_perm(60, 5), gen.fortress_chain(3) + list(range(22, 61))]
    write_genomes(os.path.join(HERE, "cli", "in_hurdles.txt"), hur)
    extra["hurdles"] = "in_hurdles.txt"
    with open(os.path.join(HERE, "cli", "in_messy.txt"), "w") as f:
        f.write("# leading comment\n>first genome   \n# c\n1 -2\n 3 (4) {5}\n;\n>second\r\n-5 4   -3 2 1 $ # trailing\n")
    extra["messy"] = "in_messy.txt"
    for nm, txt in (("dup", ">a\n1 2 3 $\n>b\n1 2 2 $\n"), ("zero", ">a\n1 2 3 $\n>b\n1 0 2 $\n"),
                    ("big", ">a\n1 2 3 $\n>b\n1 7 2 $\n"), ("count", ">a\n1 2 3 $\n>b\n1 2 $\n"),
                    ("twochrom", ">a\n1 2 $ 3 $\n>b\n1 2 3 $\n"), ("badchar", ">a\n1 2 x 3 $\n>b\n1 2 3 $\n"),
                    ("empty", "")):
        with open(os.path.join(HERE, "cli", "in_%s.txt" % nm), "w") as f:
            f.write(txt)
        extra[nm] = "in_%s.txt" % nm
    # unsigned genomes (-u): the reference's own demo inputs (data, copied)
    # plus a seeded set of identity-with-a-few-reversals rows read as unsigned
    import shutil
    for nm, src in (("uns1", "unsigned1.txt"), ("uns3", "uns_3all.txt"), ("gollan", "gollan12.txt")):
        shutil.copyfile(os.path.join(REF, "GRIMM", "GRIMM-2.01", "data", src), os.path.join(HERE, "cli", "in_%s.txt" % nm))
        os.chmod(os.path.join(HERE, "cli", "in_%s.txt" % nm), 0o644)
        extra[nm] = "in_%s.txt" % nm
    rng_u = random.Random(23)
    unsr = [list(range(1, 31))] + [[abs(x) for x in gen.k_reversal_perm(30, k, rng_u)] for k in (2, 3, 5, 6)]
    write_genomes(os.path.join(HERE, "cli", "in_unsr.txt"), unsr)
    extra["unsr"] = "in_unsr.txt"
    cases = list(CLI_CASES)
    cases += [
        ("uns1_L_u", ["-f", "{uns1}", "-L", "-u"]),
        ("uns1_L_u_g12", ["-f", "{uns1}", "-L", "-u", "-g", "1,2"]),
        ("uns1_L_u_g13_v", ["-f", "{uns1}", "-L", "-u", "-g", "1,3", "-v"]),
        ("uns1_C_u_v", ["-f", "{uns1}", "-C", "-u", "-v"]),
        ("uns3_C_u", ["-f", "{uns3}", "-C", "-u"]),
        ("uns3_L_u", ["-f", "{uns3}", "-L", "-u"]),
        ("gollan_L_u_v", ["-f", "{gollan}", "-L", "-u", "-v"]),
        ("gollan_C_u_g12_v", ["-f", "{gollan}", "-C", "-u", "-g", "1,2", "-v"]),
        ("gollan_L_u_m", ["-f", "{gollan}", "-L", "-u", "-m"]),
        ("gollan_L_u_g21_d", ["-f", "{gollan}", "-L", "-u", "-g", "2,1", "-d"]),
        ("unsr_L_u", ["-f", "{unsr}", "-L", "-u"]),
        ("unsr_C_u", ["-f", "{unsr}", "-C", "-u"]),
        ("unsr_C_u_g23_v", ["-f", "{unsr}", "-C", "-u", "-g", "2,3", "-v"]),
        ("unsr_L_u_g14_t1", ["-f", "{unsr}", "-L", "-u", "-g", "1,4", "-t", "1"]),
        ("unsr_L_u_g52_S1", ["-f", "{unsr}", "-L", "-u", "-g", "5,2", "-S", "1"]),
        ("n120_C_m", ["-f", "{n120}", "-C", "-m"]),
        ("n120_L_g23_v", ["-f", "{n120}", "-L", "-g", "2,3", "-v"]),
        ("n120_C_g13", ["-f", "{n120}", "-C", "-g", "1,3"]),
        ("n1000_L", ["-f", "{n1000}", "-L"]),
        ("n1000_C_m", ["-f", "{n1000}", "-C", "-m"]),
        ("hurdles_L_m_v", ["-f", "{hurdles}", "-L", "-m", "-v"]),
        ("hurdles_L_g12_v", ["-f", "{hurdles}", "-L", "-g", "1,2", "-v"]),
        ("hurdles_L_g21", ["-f", "{hurdles}", "-L", "-g", "2,1"]),
        ("hurdles_L_g31_v", ["-f", "{hurdles}", "-L", "-g", "3,1", "-v"]),
        ("hurdles_C_g13", ["-f", "{hurdles}", "-C", "-g", "1,3"]),
        ("hurdles_L_g21_t1", ["-f", "{hurdles}", "-L", "-g", "2,1", "-t", "1"]),
        ("hurdles_L_g13_t1", ["-f", "{hurdles}", "-L", "-g", "1,3", "-t", "1"]),
        ("n120_L_g23_t1", ["-f", "{n120}", "-L", "-g", "2,3", "-t", "1"]),
        ("hurdles_L_g21_S1", ["-f", "{hurdles}", "-L", "-g", "2,1", "-S", "1"]),
        ("hurdles_L_g13_S1", ["-f", "{hurdles}", "-L", "-g", "1,3", "-S", "1"]),
        ("n120_L_g23_S1", ["-f", "{n120}", "-L", "-g", "2,3", "-S", "1"]),
        ("n120_C_g31_S1", ["-f", "{n120}", "-C", "-g", "3,1", "-S", "1"]),
        ("messy_L", ["-f", "{messy}", "-L"]),
        ("messy_C_v", ["-f", "{messy}", "-C", "-v"]),
        ("messy_L_m", ["-f", "{messy}", "-L", "-m"]),
        ("err_dup", ["-f", "{dup}", "-L"]),
        ("err_zero", ["-f", "{zero}", "-L"]),
        ("err_big", ["-f", "{big}", "-L"]),
        ("err_count", ["-f", "{count}", "-L"]),
        ("err_twochrom", ["-f", "{twochrom}", "-L"]),
        ("err_badchar", ["-f", "{badchar}", "-L"]),
        ("err_empty", ["-f", "{empty}", "-L"]),
    ]
    manifest = []
    cwd = os.path.join(HERE, "cli")
    example_rel = os.path.join("..", "example_mgr_macro.txt")
    for name, args in cases:
        a = [x.format(f=example_rel, **extra) for x in args]
        rc, out, err = run_ref(a, cwd)
        with open(os.path.join(cwd, name + ".out"), "w") as f:
            f.write(out)
        with open(os.path.join(cwd, name + ".err"), "w") as f:
            f.write(err)
        manifest.append({"name": name, "args": a, "status": rc})
    with open(os.path.join(HERE, "cli.json"), "w") as f:
        json.dump(manifest, f, indent=1)

    # config 2
    g2 = gen.random_genomes(25, 500, 1)
    tmp = "/tmp/grsr_config2.txt"
    write_genomes(tmp, g2)
    rc, out, _ = run_ref(["-f", tmp, "-C", "-m"], "/tmp")
    rows = [ln.split("\t")[1].split() for ln in out.splitlines() if ln.startswith("genome")]
    np.savetxt(os.path.join(HERE, "config2_matrix.txt"), np.array(rows, dtype=np.int32), fmt="%d")

    # -F tables
    with open(os.path.join(HERE, "countperms_F6.txt"), "w") as f:
        for a in (["-F", "6", "-L"], ["-F", "5t", "-L"]):
            rc, out, _ = run_ref(a, "/tmp")
            f.write("### grimm %s\n%s\n" % (" ".join(a), out))

    # stats fixtures from the reference library
    blobs = {}
    for n in (9, 30, 61, 128):
        g = U.mixed_genomes(n, 24, seed=n)
        pairs = np.array([[i, j] for i in range(len(g)) for j in range(len(g))], dtype=np.int32)
        blobs["g%d" % n] = g
        blobs["s%d" % n] = U.ref_stats(g, pairs)
    g = fortress_genomes()
    pairs = np.array([[i, j] for i in range(len(g)) for j in range(len(g))], dtype=np.int32)
    blobs["gfort"] = g
    blobs["sfort"] = U.ref_stats(g, pairs)
    np.savez_compressed(os.path.join(HERE, "stats_mixed.npz"), **blobs)

    scen = []
    for n in (5, 12, 29, 47, 80):
        for s in range(8):
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
            steps = U.ref_scenario(src, dst)
            scen.append({"src": list(map(int, src)), "dst": list(map(int, dst)), "steps": steps.tolist()})
    with open(os.path.join(HERE, "scenarios_mixed.json"), "w") as f:
        json.dump(scen, f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
