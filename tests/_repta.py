"""The pair loop of GRSR's scr/reptA.sh (:53-62, :86-95, :447-468) restated for tests: for every genome
pair run deleTransp (our C restatement of scr/deleTransp.java), then `grimm -f mgr_macro2.txt -C -g 1,2`,
then map the reported blocks back to the original block ids through the four map files exactly as
scr/getTBIBlk.sh, getleftBlk.sh and getrightBlk.sh do (grep "=<id>$" / cut).  Returns what
reportA.txt prints for the pair, minus the BLAST repeat columns (bl2seq is outside the path)."""
import os
import re
import subprocess


def _map_file(path):
    """lines 'lhs=rhs' -> {rhs: lhs} (the scripts grep "=<rhs>$" and cut the lhs)."""
    out = {}
    for ln in open(path).read().split():
        lhs, rhs = ln.rsplit("=", 1)
        out.setdefault(int(rhs), lhs)        # grep prints every match; the scripts take the first token
    return out


def _chain(workdir, blk, take_right):
    """getleftBlk.sh / getrightBlk.sh: AftFilter -> BefFilter -> AftMerge (first or last) -> BefMerge."""
    r = int(_map_file(os.path.join(workdir, "AftFilter.txt"))[blk])
    r = int(_map_file(os.path.join(workdir, "BefFilter.txt"))[r])
    first, last = _map_file(os.path.join(workdir, "AftMerge.txt"))[r].split(":")
    r = int(last if take_right else first)
    return int(_map_file(os.path.join(workdir, "BefMerge.txt"))[r])


def _tbi_block(workdir, blk):
    """getTBIBlk.sh: BefFilter -> AftMerge (first, last) -> BefMerge."""
    r = int(_map_file(os.path.join(workdir, "BefFilter.txt"))[blk])
    first, last = _map_file(os.path.join(workdir, "AftMerge.txt"))[r].split(":")
    bm = _map_file(os.path.join(workdir, "BefMerge.txt"))
    return int(bm[int(first)]), int(bm[int(last)])


def genome_rows(mgr_macro_path):
    """what deleTransp.sh extracts with grep -A2 '>genomeK$' | sed 1d | sed 1d: the line of blocks"""
    rows, lines = [], open(mgr_macro_path).read().splitlines()
    for k, ln in enumerate(lines):
        if re.fullmatch(r">genome\d+", ln.strip()):
            rows.append(lines[k + 2].split())
    return rows


def pair_report(dele_transp, grimm, rows, i, j, workdir):
    """(tbi summary lines, [(kind, first, last)], [(left, right)] reversal steps) for genomes i -> j (1-based)."""
    g0, g1 = rows[i - 1], rows[j - 1]
    nblk = len(g0) - 1                                   # the trailing "$" is not a block
    r = subprocess.run([dele_transp, str(nblk)] + g0 + g1, cwd=workdir, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    tbi = r.stdout.splitlines()
    moved = []
    for ln in tbi[2:]:
        kind, *blks = ln.split()
        for b in blks:
            moved.append((kind,) + _tbi_block(workdir, int(b)))
    g = subprocess.run([grimm, "-f", "mgr_macro2.txt", "-C", "-g", "1,2"], cwd=workdir,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert g.returncode == 0, g.stderr
    dist = int(re.search(r"Reversal Distance:\s*\t(\d+)", g.stdout).group(1))
    steps = []
    for k in range(1, dist + 1):
        ln = next(x for x in g.stdout.splitlines() if x.startswith("Step %d: " % k))
        lb = int(ln.split("[")[1].split("]")[0])
        rb = int(ln.split("[")[2].split("]")[0])
        lb = _chain(workdir, lb, False) if lb > 0 else -_chain(workdir, -lb, True)      # reptA.sh:453-468
        rb = _chain(workdir, rb, True) if rb > 0 else -_chain(workdir, -rb, False)
        steps.append((lb, rb))
    return tbi, moved, steps


def parse_reportA(path):
    """{(i, j): {"total": n, "moved": [(label, first, last)], "reversals": [(l, r)]}} from reportA.txt"""
    out, cur, section = {}, None, None
    for ln in open(path).read().splitlines():
        m = re.match(r">From Genome (\d+) to (\d+), total rearrangment step\(s\): (\d+)", ln)
        if m:
            cur = out.setdefault((int(m.group(1)), int(m.group(2))), {"total": int(m.group(3)), "moved": [], "reversals": []})
            continue
        m = re.match(r"\*(.+): (\d+) step\(s\):", ln)
        if m:
            section = m.group(1)
            continue
        m = re.match(r"Step \d+: (-?\d+) through (-?\d+) ", ln)
        if m and cur is not None:
            pair = (int(m.group(1)), int(m.group(2)))
            if section == "Reversal":
                cur["reversals"].append(pair)
            else:
                cur["moved"].append((section,) + pair)
    return out
