"""Same-box A/B of library variants on config 3 (all pairs of 1000 genomes x 2000 genes): for every
library given (paths, selected through GRSR_LIB in a child process) the two-tier distance path is timed
with CUDA events and its results are compared with the block-per-pair kernel's.
Usage: python scripts/walk_ab.py lib1.so [lib2.so ...] [-- KEY=VAL,KEY=VAL ...]   (extra environments per library)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from grsr_b200 import capi, gen
G, n = 1000, 2000
g = gen.random_genomes(G, n, 7)
dev = torch.device("cuda", 0)
P = G * (G - 1) // 2
stream = torch.cuda.current_stream().cuda_stream
d_gen = torch.from_numpy(g).to(dev)
d_sinv = torch.empty_like(d_gen)
d_pairs = torch.empty((P, 2), dtype=torch.int32, device=dev)
capi.signed_inverse_dev(d_gen.data_ptr(), G, n, d_sinv.data_ptr(), stream)
capi.triangle_pairs_dev(G, 0, P, d_pairs.data_ptr(), stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(reps):
    out = torch.full((P,), -1, dtype=torch.int32, device=dev)
    for _ in range(3):
        capi.dist_pairs_dev(d_gen.data_ptr(), d_sinv.data_ptr(), n, d_pairs.data_ptr(), P, out.data_ptr(), 1, stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        capi.dist_pairs_dev(d_gen.data_ptr(), d_sinv.data_ptr(), n, d_pairs.data_ptr(), P, out.data_ptr(), 1, stream)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms)), out.cpu().numpy()
os.environ["GRSR_WALK"] = "0"
_, _, base = run(1)
os.environ["GRSR_WALK"] = "1"
med, mn, out = run(15)
print("%%-34s %%-40s median %%.3f ms  min %%.3f ms  %%.4e pairs/s  equal=%%s" %% (os.path.basename(os.environ.get("GRSR_LIB", "default")), os.environ.get("AB_TAG", ""), med, mn, P / med * 1e3, bool(np.array_equal(out, base))), flush=True)
''' % ROOT

args = sys.argv[1:]
envs = [""]
if "--" in args:
    k = args.index("--")
    envs += args[k + 1:]
    args = args[:k]
for rep in range(2):                     # two passes over all variants: drift shows as a difference between them
    for lib in args:
        for e in envs:
            env = dict(os.environ)
            env["GRSR_LIB"] = os.path.abspath(lib)
            env["AB_TAG"] = e
            for kv in e.split(","):
                if kv:
                    k, v = kv.split("=")
                    env[k] = v
            subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
