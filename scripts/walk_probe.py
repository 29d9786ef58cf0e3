"""Times the two-tier distance path (walk kernel + deferred pairs) against the block-per-pair kernel
on config 3 and compares their results.  Usage: python scripts/walk_probe.py [G] [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from grsr_b200 import capi, gen

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
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


def run(env, reps=5):
    for k in ("GRSR_WALK", "GRSR_WALK_WPB", "GRSR_WALK_OCC", "GRSR_WALK_SHIFT", "GRSR_WALK_SHARED", "GRSR_WALK_SHARED_WPB", "GRSR_WALK_IDLE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = torch.full((P,), -1, dtype=torch.int32, device=dev)
    for _ in range(2):
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
    return float(np.median(ms)), out.cpu().numpy()


base_ms, base = run({"GRSR_WALK": "0"})
print("block-per-pair      %.3f ms  %.3e pairs/s" % (base_ms, P / base_ms * 1e3), flush=True)
envs = [{"GRSR_WALK": "1"}]
for sh in (3, 4):
    for idle in (4, 6, 8, 12, 16):
        envs.append({"GRSR_WALK": "1", "GRSR_WALK_SHIFT": str(sh), "GRSR_WALK_IDLE": str(idle)})
for env in envs:
    try:
        ms, out = run(env)
    except Exception as ex:
        print(env, "failed:", ex)
        continue
    print("%-60s %.3f ms  %.3e pairs/s  equal=%s" % (env, ms, P / ms * 1e3, bool(np.array_equal(out, base))), flush=True)
