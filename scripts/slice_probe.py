"""Kernel time of each 1/8 slice of config 3's pair list on ONE GPU (what a rank of the 8-GPU run computes),
to see where the strong-scaling loss sits: python scripts/slice_probe.py [nshards]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from grsr_b200 import capi
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = bench.make_genomes()
dev = torch.device("cuda", 0)
G, n = bench.G_CFG, bench.N_CFG
P = G * (G - 1) // 2
stream = torch.cuda.current_stream().cuda_stream
d_gen = torch.from_numpy(g).to(dev); d_sinv = torch.empty_like(d_gen)
capi.signed_inverse_dev(d_gen.data_ptr(), G, n, d_sinv.data_ptr(), stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0.0
for s in range(K):
    lo, hi = P * s // K, P * (s + 1) // K
    cnt = hi - lo
    d_pairs = torch.empty((cnt, 2), dtype=torch.int32, device=dev); d_out = torch.empty((cnt,), dtype=torch.int32, device=dev)
    capi.triangle_pairs_dev(G, lo, cnt, d_pairs.data_ptr(), stream)
    for _ in range(3):
        capi.dist_pairs_sorted_dev(d_gen.data_ptr(), d_sinv.data_ptr(), n, d_pairs.data_ptr(), cnt, d_out.data_ptr(), 1, stream)
    ms = []
    for _ in range(20):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); capi.dist_pairs_sorted_dev(d_gen.data_ptr(), d_sinv.data_ptr(), n, d_pairs.data_ptr(), cnt, d_out.data_ptr(), 1, stream); b.record()
        torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    print("slice %d of %d: %6d pairs  %.3f ms  (%.2f us per pair-slot)" % (s, K, cnt, np.mean(ms), np.mean(ms) * 1e3 / cnt))
    tot += np.mean(ms)
print("sum %.3f ms" % tot)
