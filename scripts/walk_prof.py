"""One warm-up and one timed call of the two-tier distance path on config 3 (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from grsr_b200 import capi, gen
G, n = 1000, 2000
g = gen.random_genomes(G, n, 7)
dev = torch.device("cuda", 0)
P = G * (G - 1) // 2
stream = torch.cuda.current_stream().cuda_stream
d_gen = torch.from_numpy(g).to(dev)
d_sinv = torch.empty_like(d_gen)
d_pairs = torch.empty((P, 2), dtype=torch.int32, device=dev)
out = torch.empty((P,), dtype=torch.int32, device=dev)
capi.signed_inverse_dev(d_gen.data_ptr(), G, n, d_sinv.data_ptr(), stream)
capi.triangle_pairs_dev(G, 0, P, d_pairs.data_ptr(), stream)
for _ in range(int(os.environ.get("REPS", "2"))):
    capi.dist_pairs_dev(d_gen.data_ptr(), d_sinv.data_ptr(), n, d_pairs.data_ptr(), P, out.data_ptr(), 1, stream)
torch.cuda.synchronize()
print("sum", int(out.sum().item()))
