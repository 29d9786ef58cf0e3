"""One launch of the config-3 distance path on resident data (for ncu captures of dist_walk_shared_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from grsr_b200 import capi
g = bench.make_genomes()
dev = torch.device("cuda", 0)
P = bench.G_CFG * (bench.G_CFG - 1) // 2
stream = torch.cuda.current_stream().cuda_stream
d_gen = torch.from_numpy(g).to(dev); d_sinv = torch.empty_like(d_gen)
d_pairs = torch.empty((P, 2), dtype=torch.int32, device=dev); d_out = torch.empty((P,), dtype=torch.int32, device=dev)
capi.signed_inverse_dev(d_gen.data_ptr(), bench.G_CFG, bench.N_CFG, d_sinv.data_ptr(), stream)
capi.triangle_pairs_dev(bench.G_CFG, 0, P, d_pairs.data_ptr(), stream)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    capi.dist_pairs_sorted_dev(d_gen.data_ptr(), d_sinv.data_ptr(), bench.N_CFG, d_pairs.data_ptr(), P, d_out.data_ptr(), 1, stream)
torch.cuda.synchronize()
print("ok", int(d_out.sum()))
