"""Host-side logic of the multi-GPU path, on CPU: the pair-list split is a pure function of
(P, shard, num_shards), shards tile the column-by-column triangle exactly, and a world_size-2 gloo
job that gathers per-rank shards reproduces the whole matrix (the oracle stands in for the kernel
here -- this checks the plumbing, not the arithmetic)."""
import os
import subprocess
import sys

import numpy as np

import _util as U
from grsr_b200 import capi, gen


def test_triangle_order_and_shard_tiling():
    for G in (2, 3, 7, 40):
        pairs = capi.triangle_pairs(G)
        P = G * (G - 1) // 2
        assert len(pairs) == P
        idx = pairs[:, 1].astype(np.int64) * (pairs[:, 1] - 1) // 2 + pairs[:, 0]
        assert np.array_equal(idx, np.arange(P))
        for k in (1, 2, 3, 8):
            cuts = [P * s // k for s in range(k + 1)]
            assert cuts[0] == 0 and cuts[-1] == P and all(a <= b for a, b in zip(cuts, cuts[1:]))


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GRSR_ROOT"]); sys.path.insert(0, os.path.join(os.environ["GRSR_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import _util as U
from grsr_b200 import capi, gen
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
G, n = 30, 40
g = gen.random_genomes(G, n, 11)
pairs = capi.triangle_pairs(G)
P = len(pairs)
lo, hi = P * rank // world, P * (rank + 1) // world
mine = torch.from_numpy(U.orc_stats(g, pairs[lo:hi])[:, 0].copy())
sizes = [P * (r + 1) // world - P * r // world for r in range(world)]
bufs = [torch.empty(s, dtype=torch.int32) for s in sizes]
dist.all_gather(bufs, mine) if len(set(sizes)) == 1 else [dist.broadcast(bufs[r] if r != rank else mine, r) for r in range(world)]
if len(set(sizes)) != 1: bufs[rank] = mine
flat = torch.cat(bufs).numpy()
if rank == 0:
    want = U.orc_stats(g, pairs)[:, 0]
    assert np.array_equal(flat, want)
    m = np.zeros((G, G), dtype=np.int32); m[pairs[:, 0], pairs[:, 1]] = flat; m = m + m.T
    assert np.array_equal(m, m.T) and not m.diagonal().any()
    print("SHARD_OK", world, P)
dist.destroy_process_group()
'''


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GRSR_ROOT=U.ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "SHARD_OK 2 435" in r.stdout


def test_bench_workload_is_the_named_config():
    sys.path.insert(0, U.ROOT)
    import bench
    assert (bench.G_CFG, bench.N_CFG, bench.SEED_CFG) == (1000, 2000, 7)
    assert bench.BYTES_PER_PAIR == 8 * 2000 + 4
    g = bench.make_genomes()
    assert g.shape == (1000, 2000) and g.dtype == np.int32
    # the -C alignment bench.py applies equals the reference's circular_align (via the oracle)
    assert np.array_equal(g, U.orc_circular_align(gen.random_genomes(1000, 2000, 7)))
    assert (g[:, 0] == 1).all()
