#!/usr/bin/env python
"""bench.py -- throughput of the GRIMM reversal-distance / scenario hot path on the workloads
BASELINE.json quotes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline (metric / value / e2e / roofline / cpu_baseline), config 3: all 499,500 pairs of 1000
synthetic genomes x 2000 signed blocks (seed 7).  One step = one pass of the distance path over the
whole pair list (at N > 1 each rank owns a contiguous 1/N slice of the upper triangle; no collective
on the data path; strong scaling of a fixed job).

  value          pairs/s, kernels only, inputs resident in HBM, per-step CUDA events on the launch
                 stream, L2 flushed between steps, max over ranks
  e2e            pairs/s through the call a GRIMM binding makes -- grsr_dist_matrix(genomes, G, n,
                 distmatrix, first_device 0, ngpus N), ONE process, HOST buffers in, the finished
                 G x G matrix out (INTEGRATION.md section 2.1 replaces setinvmatrix by it) -- with the
                 H2D of the genome table, validation, signed inverses, pair generation, kernels and
                 the D2H of the matrix inside the timed region.  At N > 1 rank 0 makes that call on all
                 N devices while the other ranks wait on the CPU; the one-process-per-GPU shard call
                 (grsr_dist_matrix_shard) is timed beside it as e2e_multiprocess.
  roofline       algorithmic bytes (8n+4 per pair) / kernel time vs the
                 measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline   the UNMODIFIED reference (oracle/_ref/libgrimmref.so) on all host cores, on a
                 bounded sample of the same pair list (rank 0, N = 1 only); at every N the GPU
                 matrix is compared with the reference on a sample spread over all ranks' slices

Secondary legs in the same JSON line ("scenario", run by rank 0 in-process on all N devices):
  config4        BASELINE.json config 4: 10,000 hurdle/fortress-rich pairs x 5000 blocks, the first
                 CONFIG4_STEP_CAP steps of every pair's -s scenario (the whole scenarios are hours
                 on any hardware), sharded over the N devices by the library's cross-device queue
  config5        (N = 1) one pair, n = 100,000: time to first step, steps/s over a 2000-step prefix
  cli            (N = 1) wall time of bin/grimm vs the reference binary on configs 1 and 2

--impl reference times the reference CPU implementation of the headline alone (same metric/config).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from grsr_b200 import gen  # noqa: E402

G_CFG, N_CFG, SEED_CFG = 1000, 2000, 7
METRIC = "reversal_distance_pairs_per_sec"
UNIT = "pairs/s"
BYTES_PER_PAIR = 8 * N_CFG + 4          # read pi and gamma as int32, write one int32 distance
CONFIG4_PAIRS, CONFIG4_GENES, CONFIG4_STEP_CAP = 10000, 5000, 16
CONFIG5_GENES, CONFIG5_SEED, CONFIG5_PREFIX = 100000, 5, 2000
I32P = ctypes.POINTER(ctypes.c_int)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "grimm")


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def workload_config(extra=None):
    cfg = {
        "workload": "config3: all-pairs reversal distance, 1000 genomes x 2000 signed blocks "
                    "(499500 pairs), synthetic seed 7, circular-aligned (-C -m)",
        "genomes": G_CFG, "genes": N_CFG, "pairs": G_CFG * (G_CFG - 1) // 2,
        "sharding": "contiguous slices of the column-by-column upper triangle, one per rank, no collective",
    }
    if extra:
        cfg.update(extra)
    return cfg


def make_genomes():
    g = gen.random_genomes(G_CFG, N_CFG, SEED_CFG)
    # -C: align every genome on the first gene of genome 1 (host preprocessing of the CLI,
    # reference G/circ_align.c:50); genome 1 is the identity so this is a rotation / flip per row
    out = g.copy()
    a = int(g[0, 0])
    for r in range(1, G_CFG):
        row = g[r]
        pos = int(np.nonzero(np.abs(row) == abs(a))[0][0])
        if row[pos] == a:
            out[r] = np.concatenate([row[pos:], row[:pos]])
        else:
            out[r] = np.concatenate([-row[:pos + 1][::-1], -row[pos + 1:][::-1]])
    return np.ascontiguousarray(out, dtype=np.int32)


# ---- reference CPU arm ---------------------------------------------------------------------------

def load_cpu_reference():
    """(kind, callable(genomes, pairs, threads) -> distances).  Prefers the unmodified reference
    compiled in place; falls back to the oracle port (single thread per process)."""
    if os.path.exists(REF_LIB):
        lib = ctypes.CDLL(REF_LIB)

        def run(genomes, pairs, threads):
            out = np.zeros(len(pairs), dtype=np.int32)
            lib.ref_dist_pairs(genomes.ctypes.data_as(I32P), genomes.shape[1], pairs.ctypes.data_as(I32P),
                               ctypes.c_long(len(pairs)), out.ctypes.data_as(I32P), 1, threads)
            return out
        return "reference", run
    port = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(port):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    lib = ctypes.CDLL(port)

    def run1(genomes, pairs, threads):
        out = np.zeros(len(pairs), dtype=np.int32)
        ths = []
        for t in range(threads):
            lo, hi = len(pairs) * t // threads, len(pairs) * (t + 1) // threads
            sl = np.ascontiguousarray(pairs[lo:hi])
            o = out[lo:hi]
            ths.append(threading.Thread(target=lib.orc_dist_pairs, args=(
                genomes.ctypes.data_as(I32P), genomes.shape[1], sl.ctypes.data_as(I32P),
                ctypes.c_long(hi - lo), o.ctypes.data_as(I32P), 1)))
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return out
    return "port", run1


def triangle_pairs(G):
    j, i = np.tril_indices(G, k=-1)
    return np.stack([i, j], axis=1).astype(np.int32)   # column-by-column triangle, as the GPU arm


def cpu_sample_rate(genomes, pairs, seconds: float):
    """Time the CPU reference on a bounded prefix of the pair list sized for about `seconds`."""
    kind, run = load_cpu_reference()
    cores = os.cpu_count() or 1
    probe = pairs[:min(len(pairs), 64 * cores)]
    t0 = time.perf_counter()
    run(genomes, probe, cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    want = int(min(len(pairs), max(len(probe), len(probe) / dt * seconds)))
    sample = np.ascontiguousarray(pairs[:want])
    t0 = time.perf_counter()
    d = run(genomes, sample, cores)
    dt = time.perf_counter() - t0
    return {"value": len(sample) / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "first %d of %d pairs of the workload, %.1f s on %d host threads"
                      % (len(sample), len(pairs), dt, cores)}, d, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    genomes = make_genomes()
    pairs = triangle_pairs(G_CFG)
    kind, run = load_cpu_reference()
    cores = os.cpu_count() or 1
    # size one step for a few seconds so that steps+warmup stay within minutes
    probe = np.ascontiguousarray(pairs[:64 * cores])
    t0 = time.perf_counter()
    run(genomes, probe, cores)
    rate = len(probe) / max(time.perf_counter() - t0, 1e-6)
    budget = 120.0 / max(1, args.steps + args.warmup)
    per_step = int(min(len(pairs), max(len(probe), rate * min(budget, 10.0))))
    sample = np.ascontiguousarray(pairs[:per_step])
    for _ in range(args.warmup):
        run(genomes, sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(genomes, sample, cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": workload_config({"reference_step": "first %d pairs of the list per step" % per_step}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d pairs per step x %d steps on %d host threads" % (per_step, args.steps, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks --------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.limit,clocks.mem")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw, extra = [], [], set(), [], {}
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
            try:                                       # what tells two boxes of the pool apart
                pw.append(float(f[3]))
                extra = {"power_limit_w": float(f[9]), "mem_mhz": float(f[10])}
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
               "samples": len(sm)}
        if pw:
            out["power_w"] = float(np.median(pw))
        out.update(extra)
        return out


# ---- GPU arm -------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from grsr_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU implementation")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")     # waits that must not occupy the GPUs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    genomes = make_genomes()
    P = G_CFG * (G_CFG - 1) // 2
    lo, hi = P * rank // world, P * (rank + 1) // world
    cnt = hi - lo

    # ---- value: resident inputs, kernels only -------------------------------------------------------
    stream = torch.cuda.current_stream().cuda_stream
    d_gen = torch.from_numpy(genomes).to(dev)
    d_sinv = torch.empty_like(d_gen)
    d_pairs = torch.empty((cnt, 2), dtype=torch.int32, device=dev)
    d_out = torch.empty((cnt,), dtype=torch.int32, device=dev)
    capi.signed_inverse_dev(d_gen.data_ptr(), G_CFG, N_CFG, d_sinv.data_ptr(), stream)
    capi.triangle_pairs_dev(G_CFG, lo, cnt, d_pairs.data_ptr(), stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_dev():
        capi.dist_pairs_sorted_dev(d_gen.data_ptr(), d_sinv.data_ptr(), N_CFG, d_pairs.data_ptr(), cnt,
                                   d_out.data_ptr(), 1, stream)
    launches_per_step = 2       # dist_walk_shared_kernel + dist_deferred_kernel (the scratch memset is not ours)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_region0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)                 # evict L2 between timed steps (outside the event pair)
        a.record()
        step_dev()
        b.record()
    barrier()
    t_region1 = time.perf_counter()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(float(sum(ms)))
    kernel_ms = float(np.mean(ms))
    value = P * args.steps / (total_ms / 1e3)
    resident = d_out.cpu().numpy()

    # ---- e2e, one process per GPU: the shard call with host buffers ----------------------------------
    h_gen = torch.from_numpy(genomes).pin_memory()
    g_np = h_gen.numpy()
    h_out = torch.empty((cnt,), dtype=torch.int32).pin_memory()
    o_np = h_out.numpy()
    for _ in range(warm):
        capi.dist_matrix_shard(g_np, rank, world, device=local, out=o_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        capi.dist_matrix_shard(g_np, rank, world, device=local, out=o_np)
    torch.cuda.synchronize()
    mp_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    if not np.array_equal(o_np, resident):
        raise SystemExit("device-resident and host-buffer paths disagree")
    cpu_barrier()

    # ---- e2e, the bound call: rank 0 alone, grsr_dist_matrix on all N devices ---------------------------
    line = None
    if rank == 0:
        h_mat = torch.empty((G_CFG, G_CFG), dtype=torch.int32).pin_memory()
        m_np = h_mat.numpy()
        for _ in range(warm):
            capi.dist_matrix(g_np, first_device=0, ngpus=world, out=m_np)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            capi.dist_matrix(g_np, first_device=0, ngpus=world, out=m_np)
        e2e_s = time.perf_counter() - t0
        pairs = triangle_pairs(G_CFG)
        if not np.array_equal(m_np[pairs[lo:hi, 0], pairs[lo:hi, 1]], resident):
            raise SystemExit("grsr_dist_matrix and the device-resident path disagree")
        if not (np.array_equal(m_np, m_np.T) and not m_np.diagonal().any()):
            raise SystemExit("grsr_dist_matrix: matrix not symmetric with zero diagonal")
        flat = m_np[pairs[:, 0], pairs[:, 1]]

        peak, peak_src = hbm_peak()
        achieved = cnt * BYTES_PER_PAIR / (kernel_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dist_kernel_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "kernel": "grsr::dist_walk_shared_kernel (+ dist_deferred_kernel on the deferred pairs)",
                    "kernel_ms": kernel_ms,
                    "algorithmic_bytes_per_pair": BYTES_PER_PAIR, "pairs_per_launch": cnt, "peak_source": peak_src}
        cpu = None
        if not args.no_cpu:
            if world == 1:
                cpu, d_cpu, sample = cpu_sample_rate(genomes, pairs, seconds=12.0)
                if not np.array_equal(d_cpu, flat[:len(sample)]):
                    raise SystemExit("GPU distances differ from the CPU reference on the baseline sample")
                cpu["checked"] = "GPU distances equal the CPU reference on all %d sample pairs" % len(sample)
            else:
                # every rank's slice is covered: a strided sample of the whole list against the reference
                kind, run = load_cpu_reference()
                idx = np.arange(0, P, max(1, P // 16384))
                d_cpu = run(genomes, np.ascontiguousarray(pairs[idx]), os.cpu_count() or 1)
                if not np.array_equal(d_cpu, flat[idx]):
                    raise SystemExit("GPU distances differ from the CPU reference on the cross-rank sample")
                cpu = {"checked": "grsr_dist_matrix(ngpus=%d) equals the CPU %s on %d pairs spread over all "
                                  "%d slices (the timed baseline is reported at N = 1 only)" % (world, kind, len(idx), world)}
        scen = None
        if not args.no_scenario:
            scen = {"config4": config4_leg(capi, world, no_cpu=args.no_cpu)}
            if world == 1 and not args.no_cpu:
                scen["config5"] = config5_leg(capi)
                scen["cli"] = cli_leg()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config({"l2": "flushed between timed steps with a 256 MiB write; per-step CUDA events",
                                       "threads_per_block": os.environ.get("GRSR_THREADS", "auto")}),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": P * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(genomes.nbytes) * world,
                    "d2h_bytes_per_step": int(G_CFG * G_CFG * 4), "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "grsr_dist_matrix(genomes, %d, %d, distmatrix, 0, %d): one process, pinned host buffers, "
                           "full G x G matrix out" % (G_CFG, N_CFG, world)},
            "e2e_multiprocess": {"value": P * args.steps / mp_s, "unit": UNIT, "ms_per_step": mp_s / args.steps * 1e3,
                                 "h2d_bytes_per_step": int(genomes.nbytes) * world, "d2h_bytes_per_step": int(P * 4),
                                 "api": "grsr_dist_matrix_shard, one process per GPU (max over ranks)"},
            "gpu_launches": launches_per_step * args.steps * world, "clocks": clocks, "scenario": scen,
        }
    cpu_barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- scenario legs (secondary metric of BASELINE.json: scenario reversals/s) ---------------------------

def _hurdle_rows(seeds):
    return [gen.hurdle_rich_perm(CONFIG4_GENES, s) for s in seeds]


def config4_sources(k):
    """gen.hurdle_rich_perm(5000, pair_index) for pair_index < k, built on all host cores."""
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    chunks = [list(range(s, min(k, s + 64))) for s in range(0, k, 64)]
    with ProcessPoolExecutor(max_workers=cores) as ex:
        rows = [r for part in ex.map(_hurdle_rows, chunks) for r in part]
    return np.array(rows, dtype=np.int32)


def _ref_prefix_worker(args):
    """The first `cap` steps of one reference scenario (the reference prints through a global FILE*;
    ref_scenario_prefix runs print_scenario_2 in a forked child and stops it after `cap` steps)."""
    src, dst, cap = args
    lib = ctypes.CDLL(REF_LIB)
    steps = np.zeros(2 * max(cap, 1), dtype=np.int32)
    k = lib.ref_scenario_prefix(src.ctypes.data_as(I32P), dst.ctypes.data_as(I32P), len(src),
                                steps.ctypes.data_as(I32P), cap)
    return steps[:2 * max(k, 0)].reshape(-1, 2)


def config4_leg(capi, ngpus, no_cpu=False):
    """Config 4 at its stated shape: 10,000 (source, identity) pairs of 5000 blocks, source =
    gen.hurdle_rich_perm(5000, pair index); the first CONFIG4_STEP_CAP steps of every scenario through
    grsr_scenario_prefix on `ngpus` devices (one process; the library's cross-device pair queue)."""
    from concurrent.futures import ProcessPoolExecutor
    n, k, cap = CONFIG4_GENES, CONFIG4_PAIRS, CONFIG4_STEP_CAP
    src = config4_sources(k)
    dst = np.tile(np.arange(1, n + 1, dtype=np.int32), (k, 1))
    # warm-up: the same call once (contexts on all devices, peer mappings, the library's grow-only device
    # buffers and its page-locked host arena reach their size, worker threads exist)
    capi.scenario_prefix(src, dst, cap, ngpus=ngpus)
    t0 = time.perf_counter()
    steps, done, evals = capi.scenario_prefix(src, dst, cap, ngpus=ngpus)
    dt = time.perf_counter() - t0
    tot = int(sum(len(x) for x in steps))
    peak, _ = hbm_peak()
    leg = {"workload": "config4: %d hurdle/fortress-rich pairs x %d blocks (gen.hurdle_rich_perm(5000, pair index) -> "
                       "identity), -s scenario, first %d steps of every pair" % (k, n, cap),
           "pairs": k, "genes": n, "step_cap": cap, "n_gpus": ngpus, "reversals": tot, "seconds": dt,
           "value": tot / dt, "unit": "reversals/s",
           "full_evaluations": int(evals.sum()), "evaluations_per_s": float(evals.sum()) / dt,
           "api": "grsr_scenario_prefix (host buffers in, step lists out, one process, ngpus=%d)" % ngpus,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                        "achieved": tot / dt * (12 * n + 16) / 1e9, "frac": tot / dt * (12 * n + 16) / 1e9 / peak,
                        "algorithmic_bytes_per_step": 12 * n + 16,
                        "achieved_per_evaluation_bytes": float(evals.sum()) / dt * 8 * n / 1e9}}
    if not no_cpu and os.path.exists(REF_LIB):
        cores = os.cpu_count() or 1
        m = min(k, cores)
        idx = [int(q) for q in np.linspace(0, k - 1, m)]          # pairs spread over the whole batch
        t0 = time.perf_counter()
        with ProcessPoolExecutor(max_workers=cores) as ex:
            ref = list(ex.map(_ref_prefix_worker, [(src[q], dst[q], cap) for q in idx]))
        rdt = time.perf_counter() - t0
        if not all(np.array_equal(a, steps[q]) for a, q in zip(ref, idx)):
            raise SystemExit("GPU scenario prefixes differ from the CPU reference on the config-4 sample")
        rtot = sum(len(x) for x in ref)
        leg["cpu_baseline"] = {"value": rtot / rdt, "unit": "reversals/s", "cores": cores, "kind": "reference",
                               "sample": "%d pairs spread over the batch, first %d steps each, one reference process "
                                         "per pair on %d cores, %.1f s; step lists equal the GPU's" % (m, cap, cores, rdt)}
    return leg


def config5_leg(capi):
    """Config 5: one pair, n = 100,000 random signed (seed 5) -> identity, one GPU (a single scenario
    is a chain of steps: replicas only).  Time to first step and steps/s over a fixed prefix; the
    prefix is compared with the reference's (tests/golden/scenario_prefixes.npz)."""
    import random
    n = CONFIG5_GENES
    src = np.array(gen.random_signed_perm(n, random.Random(CONFIG5_SEED)), dtype=np.int32)
    dst = np.arange(1, n + 1, dtype=np.int32)
    capi.scenario_prefix(src, dst, 1)
    t0 = time.perf_counter()
    capi.scenario_prefix(src, dst, 1)
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    steps, done, _ = capi.scenario_prefix(src, dst, CONFIG5_PREFIX)
    dt = time.perf_counter() - t0
    peak, _ = hbm_peak()
    leg = {"workload": "config5: one pair, %d blocks, random signed seed %d -> identity, first %d steps" % (n, CONFIG5_SEED, CONFIG5_PREFIX),
           "genes": n, "steps": len(steps[0]), "seconds": dt, "value": len(steps[0]) / dt, "unit": "reversals/s",
           "time_to_first_step_s": first,
           "roofline_frac_hbm": len(steps[0]) / dt * (12 * n + 16) / 1e9 / peak}
    fx = os.path.join(ROOT, "tests", "golden", "scenario_prefixes.npz")
    if os.path.exists(fx):
        want = np.load(fx)["cfg5_steps"][:CONFIG5_PREFIX]
        if not np.array_equal(steps[0][:len(want)], want):
            raise SystemExit("config 5: GPU step prefix differs from the reference fixture")
        leg["checked"] = "first %d steps equal the reference's (tests/golden/scenario_prefixes.npz)" % len(want)
    return leg


def cli_leg():
    """Drop-in latency where GRSR uses it (scr/reptA.sh:61 starts one grimm per genome pair): wall time
    of the host program against the reference binary, process start-up included."""
    from grsr_b200 import build
    ours = build.CLI
    if not os.path.exists(ours):
        return {"unavailable": "grsr_b200/bin/grimm not built"}
    ex = os.path.join(ROOT, "tests", "golden", "example_mgr_macro.txt")
    tmp = os.path.join(ROOT, "gpurun_out")
    os.makedirs(tmp, exist_ok=True)
    cfg2 = os.path.join(tmp, "bench_config2.txt")
    g2 = gen.random_genomes(25, 500, 1)
    with open(cfg2, "w") as f:
        for k, row in enumerate(g2):
            f.write(">g%d\n%s $\n" % (k + 1, " ".join(str(int(x)) for x in row)))

    def wall(binary, argv, reps, env=None):
        best, out = None, None
        for _ in range(reps):
            t0 = time.perf_counter()
            r = subprocess.run([binary] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return None, r.stderr.decode()[-200:]
            best = dt if best is None else min(best, dt)
            out = r.stdout.decode()
        return best, out

    def strip(text):
        return "\n".join(ln for ln in text.splitlines() if not ln.startswith(("Running:", "Input:")))

    out = {}
    # resident mode: one `grimm --server` keeps the CUDA context, the invocations hand themselves to it
    sock = os.path.join(tmp, "bench_grsr.sock")
    if os.path.exists(sock):
        os.unlink(sock)
    srv = subprocess.Popen([ours, "--server", sock], stderr=subprocess.DEVNULL)
    for _ in range(600):
        if os.path.exists(sock) or srv.poll() is not None:
            break
        time.sleep(0.05)
    srv_env = dict(os.environ, GRSR_SERVER=sock) if os.path.exists(sock) else None
    for name, argv in (("config1_scenario_pair_1_2", ["-f", ex, "-C", "-g", "1,2"]),
                       ("config1_matrix", ["-f", ex, "-C", "-m"]),
                       ("config2_matrix_25x500", ["-f", cfg2, "-C", "-m"])):
        t_ours, o_ours = wall(ours, argv, 5)
        leg = {"argv": " ".join(argv[2:]), "grsr_wall_s": t_ours}
        if srv_env:
            t_res, o_res = wall(ours, argv, 20, env=srv_env)
            leg["grsr_resident_wall_s"] = t_res
            leg["resident_same_output"] = (o_res == o_ours)
        if os.path.exists(REF_BIN):
            t_ref, o_ref = wall(REF_BIN, argv, 5)
            leg["reference_wall_s"] = t_ref
            if t_ours is not None and t_ref is not None:
                leg["same_output"] = strip(o_ours) == strip(o_ref)
        out[name] = leg
    # start-up alone: the cheapest command that creates the CUDA context (a 2-genome, 3-gene matrix)
    tiny = os.path.join(tmp, "bench_tiny.txt")
    with open(tiny, "w") as f:
        f.write(">a\n1 2 3 $\n>b\n1 -2 3 $\n")
    t_ours, _ = wall(ours, ["-f", tiny, "-L", "-d", "-g", "1,2"], 5)
    out["process_startup_s"] = t_ours
    srv.kill()
    srv.wait()
    out["note"] = ("grsr_wall_s = one process per invocation (CUDA initialisation and context creation dominate); "
                   "grsr_resident_wall_s = the same command line with GRSR_SERVER pointing at one `grimm --server` "
                   "process; best of the repetitions")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="grsr", choices=["grsr", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-reference legs (cpu_baseline, config5, cli)")
    ap.add_argument("--no-scenario", action="store_true", help="skip the scenario legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
